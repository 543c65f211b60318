#!/bin/bash
# round 2, GPU pass f: bit-sliced path with field inputs / wide witness values (Num2Bits), full GPU test suite, small-graph probes
O=gpurun_out/r02f; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.log
P="python tools/gpu_probe.py --no-imad --reps 5"
timeout 300 $P --circuits circuit6_num2bits,circuit7_poseidon4,circuit5_poseidon --batch 65536 >> $O/probe_small.jsonl 2>> $O/probe.err
timeout 300 $P --circuits circuit6_num2bits --batch 1048576 >> $O/probe_small.jsonl 2>> $O/probe.err
GW_BITSLICE=0 timeout 300 $P --circuits circuit6_num2bits --batch 65536 >> $O/probe_small.jsonl 2>> $O/probe.err
cut -c1-150 $O/probe_small.jsonl
timeout 300 python tools/gpu_latency.py > $O/latency.jsonl 2>> $O/probe.err; cut -c1-200 $O/latency.jsonl
