#!/bin/bash
# round 2, GPU pass g: pair accumulators in OP_DOT, single-witness calls of Boolean graphs through the bit plan
O=gpurun_out/r02g; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.log
P="python tools/gpu_probe.py --no-imad --reps 3"
GW_BATCH=75776 timeout 300 $P --circuits circuit9_authV2 >> $O/probe_authv2.jsonl 2>> $O/probe.err; cut -c1-120 $O/probe_authv2.jsonl
timeout 300 $P --circuits circuit7_poseidon4,circuit5_poseidon --batch 65536 >> $O/probe_small.jsonl 2>> $O/probe.err; cut -c1-120 $O/probe_small.jsonl
timeout 300 python tools/gpu_latency.py --reps 30 > $O/latency.jsonl 2>> $O/probe.err; cut -c1-200 $O/latency.jsonl
