#!/bin/bash
# round 2, GPU pass x2: S-box link rewrite chosen by the timing model; forced / off in the test variants
O=gpurun_out/r02x; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 300 python tools/gpu_latency.py --reps 30 > $O/latency_final.jsonl 2>> $O/probe.err; cut -c1-170 $O/latency_final.jsonl
timeout 300 python tools/gpu_latency.py --reps 20 --circuits circuit7_poseidon4,poseidon2 >> $O/latency_final.jsonl 2>> $O/probe.err; tail -2 $O/latency_final.jsonl | cut -c1-170
