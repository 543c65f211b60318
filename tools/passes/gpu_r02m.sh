#!/bin/bash
# round 2, GPU pass m: dataflow latency kernel (per-warp packet streams, wait vectors) against the level kernel
O=gpurun_out/r02m; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q -k "latency or poseidon_like or golden or drop_in" > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.log
GW_LAT_BIT=0 timeout 300 python tools/gpu_latency.py --reps 20 > $O/latency_dataflow.jsonl 2>> $O/probe.err; cut -c1-200 $O/latency_dataflow.jsonl
GW_LAT_BIT=0 GW_LAT_MODE=level timeout 300 python tools/gpu_latency.py --reps 20 > $O/latency_level.jsonl 2>> $O/probe.err; cut -c1-200 $O/latency_level.jsonl
for W in 4 6; do echo "== warps $W" >> $O/latency_sweep.jsonl; GW_LAT_WARPS=$W timeout 300 python tools/gpu_latency.py --reps 10 --circuits circuit9_authV2 >> $O/latency_sweep.jsonl 2>> $O/probe.err; done; cut -c1-200 $O/latency_sweep.jsonl
