#!/bin/bash
# round 2, GPU pass q3: warp-uniform addressing in the ring / dataflow kernels
O=gpurun_out/r02q; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "bit or latency or config3 or config2" > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 300 python tools/gpu_latency.py --reps 20 > $O/latency.jsonl 2>> $O/probe.err; cut -c1-200 $O/latency.jsonl
timeout 300 python tools/gpu_probe.py --no-imad --reps 5 --circuits circuit8_sha256_512 --batch 16384 > $O/probe_sha.jsonl 2>> $O/probe.err; cut -c1-110 $O/probe_sha.jsonl
