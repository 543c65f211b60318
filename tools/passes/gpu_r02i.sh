#!/bin/bash
# round 2, GPU pass i: single-set bit kernel with a TMA-fed header ring
O=gpurun_out/r02i; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "bit or latency or golden" > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.log
timeout 300 python tools/gpu_latency.py --reps 30 --circuits circuit8_sha256_512,circuit6_num2bits > $O/latency.jsonl 2>> $O/probe.err; cut -c1-200 $O/latency.jsonl
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/gpu_latency.py --reps 1 --circuits circuit8_sha256_512 > $O/memcheck_sha_single.log 2>&1; echo "memcheck rc=$?"; tail -3 $O/memcheck_sha_single.log
