#!/bin/bash
# round 2, GPU pass t: watchdog test, full suite, parser throughput on the box's host cores
O=gpurun_out/r02t; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
nproc > $O/nproc.txt
timeout 300 python tools/parse_probe.py circuit9_authV2 60000 > $O/parse_probe.jsonl 2>> $O/probe.err; cat $O/parse_probe.jsonl
timeout 300 python tools/gpu_latency.py --reps 20 --circuits circuit9_authV2,circuit8_sha256_512 > $O/latency.jsonl 2>> $O/probe.err; cut -c1-200 $O/latency.jsonl
