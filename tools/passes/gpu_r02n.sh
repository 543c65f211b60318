#!/bin/bash
# round 2, GPU pass n: per-packet clocks of the dataflow latency kernel (profiling build), correctness of the fence-free publish
O=gpurun_out/r02n; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "latency or poseidon_like" > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 300 python tools/gpu_latency.py --reps 20 --circuits circuit9_authV2,circuit5_poseidon > $O/latency.jsonl 2>> $O/probe.err; cut -c1-200 $O/latency.jsonl
GW_LAT_EXCL=0 timeout 300 python tools/gpu_latency.py --reps 20 --circuits circuit9_authV2 > $O/latency_noexcl.jsonl 2>> $O/probe.err; cut -c1-200 $O/latency_noexcl.jsonl
GW_LIB_PATH=$PWD/circom-witnesscalc_b200/lib_variants/libcwc_prof.so GW_LAT_CLOCKS=1 GW_LAT_CLOCKS_FILE=$O/clocks_authv2.txt timeout 300 python tools/gpu_latency.py --reps 2 --circuits circuit9_authV2 > $O/latency_prof.jsonl 2>> $O/probe.err; cut -c1-200 $O/latency_prof.jsonl
gzip -f $O/clocks_authv2.txt
