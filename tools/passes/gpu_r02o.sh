#!/bin/bash
# round 2, GPU pass o: dataflow latency kernel, sweep of warp counts / exclusive sub-partition; sanitizers on it
O=gpurun_out/r02o; mkdir -p $O
run() { echo "== $*" >> $O/sweep.jsonl; env "$@" timeout 300 python tools/gpu_latency.py --reps 10 --circuits circuit9_authV2 >> $O/sweep.jsonl 2>> $O/probe.err; tail -1 $O/sweep.jsonl | cut -c1-120; }
run GW_LAT_EXCL=0 GW_LAT_WARPS=7 GW_LAT_SLOW_WARPS=3
run GW_LAT_EXCL=0 GW_LAT_WARPS=8 GW_LAT_SLOW_WARPS=4
run GW_LAT_EXCL=0 GW_LAT_WARPS=6 GW_LAT_SLOW_WARPS=3
run GW_LAT_EXCL=0 GW_LAT_WARPS=5 GW_LAT_SLOW_WARPS=3
run GW_LAT_EXCL=0 GW_LAT_WARPS=8 GW_LAT_SLOW_WARPS=3
run GW_LAT_EXCL=1 GW_LAT_WARPS=7 GW_LAT_SLOW_WARPS=3
run GW_LAT_EXCL=0 GW_LAT_WARPS=7 GW_LAT_SLOW_WARPS=2
run GW_LAT_MODE=level
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_probe.py > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $O/sanitizer_racecheck.log; grep -c "Race reported" $O/sanitizer_racecheck.log; grep "RACECHECK SUMMARY" $O/sanitizer_racecheck.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_probe.py --big > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $O/sanitizer_memcheck.log
