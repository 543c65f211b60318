#!/bin/bash
# round 2, GPU pass p: full GPU suite on the dataflow default, latency table (both kernels), per-packet clocks, racecheck of both latency kernels
O=gpurun_out/r02p; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 300 python tools/gpu_latency.py --reps 30 > $O/latency.jsonl 2>> $O/probe.err; cut -c1-200 $O/latency.jsonl
GW_LAT_BIT=0 timeout 300 python tools/gpu_latency.py --reps 20 --circuits circuit6_num2bits,circuit8_sha256_512 > $O/latency_generic_dataflow.jsonl 2>> $O/probe.err; cut -c1-200 $O/latency_generic_dataflow.jsonl
GW_LAT_MODE=level GW_LAT_BIT=0 timeout 300 python tools/gpu_latency.py --reps 20 > $O/latency_level.jsonl 2>> $O/probe.err; cut -c1-200 $O/latency_level.jsonl
GW_LIB_PATH=$PWD/circom-witnesscalc_b200/lib_variants/libcwc_prof.so GW_LAT_CLOCKS=1 GW_LAT_CLOCKS_FILE=$O/clocks_authv2.txt timeout 300 python tools/gpu_latency.py --reps 2 --circuits circuit9_authV2 > $O/latency_prof.jsonl 2>> $O/probe.err
gzip -f $O/clocks_authv2.txt
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_probe.py > $O/sanitizer_racecheck_dataflow.log 2>&1; echo "racecheck dataflow rc=$?"; grep "RACECHECK SUMMARY" $O/sanitizer_racecheck_dataflow.log
GW_LAT_MODE=level timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_probe.py > $O/sanitizer_racecheck_level.log 2>&1; echo "racecheck level rc=$?"; grep "RACECHECK SUMMARY" $O/sanitizer_racecheck_level.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python tools/sanitize_probe.py > $O/sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?"; grep "ERROR SUMMARY" $O/sanitizer_synccheck.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_probe.py --big > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; grep "ERROR SUMMARY" $O/sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool initcheck --error-exitcode 9 python tools/sanitize_probe.py > $O/sanitizer_initcheck.log 2>&1; echo "initcheck rc=$?"; grep "ERROR SUMMARY" $O/sanitizer_initcheck.log
