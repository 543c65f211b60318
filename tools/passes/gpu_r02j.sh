#!/bin/bash
# round 2, GPU pass j: sanitizers on the current kernels (incl. single-set bit kernel, wide expansion), ncu captures, launch list of the bench
O=gpurun_out/r02j; mkdir -p $O
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_probe.py --big > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $O/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_probe.py > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $O/sanitizer_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python tools/sanitize_probe.py > $O/sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?" | tee -a $O/sanitizer_synccheck.log
timeout 600 compute-sanitizer --tool initcheck --error-exitcode 9 python tools/sanitize_probe.py > $O/sanitizer_initcheck.log 2>&1; echo "initcheck rc=$?" | tee -a $O/sanitizer_initcheck.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > $O/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:eval_batch -s 1 -c 1 -o $O/prof_authv2 python tools/gpu_probe.py --circuits circuit9_authV2 --batch 75776 --reps 1 --no-imad > $O/ncu_authv2.log 2>&1; echo "ncu authv2 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bit_ -s 3 -c 3 -o $O/prof_sha_bit python tools/gpu_probe.py --circuits circuit8_sha256_512 --batch 16384 --reps 2 --no-imad > $O/ncu_sha.log 2>&1; echo "ncu sha rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bit_ -s 4 -c 4 -o $O/prof_num2bits_bit python tools/gpu_probe.py --circuits circuit6_num2bits --batch 1048576 --reps 2 --no-imad > $O/ncu_num2bits.log 2>&1; echo "ncu num2bits rc=$?"
ls -la $O
