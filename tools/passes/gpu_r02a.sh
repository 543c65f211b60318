#!/bin/bash
# round 2, GPU pass a: tests, bench (both arms), launch list, sanitizers
O=gpurun_out/r02a; mkdir -p $O
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
nproc > $O/nproc.txt; free -g >> $O/nproc.txt; numactl -H >> $O/nproc.txt 2>&1; nvidia-smi topo -m >> $O/nproc.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2>> $O/bench.err
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_probe.py --big > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $O/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_probe.py > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $O/sanitizer_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python tools/sanitize_probe.py > $O/sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?" | tee -a $O/sanitizer_synccheck.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-configs > $O/bench_under_ncu.log 2>&1
head -c 3000 $O/bench.json; echo; tail -3 $O/bench.err
