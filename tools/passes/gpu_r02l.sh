#!/bin/bash
# round 2, GPU pass l: field inputs packed by a warp bit transpose; ring heuristic
O=gpurun_out/r02l; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "bit or latency or config3 or config2 or golden" > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
P="python tools/gpu_probe.py --no-imad --reps 5"
timeout 300 $P --circuits circuit6_num2bits --batch 65536 >> $O/probe_small.jsonl 2>> $O/probe.err; timeout 300 $P --circuits circuit6_num2bits --batch 1048576 >> $O/probe_small.jsonl 2>> $O/probe.err
timeout 300 $P --circuits circuit8_sha256_512 --batch 16384 >> $O/probe_small.jsonl 2>> $O/probe.err
cut -c1-110 $O/probe_small.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file $O/launches_num2bits.csv python tools/gpu_probe.py --circuits circuit6_num2bits --batch 65536 --reps 2 --no-imad > /dev/null 2>&1; grep -v "^==" $O/launches_num2bits.csv | cut -d, -f5,15- | tail -8
