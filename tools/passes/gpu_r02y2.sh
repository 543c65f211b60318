#!/bin/bash
# round 2, GPU pass y2: ncu captures of the bit kernels of the final build (header ring, warp transpose)
O=gpurun_out/r02y; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bit_ -s 2 -c 2 -o $O/prof_sha_bit python tools/gpu_probe.py --circuits circuit8_sha256_512 --batch 16384 --reps 2 --no-imad > $O/ncu_sha.log 2>&1; echo "ncu sha rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bit_ -s 3 -c 3 -o $O/prof_num2bits_bit python tools/gpu_probe.py --circuits circuit6_num2bits --batch 65536 --reps 2 --no-imad > $O/ncu_num2bits.log 2>&1; echo "ncu num2bits rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file $O/launches_sha.csv python tools/gpu_probe.py --circuits circuit8_sha256_512 --batch 16384 --reps 2 --no-imad > /dev/null 2>&1
