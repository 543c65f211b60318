#!/bin/bash
# kernel-only thread-count sweep of the batch kernel on one graph (witness rows wrap so HBM capacity is not the limit),
# then one full ncu capture.  usage: bash tools/gpu_sweep.sh <tag> [circuit] [ncu_threads]
TAG=${1:-sweep}; CIRCUIT=${2:-circuit9_authV2}; NT=${3:-352}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,memory.total,memory.used --format=csv > $OUT/smi.csv 2>&1
for T in 128 256 320 352 384 448 512; do
  GW_DEBUG_OUT_WRAP=4096 timeout 300 python tools/gpu_probe.py --no-imad --circuits $CIRCUIT --batch $((148*T)) --reps 2 2>&1 | \
    python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print('T', $T, d['circuit'], 'B', d['B'], 'ms', d['ms'], 'wit/s', d['witness_per_s'])
" | tee -a $OUT/sweep.log
done
GW_DEBUG_OUT_WRAP=4096 timeout 900 ncu --set full --clock-control none --import-source on -k regex:eval_batch -s 1 -c 1 -o $OUT/prof_${CIRCUIT}_T$NT \
  python tools/gpu_probe.py --circuits $CIRCUIT --batch $((148*NT)) --reps 1 --no-imad > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT
