#!/bin/bash
# One short gpurun call: GPU parity tests, a thread-count sweep of the batch kernel, the headline bench.
# usage: bash tools/gpu_quick.sh <tag> [sweep threads...]
TAG=${1:-q}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
for T in "$@"; do
  GW_DEBUG_OUT_WRAP=4096 timeout 300 python tools/gpu_probe.py --no-imad --circuits ${CIRCUIT:-circuit9_authV2} --batch $((148*T)) --reps 2 2>&1 | \
    python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print('T', $T, d['circuit'], 'B', d['B'], 'ms', d['ms'], 'wit/s', d['witness_per_s'])
" | tee -a $OUT/sweep.log
done
if [ -z "$NO_BENCH" ]; then
timeout 1500 python bench.py --steps 2 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -3 $OUT/bench.err
cat $OUT/bench.json
fi
