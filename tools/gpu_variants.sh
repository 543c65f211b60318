#!/bin/bash
# kernel-only timing of the config graphs under different launch/plan variants (one gpurun call)
# usage: [CIRCUITS=..] [PROBE_ARGS=..] bash tools/gpu_variants.sh <tag> "<ENV=VAL ...>" ...
OUT=gpurun_out/${1:-var}; shift
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
CIRCUITS=${CIRCUITS:-circuit7_poseidon4,circuit6_num2bits,circuit8_sha256_512,circuit9_authV2}
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg timeout 600 python tools/gpu_probe.py --no-imad --circuits $CIRCUITS ${PROBE_ARGS} 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['circuit'], 'B', d['B'], 'ms', d['ms'], 'wit/s', d['witness_per_s'])
" | tee -a $OUT/variants.log
done
