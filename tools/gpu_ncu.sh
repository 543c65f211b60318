#!/bin/bash
# usage: bash tools/gpu_ncu.sh <tag> <circuit> <batch> [ENV=VAL ...]   -> gpurun_out/<tag>/prof_<circuit>_<name>.ncu-rep
TAG=$1; CIRCUIT=$2; BATCH=$3; shift 3
OUT=gpurun_out/$TAG; mkdir -p $OUT
NAME=$(echo "$*" | tr ' =' '__')
env "$@" timeout 900 ncu --set full --clock-control none --import-source on -k regex:eval_batch -s 1 -c 1 -o $OUT/prof_${CIRCUIT}_${NAME} \
  python tools/gpu_probe.py --circuits $CIRCUIT --batch $BATCH --reps 1 --no-imad > $OUT/ncu_${CIRCUIT}_${NAME}.log 2>&1
echo "ncu rc=$? $NAME"; tail -2 $OUT/ncu_${CIRCUIT}_${NAME}.log | cut -c1-200
