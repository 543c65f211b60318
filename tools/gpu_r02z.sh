#!/bin/bash
# round 2, GPU pass z2: 17 warps per SM (launch bounds 544, 96 registers) against 16
O=gpurun_out/r02z; mkdir -p $O
P="python tools/gpu_probe.py --no-imad --reps 3 --circuits circuit9_authV2"
run() { echo "== $*" >> $O/probe_t544.jsonl; env "$@" timeout 300 $P >> $O/probe_t544.jsonl 2>> $O/probe.err; echo "$* $(tail -1 $O/probe_t544.jsonl | cut -c30-110)"; }
run GW_BATCH=75776
run GW_BATCH=80512 GW_THREADS=544 GW_LIB_PATH=$PWD/circom-witnesscalc_b200/lib_variants/libcwc_t544.so
run GW_BATCH=75776 GW_THREADS=512 GW_LIB_PATH=$PWD/circom-witnesscalc_b200/lib_variants/libcwc_t544.so
run GW_BATCH=80512 GW_THREADS=544 GW_REGS=11 GW_LIB_PATH=$PWD/circom-witnesscalc_b200/lib_variants/libcwc_t544.so
