#!/bin/bash
# round 2, GPU pass s: where the interpreter's ~650 cycles per packet go (profiling build, timing experiments: results of dbg runs are not checked)
O=gpurun_out/r02s; mkdir -p $O
export GW_LIB_PATH=$PWD/circom-witnesscalc_b200/lib_variants/libcwc_prof.so
for D in 0 1 2 4 7; do echo "== dbg $D" >> $O/dbg.jsonl; GW_LAT_DBG=$D timeout 300 python tools/gpu_latency.py --reps 10 --circuits circuit9_authV2 >> $O/dbg.jsonl 2>> $O/probe.err; echo "dbg $D $(tail -1 $O/dbg.jsonl | cut -c30-95)"; done
