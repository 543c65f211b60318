#!/bin/bash
# round 2, GPU pass w2: how waiting warps should poll (sleep between polls of a progress counter)
O=gpurun_out/r02w; mkdir -p $O
run() { echo "== $*" >> $O/poll.jsonl; env "$@" timeout 300 python tools/gpu_latency.py --reps 10 --circuits circuit9_authV2 >> $O/poll.jsonl 2>> $O/probe.err; echo "$* $(tail -1 $O/poll.jsonl | cut -c50-95)"; }
for NS in 0 20 50 100 200 400; do run GW_LAT_POLL_NS=$NS; done
run GW_LAT_POLL_NS=50 GW_LAT_WARPS=7
run GW_LAT_POLL_NS=100 GW_LAT_WARPS=7
run GW_LAT_POLL_NS=100 GW_LAT_WARPS=8 GW_LAT_SLOW_WARPS=4
