#!/bin/bash
# round 2, GPU pass v (final, N GPUs): bench + reference arm under torchrun as the driver launches them (strong scaling)
O=gpurun_out/r02v; mkdir -p $O
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $O/bench_ref_${N}gpu.json 2> $O/bench_ref_${N}gpu.err; echo "ref$N rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 3 --warmup 3 > $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.err; echo "bench$N rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r02v/bench_${N}gpu.json'))
print({k:d[k] for k in ("value","n_gpus","scaling","ms_per_step")}, "e2e", round(d["e2e"]["value"]), round(d["e2e"]["d2h_GBps"],1), "frac", d["roofline"]["frac"])
r=json.load(open('gpurun_out/r02v/bench_ref_${N}gpu.json')); print("ref", r["value"], r["cpu_baseline"]["cores"])
PY
