#!/bin/bash
# round 2, GPU pass u (final, 1 GPU): bench line + reference arm as the driver runs them, launch list of the same command
O=gpurun_out/r02u; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/gpu.txt; nproc > $O/nproc.txt
timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 3 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > $O/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02u/bench.json'))
print("value",d["value"],"frac",d["roofline"]["frac"],"e2e",d["e2e"]["value"], "launches", d["gpu_launches"], d["clocks"])
for c in d["configs"]:
    print({k:c.get(k) for k in ("config","ms_per_step","witnesses_per_s","roofline","gpu_kernel_ms","gpu_call_ms","cpu_port_1thread_ms","error","bit_sliced","parity_rows_bit_exact","wtns_bit_exact") if c.get(k) is not None})
r=json.load(open('gpurun_out/r02u/bench_ref.json')); print("ref", r["value"], r["cpu_baseline"])
PY
