#!/bin/bash
# round 2, GPU pass h: full bench line (all configs), reference arm, single-set bit kernel latency
O=gpurun_out/r02h; mkdir -p $O
timeout 300 python tools/gpu_latency.py --reps 30 --circuits circuit8_sha256_512,circuit6_num2bits > $O/latency.jsonl 2>> $O/probe.err; cut -c1-200 $O/latency.jsonl
timeout 900 python bench.py --steps 3 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02h/bench.json'))
print("value",d["value"],"frac",d["roofline"]["frac"],"e2e",d["e2e"]["value"], d["config"]["workload"])
for c in d["configs"]:
    print({k:c.get(k) for k in ("config","ms_per_step","witnesses_per_s","roofline","gpu_kernel_ms","cpu_port_1thread_ms","error","bit_sliced") if c.get(k) is not None})
print(open('gpurun_out/r02h/bench_ref.json').read()[:600])
PY
