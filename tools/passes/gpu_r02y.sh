#!/bin/bash
# round 2, GPU pass y (last): what the driver runs at round end -- GPU suite, smoke, bench both arms
O=gpurun_out/r02y; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 900 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02y/bench.json'))
print("value",d["value"],"frac",d["roofline"]["frac"],"e2e",d["e2e"]["value"], "verified", d["extra"]["verified_against_oracle"], d["extra"]["verified_rows"], d["e2e"]["verified_against_oracle"], d["e2e"]["verified_rows"])
for c in d["configs"]:
    print({k:c.get(k) for k in ("config","ms_per_step","gpu_kernel_ms","cpu_port_1thread_ms","error","parity_rows_bit_exact","wtns_bit_exact") if c.get(k) is not None})
r=json.load(open('gpurun_out/r02y/bench_ref.json')); print("ref", r["value"], r["impl"], r["e2e"])
PY
