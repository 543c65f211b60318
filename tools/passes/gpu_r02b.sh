#!/bin/bash
# round 2, GPU pass b: tests with the bit-sliced path, sanitizers after the flag/convergence fix, bench, ncu captures
O=gpurun_out/r02b; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_probe.py --big > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $O/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_probe.py > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $O/sanitizer_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python tools/sanitize_probe.py > $O/sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?" | tee -a $O/sanitizer_synccheck.log
timeout 900 python bench.py --steps 3 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02b/bench.json'))
print("value",d["value"],"frac",d["roofline"]["frac"],"e2e",d["e2e"]["value"])
for c in d["configs"]:
    print({k:c.get(k) for k in ("config","ms_per_step","witnesses_per_s","roofline","parity_rows_bit_exact","gpu_kernel_ms","cpu_port_1thread_ms","error")})
PY
# probes: SHA-256 at 16384 / 65536 / 262144 sets, bit path on and off
for B in 16384 65536; do timeout 300 python tools/gpu_probe.py --circuits circuit8_sha256_512 --batch $B --reps 5 --no-imad >> $O/probe_sha.jsonl 2>> $O/probe.err; done
GW_BITSLICE=0 timeout 300 python tools/gpu_probe.py --circuits circuit8_sha256_512 --batch 16384 --reps 3 --no-imad >> $O/probe_sha_generic.jsonl 2>> $O/probe.err
cut -c1-160 $O/probe_sha.jsonl
# launch list + full captures of the bit kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/launches_sha.csv python tools/gpu_probe.py --circuits circuit8_sha256_512 --batch 16384 --reps 2 --no-imad > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bit_ -s 3 -c 3 -o $O/prof_sha_bit python tools/gpu_probe.py --circuits circuit8_sha256_512 --batch 16384 --reps 2 --no-imad > $O/ncu_sha.log 2>&1; echo "ncu sha rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:eval_batch -s 1 -c 1 -o $O/prof_authv2 python tools/gpu_probe.py --circuits circuit9_authV2 --batch 75776 --reps 1 --no-imad > $O/ncu_authv2.log 2>&1; echo "ncu authv2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:eval_batch -s 1 -c 1 -o $O/prof_num2bits python tools/gpu_probe.py --circuits circuit6_num2bits --batch 65536 --reps 1 --no-imad > $O/ncu_num2bits.log 2>&1; echo "ncu num2bits rc=$?"
ls -la $O
