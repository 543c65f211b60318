#!/bin/bash
# round 2, GPU pass c: kernel variants (OP_DOT shapes, threads per CTA), bit kernel with deeper prefetch, synccheck, bench
O=gpurun_out/r02c; mkdir -p $O
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python tools/sanitize_probe.py > $O/sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?" | tee -a $O/sanitizer_synccheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_probe.py > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $O/sanitizer_racecheck.log
P="python tools/gpu_probe.py --no-imad --reps 3"
for B in 16384 65536; do timeout 300 $P --circuits circuit8_sha256_512 --batch $B >> $O/probe_sha.jsonl 2>> $O/probe.err; done
cut -c1-120 $O/probe_sha.jsonl
run() { tag=$1; shift; echo "== $tag" | tee -a $O/probe_authv2.jsonl; env "$@" timeout 300 $P --circuits circuit9_authV2 >> $O/probe_authv2.jsonl 2>> $O/probe.err; tail -1 $O/probe_authv2.jsonl | cut -c1-110; }
run shapes1_t512 GW_BATCH=75776
run shapes0_t512 GW_BATCH=75776 GW_DOT_SHAPES=0
run shapes1_t512_regs10 GW_BATCH=75776 GW_REGS=10
run shapes1_t512_regs13 GW_BATCH=75776 GW_REGS=13
run t640_regs10 GW_BATCH=94720 GW_REGS=10 GW_THREADS=640 GW_LIB_PATH=$PWD/circom-witnesscalc_b200/lib_variants/libcwc_t640.so
run t640_regs9 GW_BATCH=94720 GW_REGS=9 GW_THREADS=640 GW_LIB_PATH=$PWD/circom-witnesscalc_b200/lib_variants/libcwc_t640.so
run t768_regs8 GW_BATCH=98304 GW_REGS=8 GW_THREADS=768 GW_LIB_PATH=$PWD/circom-witnesscalc_b200/lib_variants/libcwc_t768.so
run t512_half GW_BATCH=37888
run t512_quarter GW_BATCH=18944
timeout 300 $P --circuits circuit7_poseidon4,circuit5_poseidon,circuit6_num2bits --batch 65536 >> $O/probe_small.jsonl 2>> $O/probe.err
timeout 900 python bench.py --steps 3 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02c/bench.json'))
print("value",d["value"],"frac",d["roofline"]["frac"],"e2e",d["e2e"]["value"], d["config"]["workload"])
for c in d["configs"]:
    print({k:c.get(k) for k in ("config","ms_per_step","witnesses_per_s","roofline","gpu_kernel_ms","cpu_port_1thread_ms","error") if c.get(k) is not None})
PY
