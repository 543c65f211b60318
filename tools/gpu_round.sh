#!/bin/bash
# One gpurun call: GPU parity tests, kernel probe, headline bench, ncu launch list + one full capture.
# Usage (from the repo root, on the GPU box): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/smi.csv 2>&1
nproc > $OUT/nproc.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 600 python tools/gpu_probe.py > $OUT/probe.jsonl 2> $OUT/probe.err; echo "probe rc=$?"
cat $OUT/probe.jsonl | cut -c1-400
timeout 1500 python bench.py --steps 2 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
cat $OUT/bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"
cat $OUT/bench_ref.json
# launch list of the bench command (reduced steps; numbers printed under ncu are not bench values)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $OUT/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
# full capture of the batch kernel at the bench's occupancy (56 832 input sets = 384 threads per SM).  ncu replays the
# launch ~40 times and would have to save/restore the 146 GB of witness output each time, so for THIS capture only the
# witness rows wrap modulo 4096 rows (GW_DEBUG_OUT_WRAP: same instruction stream, same stores, smaller footprint).
GW_DEBUG_OUT_WRAP=4096 timeout 900 ncu --set full --clock-control none --import-source on -k regex:eval_batch -s 1 -c 1 -o $OUT/prof_authv2 \
  python tools/gpu_probe.py --circuits circuit9_authV2 --batch 56832 --reps 1 --no-imad > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:eval_batch -s 1 -c 1 -o $OUT/prof_poseidon4 \
  python tools/gpu_probe.py --circuits circuit7_poseidon4 --batch 75776 --reps 1 --no-imad > $OUT/ncu_full_p4.log 2>&1; echo "ncu full p4 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:eval_batch -s 1 -c 1 -o $OUT/prof_sha256 \
  python tools/gpu_probe.py --circuits circuit8_sha256_512 --batch 16384 --reps 1 --no-imad > $OUT/ncu_full_sha.log 2>&1; echo "ncu full sha rc=$?"
timeout 300 python tools/pcie_probe.py > $OUT/pcie.jsonl 2>&1; cat $OUT/pcie.jsonl
timeout 600 python tools/gpu_latency.py --reps 50 > $OUT/latency.jsonl 2> $OUT/latency.err; echo "latency rc=$?"; cat $OUT/latency.jsonl
ls -la $OUT
