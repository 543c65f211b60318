#!/bin/bash
# round 2, GPU pass d (2 GPUs): bench under torchrun exactly as the driver launches it (strong scaling, gloo), both arms
O=gpurun_out/r02d; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1; nproc >> $O/topo.txt; free -g >> $O/topo.txt
for f in /sys/bus/pci/devices/*/numa_node; do d=$(dirname $f); if [ -e $d/class ] && grep -q "^0x0302" $d/class; then echo "$d numa $(cat $f)" >> $O/topo.txt; fi; done
ls /sys/devices/system/node/ >> $O/topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "bench2 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/bench_ref_2gpu.json 2>> $O/bench_2gpu.err; echo "ref2 rc=$?"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "stream or multi_gpu" > $O/pytest_2gpu.log 2>&1; tail -2 $O/pytest_2gpu.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02d/bench_2gpu.json'))
print({k:d[k] for k in ("value","n_gpus","scaling","ms_per_step")}, d["config"]["workload"])
print("e2e", d["e2e"])
PY
tail -5 $O/bench_2gpu.err
