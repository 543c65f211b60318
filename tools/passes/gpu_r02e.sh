#!/bin/bash
# round 2, GPU pass e (8 GPUs): bench under torchrun as the driver launches it (strong scaling), PCIe ceiling probes
O=gpurun_out/r02e; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1; nproc >> $O/topo.txt; free -g >> $O/topo.txt; ls /sys/devices/system/node/ >> $O/topo.txt 2>&1
for f in /sys/bus/pci/devices/*/numa_node; do d=$(dirname $f); if [ -e $d/class ] && grep -q "^0x0302" $d/class; then echo "$d numa $(cat $f)" >> $O/topo.txt; fi; done
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 3 --warmup 3 > $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.err; echo "bench$N rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 tools/pcie_probe_multi.py > $O/pcie_${N}gpu.jsonl 2> $O/pcie.err; echo "pcie rc=$?"
timeout 120 python tools/pcie_probe_multi.py > $O/pcie_1gpu.jsonl 2>> $O/pcie.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu > $O/bench_4gpu.json 2> $O/bench_4gpu.err; echo "bench4 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02e/bench_*gpu.json')):
    d=json.load(open(f))
    print(f, {k:d[k] for k in ("value","n_gpus","scaling","ms_per_step")}, "e2e", round(d["e2e"]["value"]), round(d["e2e"]["d2h_GBps"],1), d["e2e"].get("in_library_n_gpus",{}).get("value"))
PY
cat $O/pcie_*gpu.jsonl | cut -c1-220
