#!/bin/bash
# round 2, GPU pass z3 (8 GPUs): aggregate device->host rate against the number of GPUs copying at once
O=gpurun_out/r02z; mkdir -p $O
N=$(nvidia-smi -L | wc -l)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/pcie_probe_subsets.py > $O/pcie_subsets_${N}gpu.jsonl 2> $O/pcie_subsets.err; echo "rc=$?"; cut -c1-200 $O/pcie_subsets_${N}gpu.jsonl
