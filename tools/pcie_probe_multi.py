"""All GPUs of one box copying device->host (and host->device) at the same time: the aggregate ceiling of the
end-to-end number at N GPUs.  Run under torchrun; rank 0 prints one JSON line per direction."""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    n = 4 << 30
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for name, dst, src in (("d2h", h, d), ("h2d", d, h)):
        dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"probe": "pcie_all_gpus", "dir": name, "n_gpus": world, "GiB_per_gpu": 4,
                              "slowest_rank_GBps": round(n / t.item() / 1e9, 2), "aggregate_GBps": round(world * n / t.item() / 1e9, 2)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
