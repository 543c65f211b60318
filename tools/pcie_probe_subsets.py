"""How the aggregate device->host rate of a box depends on HOW MANY of its GPUs copy at once: under torchrun with one rank
per GPU, the first k ranks copy 4 GiB each into pinned host memory at the same time (the others wait), k = 1 .. world.
Also which GPUs: the even ones / one per pair.  rank 0 prints one JSON line per experiment."""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("gloo")
    n = 4 << 30
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    h = torch.empty(n, dtype=torch.uint8).pin_memory()

    def timed(active, label, reps=3):
        mine = rank in active
        if mine:
            h.copy_(d, non_blocking=True); torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        if mine:
            for _ in range(reps):
                h.copy_(d, non_blocking=True)
            torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps if mine else 0.0
        t = torch.tensor([dt], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"probe": "pcie_subsets", "what": label, "gpus_copying": sorted(active), "slowest_GBps": round(n / t.item() / 1e9, 2),
                              "aggregate_GBps": round(len(active) * n / t.item() / 1e9, 2)}), flush=True)

    for k in range(1, world + 1):
        timed(set(range(k)), f"d2h, GPUs 0..{k - 1}")
    if world >= 8:
        timed({0, 2, 4, 6}, "d2h, the even GPUs")
        timed({0, 1, 4, 5}, "d2h, GPUs 0 1 4 5")
        timed({0, 4}, "d2h, GPUs 0 and 4")
        timed({0, 1}, "d2h, GPUs 0 and 1")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
