"""All GPUs of one box copying device->host (and host->device) at the same time: the aggregate ceiling of the
end-to-end number at N GPUs, and what does (not) move it: the way the host buffer is pinned (cudaHostAlloc vs
cudaHostRegister of transparent-huge-page memory), the copy granularity, two copies in flight per GPU.
Run under torchrun (gloo is used for the barrier only); rank 0 prints one JSON line per experiment."""
import ctypes
import json
import mmap
import os
import time

import torch
import torch.distributed as dist


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("gloo")
    n = 4 << 30
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    h_alloc = torch.empty(n, dtype=torch.uint8).pin_memory()                       # cudaHostAlloc
    # cudaHostRegister of anonymous memory advised to use transparent huge pages, touched by this rank first
    mm = mmap.mmap(-1, n, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    try:
        mm.madvise(mmap.MADV_HUGEPAGE)
    except (AttributeError, OSError):
        pass
    h_reg = torch.frombuffer(mm, dtype=torch.uint8)
    h_reg.fill_(1)
    rc = torch.cuda.cudart().cudaHostRegister(h_reg.data_ptr(), n, 0)
    registered = int(rc) == 0
    s2 = torch.cuda.Stream()

    def timed(fn, label, reps=3):
        fn(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        t = torch.tensor([dt], dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"probe": "pcie_all_gpus", "what": label, "n_gpus": world, "GiB_per_gpu": n >> 30,
                              "slowest_rank_GBps": round(n / t.item() / 1e9, 2), "aggregate_GBps": round(world * n / t.item() / 1e9, 2)}), flush=True)

    def chunks(dst, src, k):
        step = n // k
        for i in range(k):
            dst[i * step:(i + 1) * step].copy_(src[i * step:(i + 1) * step], non_blocking=True)

    def two_streams(dst, src):
        half = n // 2
        dst[:half].copy_(src[:half], non_blocking=True)
        with torch.cuda.stream(s2):
            dst[half:].copy_(src[half:], non_blocking=True)

    timed(lambda: h_alloc.copy_(d, non_blocking=True), "d2h cudaHostAlloc, one 4 GiB copy")
    timed(lambda: d.copy_(h_alloc, non_blocking=True), "h2d cudaHostAlloc, one 4 GiB copy")
    timed(lambda: chunks(h_alloc, d, 64), "d2h cudaHostAlloc, 64 x 64 MiB copies")
    timed(lambda: two_streams(h_alloc, d), "d2h cudaHostAlloc, two streams x 2 GiB")
    if registered:
        timed(lambda: h_reg.copy_(d, non_blocking=True), "d2h cudaHostRegister (THP-advised anonymous memory), one 4 GiB copy")
        timed(lambda: d.copy_(h_reg, non_blocking=True), "h2d cudaHostRegister (THP-advised anonymous memory), one 4 GiB copy")
        torch.cuda.cudart().cudaHostUnregister(h_reg.data_ptr())
    elif rank == 0:
        print(json.dumps({"probe": "pcie_all_gpus", "what": "cudaHostRegister failed", "rc": int(rc)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
