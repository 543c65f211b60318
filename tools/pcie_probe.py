"""Host<->device copy ceiling of the box (pinned memory, one stream): the bound of the end-to-end number, whose
device->host copy of the full witness (32*W bytes per input set) dominates.  One JSON line per direction and size."""
import json
import sys
import torch


def main():
    assert torch.cuda.is_available(), "needs a GPU"
    for mb in (256, 4096):
        n = mb << 20
        h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        d = torch.empty(n, dtype=torch.uint8, device="cuda")
        for name, dst, src in (("h2d", d, h), ("d2h", h, d)):
            best = 1e30
            for rep in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                dst.copy_(src, non_blocking=True)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            print(json.dumps({"probe": "pcie", "dir": name, "MiB": mb, "GBps": round(n / (best * 1e-3) / 1e9, 2)}))
        del h, d
    return 0


if __name__ == "__main__":
    sys.exit(main())
