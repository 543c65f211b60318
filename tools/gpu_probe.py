#!/usr/bin/env python3
"""Quick on-GPU probe: integer-pipe microbenchmark + kernel-only timing of the batch evaluator on the
golden graphs (inputs/outputs resident in HBM).  Prints one JSON line per measurement."""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import util  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--circuits", default="circuit5_poseidon,circuit7_poseidon4,circuit6_num2bits,circuit8_sha256_512,circuit9_authV2")
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--no-imad", action="store_true")
a = ap.parse_args()

cwc = importlib.import_module("circom-witnesscalc_b200")
dev = torch.device("cuda:0")
if not a.no_imad:
    for which, nm in enumerate(["imad.wide.u32 carry rows", "mad.lo.u32", "mad.hi.u32", "add.u32"]):
        r = cwc.microbench_imad(0, which)
        print(json.dumps({"microbench": nm, "Gops_per_s": round(r / 1e9, 1)}), flush=True)

for name in a.circuits.split(","):
    g = cwc.Graph(util.golden_graph(name))
    I, W = g.n_inputs, g.n_witness
    B = int(os.environ.get("GW_BATCH", "0")) or a.batch or max(1024, min(148 * 256 * 2, int(110e9 // (32 * W))))
    if a.batch and not os.environ.get("GW_DEBUG_OUT_WRAP"):
        B = min(B, int(160e9 // (32 * W)))
    rng = np.random.default_rng(9)
    vals = util.random_field_batch(rng, (min(B, 4096), I))
    if "sha256" in name:
        vals[:] = 0
        vals[:, :, 0] = rng.integers(0, 2, size=vals.shape[:2], dtype=np.uint64)
    vals[:, 0, :] = 0
    vals[:, 0, 0] = 1
    host = torch.from_numpy(vals.view(np.uint8).reshape(-1, I * 32))
    reps = (B + host.shape[0] - 1) // host.shape[0]
    d_in = host.repeat(reps, 1)[:B].contiguous().to(dev)
    wrap = int(os.environ.get("GW_DEBUG_OUT_WRAP", "0"))     # profiling aid: all witnesses share `wrap` output rows
    d_out = torch.empty((wrap or B, W * 32), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    g.calc_witness_batch_device(0, d_in.data_ptr(), B, d_out.data_ptr(), None, stream)   # warm-up
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.calc_witness_batch_device(0, d_in.data_ptr(), B, d_out.data_ptr(), None, stream)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    info = g.info
    print(json.dumps({"circuit": name, "B": B, "ms": round(best, 3), "witness_per_s": round(B / best * 1e3, 1),
                      "node_ops_per_s": round(B * info["n_ops"] / best * 1e3, 1),
                      "mul_per_s": round(B * info["n_mul"] / best * 1e3, 1),
                      "out_GBps": round(B * W * 32 / best / 1e6, 1), "info": info}), flush=True)
    del d_in, d_out
    torch.cuda.empty_cache()
