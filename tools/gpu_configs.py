#!/usr/bin/env python3
"""BASELINE.json configs 1-5 at their stated sizes on one GPU: kernel-only time (inputs and witnesses resident in HBM,
CUDA events, best of --reps), node-ops/s, algorithmic HBM GB/s, and the C restatement of the reference on the host
(bounded sample, all threads) beside it.  One JSON line per config.  (Config 4's multi-GPU numbers are bench.py's.)"""
import argparse
import importlib
import json
import os
import statistics
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cref, pyoracle as po  # noqa: E402
from tests import util  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--only", default="")
a = ap.parse_args()
cwc = importlib.import_module("circom-witnesscalc_b200")
dev = torch.device("cuda:0")
cores = os.cpu_count() or 1
free_b, total_b = torch.cuda.mem_get_info()


def inputs_for(name, g, B, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    n_u = min(B, 8192)
    vals = util.random_field_batch(rng, (n_u, g.n_inputs))
    if "sha256" in name:
        vals[:] = 0
        vals[:, :, 0] = rng.integers(0, 2, size=(n_u, g.n_inputs), dtype=np.uint64)
    for key in ("authClaimNonRevMtpNoAux", "gistMtpNoAux"):
        if key in g.input_signals:
            off, ln = g.input_signals[key]
            vals[:, off:off + ln, :] = 0
            vals[:, off:off + ln, 0] = rng.integers(0, 2, size=(n_u, ln), dtype=np.uint64)
    vals[:, 0, :] = 0
    vals[:, 0, 0] = 1
    return vals.view(np.uint8).reshape(n_u, g.n_inputs, 32)


def batch_config(tag, name, B, seed):
    data = util.golden_graph(name)
    g = cwc.Graph(data)
    I, W = g.n_inputs, g.n_witness
    host = inputs_for(name, g, B, seed)
    d_u = torch.from_numpy(host.reshape(host.shape[0], -1)).to(dev)
    d_in = d_u.repeat((B + host.shape[0] - 1) // host.shape[0], 1)[:B].contiguous()
    # chunk so that a chunk's witnesses fit HBM (whole quads of warps per SM)
    max_chunk = int((free_b - (8 << 30)) // (W * 32))
    chunk = B if B <= max_chunk else max(148 * 32, max_chunk // (148 * 128) * (148 * 128))
    d_out = torch.empty((chunk, W * 32), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    chunks = [(lo, min(lo + chunk, B)) for lo in range(0, B, chunk)]

    def step():
        for lo, hi in chunks:
            g.calc_witness_batch_device(0, d_in[lo:hi].data_ptr(), hi - lo, d_out.data_ptr(), None, stream)
    step()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    # parity of the last chunk's first rows against the C oracle
    cg = cref.CGraph(data)
    lo, hi = chunks[-1]
    k = min(8, hi - lo)
    ok = bool((cg.evaluate_batch(d_in[lo:lo + k].cpu().numpy().reshape(k, I, 32), min(k, cores)) ==
               d_out[:k].cpu().numpy().reshape(k, W, 32)).all())
    # CPU restatement on a bounded sample
    n_cpu = min(B, max(cores, int(4.0 / max(1e-5, _cpu_one(cg, host[:1]))) // cores * cores))
    reps_in = np.ascontiguousarray(np.resize(host, (n_cpu, I, 32)))
    t0 = time.perf_counter(); cg.evaluate_batch(reps_in, cores); dt = time.perf_counter() - t0
    info = g.info
    print(json.dumps({"config": tag, "circuit": name, "B": B, "launches": len(chunks), "sets_per_launch": chunk, "kernel_ms": round(best, 3),
                      "witness_per_s": round(B / best * 1e3, 1), "node_ops_per_s": round(B * info["n_ops"] / best * 1e3, 1),
                      "field_mul_per_s": round(B * info["n_mul"] / best * 1e3, 1),
                      "hbm_algorithmic_GBps": round(B * 32 * (I - 1 + W) / best / 1e6, 1), "bit_exact_sample": ok,
                      "cpu_port_witness_per_s": round(n_cpu / dt, 1), "cpu_cores": cores, "cpu_sample_sets": n_cpu,
                      "speedup_kernel_vs_cpu": round(B / best * 1e3 / (n_cpu / dt), 1),
                      "n_ops": info["n_ops"], "n_mul": info["n_mul"], "n_div": info["n_div"], "I": I, "W": W, "n_instrs": info["n_instrs"],
                      "n_spill": info["n_spill"]}), flush=True)
    del d_in, d_out, d_u
    torch.cuda.empty_cache()


def _cpu_one(cg, row):
    t0 = time.perf_counter(); cg.evaluate_batch(row, 1); return time.perf_counter() - t0


def single_config(tag, name, reps=50):
    """one witness: JSON + graph -> .wtns through gw_calc_witness semantics (pre-loaded graph), latency mode"""
    data = util.golden_graph(name)
    nodes, wit, imap = po.deserialize_graph(data)
    g = cwc.Graph(data)
    buf = po.build_input_buffer(nodes, imap, po.deserialize_inputs(util.golden_inputs(name)))
    row = np.frombuffer(util.pack_u256(buf), dtype=np.uint8).reshape(g.n_inputs, 32)
    out, _ = g.calc_witness_latency(row)
    ok = po.wtns_from_witness(util.unpack_u256(out.tobytes())) == util.golden_wtns(name)
    ok = ok and g.calc_witness_wtns(util.golden_inputs(name)) == util.golden_wtns(name)
    kern, wall, e2e = [], [], []
    for _ in range(reps):
        t0 = time.perf_counter(); _, ms = g.calc_witness_latency(row); wall.append((time.perf_counter() - t0) * 1e3); kern.append(ms)
        t0 = time.perf_counter(); g.calc_witness_wtns(util.golden_inputs(name)); e2e.append((time.perf_counter() - t0) * 1e3)
    cg = cref.CGraph(data)
    cpu = []
    for _ in range(30):
        t0 = time.perf_counter(); cg.evaluate_batch(row.reshape(1, g.n_inputs, 32), 1); cpu.append((time.perf_counter() - t0) * 1e3)
    print(json.dumps({"config": tag, "circuit": name, "wtns_bit_exact": bool(ok), "gpu_kernel_ms_median": round(statistics.median(kern), 3),
                      "gpu_call_ms_median": round(statistics.median(wall), 3), "gpu_json_to_wtns_ms_median": round(statistics.median(e2e), 3),
                      "cpu_port_1thread_ms_median": round(statistics.median(cpu), 3), "reps": reps}), flush=True)


todo = [("1", lambda: single_config("1: circuit5_poseidon single witness", "circuit5_poseidon")),
        ("2", lambda: batch_config("2a: circuit6_num2bits x 65536", "circuit6_num2bits", 65536, 6)),
        ("2", lambda: batch_config("2b: circuit7_poseidon4 x 65536", "circuit7_poseidon4", 65536, 7)),
        ("3", lambda: batch_config("3: circuit8_sha256_512 x 16384", "circuit8_sha256_512", 16384, 8)),
        ("4", lambda: batch_config("4: circuit9_authV2 x 262144 (1 GPU)", "circuit9_authV2", 262144, 9)),
        ("5", lambda: single_config("5: circuit9_authV2 single witness, latency mode", "circuit9_authV2"))]
for k, fn in todo:
    if not a.only or k in a.only.split(","):
        fn()
