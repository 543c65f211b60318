#!/usr/bin/env python3
"""BASELINE config 5: single-witness latency of circuit9_authV2 (and others) on one GPU, latency mode,
against the C restatement of the reference on one host thread.  Median of --reps runs after warm-up."""
import argparse
import importlib
import json
import os
import statistics
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cref, pyoracle as po  # noqa: E402
from tests import util  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--circuits", default="circuit5_poseidon,circuit6_num2bits,circuit8_sha256_512,circuit9_authV2")
ap.add_argument("--reps", type=int, default=100)
a = ap.parse_args()
cwc = importlib.import_module("circom-witnesscalc_b200")
for name in a.circuits.split(","):
    data = util.golden_graph(name)
    nodes, wit, imap = po.deserialize_graph(data)
    g = cwc.Graph(data)
    buf = po.build_input_buffer(nodes, imap, po.deserialize_inputs(util.golden_inputs(name)))
    row = np.frombuffer(util.pack_u256(buf), dtype=np.uint8).reshape(g.n_inputs, 32)
    out, _ = g.calc_witness_latency(row)
    ok = po.wtns_from_witness(util.unpack_u256(out.tobytes())) == util.golden_wtns(name)
    kern, wall = [], []
    for _ in range(a.reps):
        t0 = time.perf_counter()
        _, ms = g.calc_witness_latency(row)
        wall.append((time.perf_counter() - t0) * 1e3)
        kern.append(ms)
    cg = cref.CGraph(data)
    cpu = []
    for _ in range(max(5, min(a.reps, 30))):
        t0 = time.perf_counter()
        cg.evaluate_batch(row.reshape(1, g.n_inputs, 32), 1)
        cpu.append((time.perf_counter() - t0) * 1e3)
    js = json.dumps(util.golden_inputs(name))
    t0 = time.perf_counter()
    w = g.calc_witness_wtns(util.golden_inputs(name))
    t_json = (time.perf_counter() - t0) * 1e3
    print(json.dumps({"circuit": name, "bit_exact": ok, "gpu_kernel_ms_median": round(statistics.median(kern), 3),
                      "gpu_call_ms_median": round(statistics.median(wall), 3), "gpu_json_to_wtns_ms": round(t_json, 3),
                      "cpu_1thread_ms_median": round(statistics.median(cpu), 3), "reps": a.reps,
                      "levels": len(g.info) and None}), flush=True)
