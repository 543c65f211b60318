#!/usr/bin/env python3
"""Latency-mode variants of one circuit on one GPU: the plan options are read from the environment when the
latency plan of a Graph is first built, so every variant loads the graph anew.  One JSON line per variant."""
import argparse
import importlib
import json
import os
import statistics
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle as po  # noqa: E402
from tests import util  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--circuit", default="circuit9_authV2")
ap.add_argument("--reps", type=int, default=15)
ap.add_argument("--variants", default="")
ap.add_argument("--clocks", default="")      # file for the per-level cycle counts of the first variant
a = ap.parse_args()
cwc = importlib.import_module("circom-witnesscalc_b200")
data = util.golden_graph(a.circuit)
nodes, wit, imap = po.deserialize_graph(data)
buf = po.build_input_buffer(nodes, imap, po.deserialize_inputs(util.golden_inputs(a.circuit)))
row = np.frombuffer(util.pack_u256(buf), dtype=np.uint8).reshape(-1, 32)
want = util.golden_wtns(a.circuit)
variants = [dict(kv.split("=") for kv in v.split(",") if kv) for v in a.variants.split(";")] if a.variants else [{}]
KEYS = ["GW_LAT_WARPS", "GW_LAT_SLOW_WARPS", "GW_LAT_D", "GW_LAT_SPLIT", "GW_LAT_FUSE", "GW_LAT_CLOCKS", "GW_LAT_CLOCKS_FILE", "GW_LAT_DBG", "GW_LAT_GRID"]
for vi, v in enumerate(variants):
    for k in KEYS:
        os.environ.pop(k, None)
    for k, x in v.items():
        os.environ["GW_LAT_" + k.upper()] = x
    if vi == 0 and a.clocks:
        os.environ["GW_LAT_CLOCKS"] = "1"
        os.environ["GW_LAT_CLOCKS_FILE"] = a.clocks
    g = cwc.Graph(data)
    out, _ = g.calc_witness_latency(row)
    ok = po.wtns_from_witness(util.unpack_u256(out.tobytes())) == want
    os.environ.pop("GW_LAT_CLOCKS", None)
    kern = []
    for _ in range(a.reps):
        _, ms = g.calc_witness_latency(row)
        kern.append(ms)
    print(json.dumps({"circuit": a.circuit, "variant": v, "bit_exact": ok, "kernel_ms_median": round(statistics.median(kern), 3),
                      "kernel_ms_min": round(min(kern), 3)}), flush=True)
    del g
