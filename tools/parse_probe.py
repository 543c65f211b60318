#!/usr/bin/env python3
"""Throughput of the batch input path (gw_inputs_parse_batch: JSON Lines -> packed n_sets x I x 32 B, host threads only):
records per second for authV2-shaped records against the kernel's consumption rate.  No GPU needed."""
import importlib
import json
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import util  # noqa: E402

cwc = importlib.import_module("circom-witnesscalc_b200")
name = sys.argv[1] if len(sys.argv) > 1 else "circuit9_authV2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
g = cwc.Graph(util.golden_graph(name))
base = json.loads(util.golden_inputs(name))
rnd = random.Random(1)
M = cwc.M


def rand_like(v):
    if isinstance(v, list):
        return [rand_like(x) for x in v]
    return str(rnd.randrange(M))


recs = [json.dumps({k: rand_like(v) for k, v in base.items()}) for _ in range(64)]
text = ("\n".join(recs[i % 64] for i in range(n)) + "\n").encode()
for threads in (1, 4, 0):
    t0 = time.perf_counter()
    arr = g.parse_inputs_batch(text, threads)
    dt = time.perf_counter() - t0
    assert arr.shape[0] == n
    print(json.dumps({"probe": "gw_inputs_parse_batch", "circuit": name, "records": n, "bytes_per_record": len(text) // n, "threads": threads or os.cpu_count(),
                      "records_per_s": round(n / dt), "MB_per_s": round(len(text) / dt / 1e6, 1)}), flush=True)
