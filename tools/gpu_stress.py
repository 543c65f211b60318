#!/usr/bin/env python3
"""Randomised cross-check of every device path on one GPU: for many random graphs (all ops; Poseidon-shaped; Boolean;
field inputs taken apart into bits) the single-witness kernels (dataflow plan, level plan, bit plan) and the batch
kernels must agree with the Python oracle bit for bit.  Longer than the test suite would tolerate; run under gpurun."""
import importlib
import os
import random
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle as po  # noqa: E402
from tests import util  # noqa: E402
from tests.test_bitplan import boolean_graph, field_bits_graph  # noqa: E402

seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 1
budget_s = float(sys.argv[2]) if len(sys.argv) > 2 else 120.0
cwc = importlib.import_module("circom-witnesscalc_b200")
t_end = time.time() + budget_s
n_graphs = n_checks = 0
seed = seed0
while time.time() < t_end:
    rnd = random.Random(seed)
    kind = seed % 4
    if kind == 0:
        nodes, wit, imap = util.random_graph(rnd, n_ops=rnd.choice([100, 500, 2000]))
        n_in = 6
        rows = [[1] + [util.random_value(rnd) if rnd.random() < 0.8 else rnd.randrange(1 << 256) for _ in range(n_in)] for _ in range(40)]
    elif kind == 1:
        nodes, wit, imap = util.poseidon_like_graph(rnd, rnd.choice([2, 3, 5]), rnd.choice([3, 8, 30]))
        n_in = 6
        rows = [[1] + [util.random_value(rnd) for _ in range(n_in)] for _ in range(40)]
    elif kind == 2:
        n_in = rnd.choice([5, 24, 70])
        nodes, wit, imap = boolean_graph(rnd, n_inputs=n_in, n_gates=rnd.choice([60, 300, 1500]))
        rows = [[1] + [rnd.randrange(2) for _ in range(n_in)] for _ in range(70)]
        rows[rnd.randrange(70)][1 + rnd.randrange(n_in)] = rnd.randrange(po.M)
    else:
        nf, nb = rnd.choice([1, 3]), rnd.choice([0, 6])
        nodes, wit, imap = field_bits_graph(rnd, nf, nb)
        n_in = nf + nb
        rows = [[1] + [rnd.choice([0, po.M - 1, po.M, (1 << 256) - 1, rnd.randrange(1 << 256)]) for _ in range(nf)] + [rnd.randrange(2) for _ in range(nb)] for _ in range(45)]
    data = po.serialize_graph(nodes, wit, imap)
    want = [po.evaluate(nodes, r, wit, "circom") for r in rows]
    inp = np.frombuffer(b"".join(util.pack_u256(r) for r in rows), dtype=np.uint8).reshape(len(rows), n_in + 1, 32)
    for env in ({}, {"GW_LAT_MODE": "level"}, {"GW_BITSLICE": "0"}, {"GW_LAT_LINKS": "2"}):
        for k, v in env.items():
            os.environ[k] = v
        g = cwc.Graph(data)
        out = g.calc_witness_batch(inp)
        for b in range(len(rows)):
            assert util.unpack_u256(out[b].tobytes()) == want[b], ("batch", seed, env, b)
        for b in (0, len(rows) // 2, len(rows) - 1):
            lat, _ = g.calc_witness_latency(inp[b])
            assert util.unpack_u256(lat.tobytes()) == want[b], ("latency", seed, env, b)
            n_checks += 1
        for k in env:
            del os.environ[k]
    n_graphs += 1
    seed += 1
print(f"gpu_stress OK: {n_graphs} graphs (seeds {seed0}..{seed - 1}), {n_checks} single-witness checks, every batch row checked")
