#!/usr/bin/env python3
"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): both kernels on small graphs,
results checked against the oracle.  Run as `compute-sanitizer --tool <t> python tools/sanitize_probe.py [--big]`."""
import importlib
import os
import random
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle as po  # noqa: E402
from tests import util  # noqa: E402

cwc = importlib.import_module("circom-witnesscalc_b200")
big = "--big" in sys.argv
rnd = random.Random(3)
n_ok = 0
# throughput kernel: golden graphs with spills, Div batching, narrow values; ragged batch
for name, B in (("circuit5_poseidon", 300), ("circuit2", 97), ("circuit6_num2bits", 200), ("poseidon2", 64)):
    data = util.golden_graph(name)
    nodes, wit, imap = po.deserialize_graph(data)
    g = cwc.Graph(data)
    rng = np.random.default_rng(1)
    vals = util.random_field_batch(rng, (B, g.n_inputs))
    vals[:, 0, :] = 0
    vals[:, 0, 0] = 1
    out = g.calc_witness_batch(vals.view(np.uint8).reshape(B, g.n_inputs, 32))
    for b in (0, B // 2, B - 1):
        assert util.unpack_u256(out[b].tobytes()) == po.evaluate(nodes, util.limbs_to_ints(vals[b]), wit), (name, b)
    n_ok += 1
    # latency kernel on the reference's fixture
    assert cwc.calc_witness_wtns(util.golden_inputs(name), data) == util.golden_wtns(name), name
    n_ok += 1
# random graphs over all ops through both kernels (tiny register file -> spills)
os.environ["GW_REGS"] = "5"
for t in range(3):
    nodes, wit, imap = util.random_graph(rnd, n_ops=300)
    g = cwc.Graph(po.serialize_graph(nodes, wit, imap))
    rows = [[1] + [util.random_value(rnd) for _ in range(6)] for _ in range(40)]
    inp = np.frombuffer(b"".join(util.pack_u256(r) for r in rows), dtype=np.uint8).reshape(40, 7, 32)
    out = g.calc_witness_batch(inp)
    lat, _ = g.calc_witness_latency(inp[3])
    assert util.unpack_u256(out[3].tobytes()) == po.evaluate(nodes, rows[3], wit, "circom") == util.unpack_u256(lat.tobytes())
    n_ok += 1
# bit-sliced path (Boolean graph, ragged batch, two input sets that break the bit contract -> generic fallback)
os.environ.pop("GW_REGS", None)
from tests.test_bitplan import boolean_graph  # noqa: E402
nodes, wit, imap = boolean_graph(rnd, n_inputs=24, n_gates=200)
g = cwc.Graph(po.serialize_graph(nodes, wit, imap))
rows = [[1] + [rnd.randrange(2) for _ in range(24)] for _ in range(70)]
rows[5][3] = 7
rows[69][24] = po.M - 1
inp = np.frombuffer(b"".join(util.pack_u256(r) for r in rows), dtype=np.uint8).reshape(70, 25, 32)
out = g.calc_witness_batch(inp)
for b in (0, 5, 33, 69):
    assert util.unpack_u256(out[b].tobytes()) == po.evaluate(nodes, rows[b], wit, "circom"), b
n_ok += 1
# the same Boolean graph, ONE input set: single-set bit kernel (TMA-fed header ring, mbarriers), then a set that breaks the
# contract (generic latency kernel in the same call)
for row in (rows[0], rows[5]):
    lat, _ = g.calc_witness_latency(np.frombuffer(util.pack_u256(row), dtype=np.uint8).reshape(25, 32))
    assert util.unpack_u256(lat.tobytes()) == po.evaluate(nodes, row, wit, "circom")
n_ok += 1
# field inputs taken apart into bits, wide witness values (bit_expand_wide_kernel)
from tests.test_bitplan import field_bits_graph  # noqa: E402
nodes, wit, imap = field_bits_graph(rnd, 3, 6)
g = cwc.Graph(po.serialize_graph(nodes, wit, imap))
rows = [[1] + [rnd.choice([0, po.M - 1, po.M + 5, (1 << 256) - 1, rnd.randrange(1 << 256)]) for _ in range(3)] + [rnd.randrange(2) for _ in range(6)] for _ in range(45)]
rows[7][4] = 9
inp = np.frombuffer(b"".join(util.pack_u256(r) for r in rows), dtype=np.uint8).reshape(45, 10, 32)
out = g.calc_witness_batch(inp)
for b in (0, 7, 31, 32, 44):
    assert util.unpack_u256(out[b].tobytes()) == po.evaluate(nodes, rows[b], wit, "circom"), b
n_ok += 1
# streaming path
g = cwc.Graph(util.golden_graph("circuit5_poseidon"))
got = []
inp = np.zeros((70, g.n_inputs, 32), dtype=np.uint8)
inp[:, :, 0] = 1
g.calc_witness_batch_stream(inp.ctypes.data, 70, lambda d, f, r, fl: got.append((f, r.shape[0])) and 0, chunk_sets=32)
assert got == [(0, 32), (32, 32), (64, 6)], got
n_ok += 1
if big:
    for name in ("circuit8_sha256_512", "circuit9_authV2"):
        assert cwc.calc_witness_wtns(util.golden_inputs(name), util.golden_graph(name)) == util.golden_wtns(name), name
        n_ok += 1
print("sanitize_probe OK:", n_ok, "checks")
