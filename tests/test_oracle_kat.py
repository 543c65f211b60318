"""Pins the oracles (oracle/pyoracle.py, oracle/ref_eval.c) against every vector the reference's
own tests hold for the hot path and against circuit-level known answers."""
import hashlib
import json
import random

import numpy as np
import pytest

from oracle import cref
from tests import util
from tests.util import M, po

D = po.DUO


def both(op, a, b):
    r = po.eval_duo(op, a % M, b % M)
    assert cref.op_duo(op, a % M, b % M) == r
    return r


def test_reference_unit_vectors_graph_rs():
    # src/graph.rs:780-785 test_ok: shl(4, 2) == 16
    assert both(D["Shl"], 4, 2) == 16
    # src/graph.rs:788-800 test_div
    assert both(D["Div"], 2, 3) == 7296080957279758407415468581752425029516121466805344781232734728858602831873
    assert both(D["Div"], 6, 2) == 3
    assert both(D["Div"], 7, 2) == 10944121435919637611123202872628637544274182200208017171849102093287904247812
    # src/graph.rs:803-815 test_idiv
    assert both(D["Idiv"], 2, 3) == 0 and both(D["Idiv"], 6, 2) == 3 and both(D["Idiv"], 7, 2) == 3
    # src/graph.rs:818-826 test_fr_mod
    assert both(D["Mod"], 7, 2) == 1 and both(D["Mod"], 7, 9) == 7
    # src/graph.rs:850-883 test_u_gte
    assert both(D["Geq"], 10, 3) == 1 and both(D["Geq"], 3, 3) == 1 and both(D["Geq"], 2, 3) == 0
    assert both(D["Geq"], M - 1, 3) == 0          # -1 >= 3
    assert both(D["Geq"], M - 1, M - 2) == 1      # -1 >= -2
    assert both(D["Geq"], M - 2, M - 1) == 0
    assert both(D["Geq"], M - 2, M - 2) == 1


def test_documented_edge_semantics():
    assert both(D["Div"], 5, 0) == 0 and both(D["Idiv"], 5, 0) == 0 and both(D["Mod"], 5, 0) == 0   # graph.rs:109-121
    assert both(D["Shr"], 12345, 0) == 12345 and both(D["Shr"], M - 1, 254) == 0 and both(D["Shr"], M - 1, M - 1) == 0
    assert both(D["Shr"], M - 1, 253) == (M - 1) >> 253 and both(D["Shl"], 1, 253) == 1 << 253
    assert both(D["Shl"], 7, 254) == 0 and both(D["Shl"], 7, 0) == 7
    assert both(D["Lt"], M - 1, 0) == 1 and both(D["Gt"], M - 1, 0) == 0        # negative < positive
    assert both(D["Leq"], (M >> 1), (M >> 1) + 1) == 0                          # halfM is positive, halfM+1 negative
    assert both(D["Land"], 2, 0) == 0 and both(D["Lor"], 2, 0) == 1
    assert po.eval_uno(0, 0) == 0 and po.eval_uno(0, 5) == M - 5 and cref.op_uno(0, 5) == M - 5
    assert po.eval_tres(0, 0, 7, 9) == 9 and po.eval_tres(0, 3, 7, 9) == 7
    with pytest.raises(po.ReferenceUndefined):
        po.eval_duo(D["Shl"], M - 1, 200)         # reference panics (graph.rs:634)
    with pytest.raises(po.ReferenceUndefined):
        po.eval_duo(D["Bxor"], 1 << 253, (1 << 253) ^ M)
    with pytest.raises(po.ReferenceUndefined):
        po.eval_duo(D["Pow"], 2, 3)


def test_c_oracle_equals_python_oracle_on_random_operands():
    rnd = random.Random(3)
    for _ in range(4000):
        a, b = util.random_value(rnd), util.random_value(rnd)
        for op in range(20):
            if op == 4:
                continue
            assert cref.op_duo(op, a, b) == po.eval_duo(op, a, b, "circom"), (op, a, b)
    for _ in range(20):
        a, b = util.random_value(rnd), util.random_value(rnd)
        assert cref.op_duo(4, a, b) == pow(a, b, M)


def _eval_golden(name, inputs):
    nodes, wit, imap = po.deserialize_graph(util.golden_graph(name))
    buf = po.build_input_buffer(nodes, imap, inputs)
    return po.evaluate(nodes, buf, wit)


def test_small_circuit_known_answers():
    # circuit1 105*303+2, circuit3 +3, circuit2 IsZero(105)=0 -> 2, circuit4, circuit6
    assert _eval_golden("circuit1", {"a": [105], "b": [303]})[1] == 31817
    assert _eval_golden("circuit3", {"a": [105], "b": [303]})[1] == 31818
    assert _eval_golden("circuit2", {"a": [105], "b": [303]})[1] == 2
    assert _eval_golden("circuit2", {"a": [0], "b": [303]})[1] == 305
    for v in (105, (1 << 250) + 12345, M - 1):
        assert _eval_golden("circuit6_num2bits", {"a": [v]})[1] == (v >> 16) & ((1 << 216) - 1)


def test_poseidon_known_answers():
    # test_deps/circomlib/test/poseidoncircuit.js:26-78
    assert _eval_golden("poseidon2", {"inputs": [1, 2]})[1] == \
        7853200120776062878684798364095072458815029376092732009249414926327459813530
    assert _eval_golden("poseidon2", {"inputs": [3, 4]})[1] == \
        14763215145315200506921711489642608356394854266165572616578112107564877678998


def test_sha256_known_answers():
    nodes, wit, imap = po.deserialize_graph(util.golden_graph("circuit8_sha256_512"))
    cg = cref.CGraph(util.golden_graph("circuit8_sha256_512"))
    rnd = random.Random(8)
    msgs = [bytes(range(1, 65)), bytes(64), bytes([255] * 64), rnd.randbytes(64)]
    rows = []
    for m in msgs:
        bits = [(byte >> (7 - k)) & 1 for byte in m for k in range(8)]
        rows.append([1] + bits)
    inp = np.frombuffer(b"".join(util.pack_u256(r) for r in rows), dtype=np.uint8).reshape(len(rows), 513, 32)
    out = cg.evaluate_batch(inp, 2)
    for i, m in enumerate(msgs):
        w = util.unpack_u256(out[i].tobytes())
        digest = bytes(int("".join(str(b) for b in w[1 + 8 * k:9 + 8 * k]), 2) for k in range(32))
        assert digest == hashlib.sha256(m).digest()
    assert util.unpack_u256(out[0].tobytes()) == po.evaluate(nodes, rows[0], wit)
    assert hashlib.sha256(msgs[0]).hexdigest() == "20a7ec84684f7fe124cb3727d049734ab0b7da2f52fcafbcef989ecfd91e870b"


def test_authv2_valid_proof_request():
    inp = po.deserialize_inputs(util.golden_inputs("circuit9_authV2"))
    w = _eval_golden("circuit9_authV2", inp)
    assert w[1] == inp["genesisID"][0]          # profileNonce == 0 -> userID == genesisID (authV2.circom:80)
    assert len(w) == util.manifest()["circuit9_authV2"]["n_witness"]


@pytest.mark.parametrize("name", sorted(util.manifest().keys()))
def test_golden_wtns_reproduced_by_both_oracles(name):
    man = util.manifest()[name]
    data = util.golden_graph(name)
    assert hashlib.sha256(data).hexdigest() == man["graph_sha256"]
    want = util.golden_wtns(name)
    assert hashlib.sha256(want).hexdigest() == man["wtns_sha256"]
    assert man["stats"]["constraints_violated"] == 0      # every === of the circom sources held when generated
    w = po.calc_witness(util.golden_inputs(name), data)
    assert po.wtns_from_witness(w) == want
    nodes, _, imap = po.deserialize_graph(data)
    buf = po.build_input_buffer(nodes, imap, po.deserialize_inputs(util.golden_inputs(name)))
    cg = cref.CGraph(data)
    arr = np.frombuffer(util.pack_u256(buf + [0] * (cg.n_inputs - len(buf))), dtype=np.uint8).reshape(1, cg.n_inputs, 32)
    assert po.wtns_from_witness(util.unpack_u256(cg.evaluate_batch(arr)[0].tobytes())) == want
