"""Plan compiler (csrc/plan.cpp: liveness, constant table, register allocation, Belady spilling,
witness placement) executed by the host simulator with the same alu_exec() the kernel uses, against
the Python oracle.  No GPU needed."""
import random

import pytest

from tests import util
from tests.util import po


def _check(nodes, wit, imap, n_regs, rnd, n_rows=3, n_in=6):
    g = util.SimGraph(po.serialize_graph(nodes, wit, imap), n_regs)
    for _ in range(n_rows):
        inp = [1] + [util.random_value(rnd) if rnd.random() < 0.8 else rnd.randrange(1 << 256) for _ in range(n_in)]
        got, _ = g.eval(inp)
        assert got == po.evaluate(nodes, inp, wit, "circom")
    return g


@pytest.mark.parametrize("n_regs", [4, 5, 8, 24, 64])
def test_random_graphs_all_ops(n_regs):
    rnd = random.Random(100 + n_regs)
    for _ in range(25):
        nodes, wit, imap = util.random_graph(rnd, n_ops=250)
        g = _check(nodes, wit, imap, n_regs, rnd)
        if n_regs <= 5:
            assert g.info["spill_st"] > 0          # tiny register files must spill, and still be exact


def test_edge_graphs():
    rnd = random.Random(5)
    # empty witness / constants only / inputs only / duplicated witness entries / dead nodes
    _check([(po.K_INPUT, 0)], [], {}, 8, rnd, n_in=0)
    _check([(po.K_INPUT, 0), (po.K_CONST, 7)], [1, 1, 0, 1], {}, 8, rnd, n_in=0)
    nodes = [(po.K_INPUT, i) for i in range(7)] + [(po.K_DUO, 0, 1, 2), (po.K_DUO, 2, 7, 7), (po.K_DUO, 0, 3, 3)]
    _check(nodes, [0, 8, 8, 3, 3, 7, 1], {"x": (1, 6)}, 8, rnd)
    # TernCond that is itself a witness signal, with constant operands
    nodes = [(po.K_INPUT, 0), (po.K_INPUT, 1), (po.K_CONST, 5), (po.K_CONST, 9), (po.K_TRES, 0, 1, 2, 3), (po.K_TRES, 0, 2, 1, 4)]
    _check(nodes, [0, 4, 5, 4], {"x": (1, 1)}, 4, rnd, n_in=1)
    # an input index gap (tree_shake removes unused inputs)
    nodes = [(po.K_INPUT, 0), (po.K_INPUT, 3), (po.K_DUO, 2, 1, 1)]
    g = util.SimGraph(po.serialize_graph(nodes, [0, 2], {"x": (1, 3)}), 8)
    assert g.info["I"] == 4 and g.eval([1, 0, 0, 21])[0] == [1, 42]


def test_register_pressure_chain():
    """a wide fan-in followed by uses in reverse order defeats any small register file"""
    rnd = random.Random(9)
    nodes = [(po.K_INPUT, i) for i in range(7)]
    first = len(nodes)
    for k in range(60):
        nodes.append((po.K_DUO, 0, 1 + k % 6, 1 + (k + 1) % 6))
    acc = first
    for k in range(59, -1, -1):
        nodes.append((po.K_DUO, 2, acc, first + k))
        acc = len(nodes) - 1
    g = _check(nodes, [0, acc], {"x": (1, 6)}, 6, rnd)
    assert g.info["n_spill"] >= 40


@pytest.mark.parametrize("name,n_regs", [("circuit5_poseidon", 8), ("circuit6_num2bits", 24), ("circuit11_key_expansion", 16),
                                         ("circuit8_sha256_512", 24), ("circuit9_authV2", 24), ("circuit9_authV2", 12)])
def test_golden_circuits(name, n_regs):
    data = util.golden_graph(name)
    g = util.SimGraph(data, n_regs)
    got, st = g.eval(g.inputs_from_json(util.golden_inputs(name)))
    assert st == 0 and po.wtns_from_witness(got) == util.golden_wtns(name)
    man = util.manifest()[name]
    assert g.info["W"] == man["n_witness"] and g.info["I"] == man["n_inputs"] and g.info["n_nodes"] == man["n_nodes"]
