"""GPU parity tests: the CUDA path, called through the C ABI, against the Python oracle and the
committed golden .wtns files.  Bit-exact comparison everywhere (integer arithmetic)."""
import importlib
import random

import numpy as np
import pytest

from tests import util
from tests.util import po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cwc():
    mod = importlib.import_module("circom-witnesscalc_b200")
    assert mod.device_count() >= 1, "no CUDA device visible"
    return mod


GOLDEN = ["circuit1", "circuit2", "circuit3", "circuit4", "circuit5_poseidon", "circuit6_num2bits",
          "circuit7_poseidon4", "poseidon2", "circuit11_key_expansion", "circuit8_sha256_512", "circuit9_authV2"]


@pytest.mark.parametrize("name", GOLDEN)
def test_golden_wtns_drop_in(cwc, name):
    """gw_calc_witness(inputs.json, graph.bin) == committed .wtns, byte for byte (test_circuits.sh:81 `cmp`)."""
    got = cwc.calc_witness_wtns(util.golden_inputs(name), util.golden_graph(name))
    assert got == util.golden_wtns(name)


def test_random_graphs_all_ops(cwc):
    rnd = random.Random(1234)
    for t in range(12):
        nodes, wit, imap = util.random_graph(rnd, n_ops=400)
        g = cwc.Graph(po.serialize_graph(nodes, wit, imap))
        B = 96 + t            # not a multiple of the CTA size: ragged last tile
        rows = [[1] + [util.random_value(rnd) if rnd.random() < 0.8 else rnd.randrange(1 << 256) for _ in range(6)]
                for _ in range(B)]
        inp = np.frombuffer(b"".join(util.pack_u256(r) for r in rows), dtype=np.uint8).reshape(B, 7, 32)
        out = g.calc_witness_batch(inp)
        for b in range(0, B, 7):
            want = po.evaluate(nodes, rows[b], wit, "circom")
            assert util.unpack_u256(out[b].tobytes()) == want, (t, b)


def test_reference_undefined_flags(cwc):
    """cases where the reference panics get the circom value and a per-set flag"""
    nodes = [(po.K_INPUT, 0), (po.K_INPUT, 1), (po.K_INPUT, 2),
             (po.K_DUO, po.DUO["Shl"], 1, 2), (po.K_DUO, po.DUO["Bxor"], 1, 2), (po.K_DUO, po.DUO["Pow"], 1, 2)]
    g = cwc.Graph(po.serialize_graph(nodes, [0, 3, 4, 5], {"a": (1, 1), "b": (2, 1)}))
    M = po.M
    rows = [[1, 5, 3], [1, M - 1, 200], [1, M, 0], [1, (1 << 253), (1 << 253) ^ M]]
    inp = np.frombuffer(b"".join(util.pack_u256(r) for r in rows), dtype=np.uint8).reshape(4, 3, 32)
    out, flags = g.calc_witness_batch(inp, want_flags=True)
    for b, r in enumerate(rows):
        assert util.unpack_u256(out[b].tobytes()) == po.evaluate(nodes, r, [0, 3, 4, 5], "circom")
    assert flags[0] == 4 and flags[1] & 1 and flags[3] & 2


def test_batch_matches_single_and_properties(cwc):
    """batch of random inputs on Poseidon(2): every row equals the oracle; duplicated rows give
    duplicated witnesses (purity); size-independent check on a larger batch via row sampling."""
    name = "poseidon2"
    data = util.golden_graph(name)
    nodes, wit, imap = po.deserialize_graph(data)
    g = cwc.Graph(data)
    rng = np.random.default_rng(7)
    B = 4096 + 37
    vals = util.random_field_batch(rng, (B, g.n_inputs))
    vals[:, 0, :] = 0
    vals[:, 0, 0] = 1
    vals[B - 1] = vals[0]
    inp = vals.view(np.uint8).reshape(B, g.n_inputs, 32)
    out = g.calc_witness_batch(inp)
    assert (out[B - 1] == out[0]).all()
    for b in [0, 1, 31, 32, 127, 128, 4095, 4096, B - 2]:
        row = util.limbs_to_ints(vals[b])
        assert util.unpack_u256(out[b].tobytes()) == po.evaluate(nodes, row, wit)


def test_bad_inputs_return_errors(cwc):
    data = util.golden_graph("circuit1")
    with pytest.raises(cwc.WitnessCalcError):
        cwc.calc_witness_wtns('{"a": "1", "zzz": "2"}', data)          # unknown key (reference: panic)
    with pytest.raises(cwc.WitnessCalcError):
        cwc.calc_witness_wtns('{"a": ["1", "2"]}', data)                # wrong length (reference: panic)
    with pytest.raises(cwc.WitnessCalcError):
        cwc.calc_witness_wtns('{"a": -1}', data)                        # lib.rs:211-214
    with pytest.raises(cwc.WitnessCalcError):
        cwc.calc_witness_wtns('{"a": "1"', data)                        # invalid JSON (reference: panic)
    with pytest.raises(cwc.WitnessCalcError):
        cwc.calc_witness_wtns('{"a": "1"}', b"not a graph file at all....")


LATENCY_ENVS = [{}, {"GW_LAT_CHAIN": "0"}, {"GW_LAT_WARPS": "2", "GW_LAT_SLOW_WARPS": "1", "GW_LAT_D": "1"},
                {"GW_LAT_WARPS": "5", "GW_LAT_D": "60", "GW_LAT_SPLIT": "0"}, {"GW_LAT_FUSE": "0"}]


@pytest.mark.parametrize("env", LATENCY_ENVS)
def test_latency_mode_random_graphs_and_golden(cwc, monkeypatch, env):
    """single-witness latency mode (one CTA: level-synchronous main warps, asynchronous slow warps, lane chains):
    gw_calc_witness_latency.  The plan options are read when the latency plan of a graph is first built."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    rnd = random.Random(4321)
    for t in range(10):
        nodes, wit, imap = util.random_graph(rnd, n_ops=500)
        g = cwc.Graph(po.serialize_graph(nodes, wit, imap))
        for _ in range(2):
            row = [1] + [util.random_value(rnd) if rnd.random() < 0.8 else rnd.randrange(1 << 256) for _ in range(6)]
            out, ms = g.calc_witness_latency(np.frombuffer(util.pack_u256(row), dtype=np.uint8).reshape(7, 32))
            assert util.unpack_u256(out.tobytes()) == po.evaluate(nodes, row, wit, "circom"), t
    for name in ("circuit5_poseidon", "circuit6_num2bits", "circuit8_sha256_512", "circuit9_authV2"):
        data = util.golden_graph(name)
        nodes, wit, imap = po.deserialize_graph(data)
        g = cwc.Graph(data)
        buf = po.build_input_buffer(nodes, imap, po.deserialize_inputs(util.golden_inputs(name)))
        if env.get("GW_LAT_FUSE") == "0" and name == "circuit9_authV2":
            # one instruction per graph node keeps more values alive than the shared-memory value file holds: the
            # latency entry point says so, and the drop-in gw_calc_witness falls back to the throughput kernel
            with pytest.raises(cwc.WitnessCalcError, match="too wide"):
                g.calc_witness_latency(np.frombuffer(util.pack_u256(buf), dtype=np.uint8).reshape(g.n_inputs, 32))
            assert cwc.calc_witness_wtns(util.golden_inputs(name), data) == util.golden_wtns(name)
            continue
        for _ in range(3):                                   # repeated launches reuse the uploaded plan
            out, ms = g.calc_witness_latency(np.frombuffer(util.pack_u256(buf), dtype=np.uint8).reshape(g.n_inputs, 32))
            assert po.wtns_from_witness(util.unpack_u256(out.tobytes())) == util.golden_wtns(name)
        print(f"latency mode {name} {env}: {ms:.2f} ms kernel")


def test_poseidon_like_graphs_batch_and_latency(cwc):
    """random graphs with Poseidon's shapes (OP_POW5, straight-line OP_DOT shapes, chains, divisions) on the device:
    a small batch through the throughput kernel and single witnesses through the latency kernel"""
    for seed in range(24):
        rnd = random.Random(5000 + seed)
        nodes, wit, imap = util.poseidon_like_graph(rnd, rnd.choice([2, 3, 5]), rnd.choice([3, 8, 20]))
        g = cwc.Graph(po.serialize_graph(nodes, wit, imap))
        rows = [[1] + [util.random_value(rnd) if rnd.random() < 0.8 else rnd.randrange(1 << 256) for _ in range(6)] for _ in range(5)]
        inp = np.frombuffer(b"".join(util.pack_u256(r) for r in rows), dtype=np.uint8).reshape(len(rows), 7, 32)
        out = g.calc_witness_batch(inp)
        for b, row in enumerate(rows):
            want = po.evaluate(nodes, row, wit, "circom")
            assert util.unpack_u256(out[b].tobytes()) == want, (seed, b)
            if b < 2:
                lat, _ = g.calc_witness_latency(inp[b])
                assert util.unpack_u256(lat.tobytes()) == want, (seed, b, "latency")


def test_single_witness_through_batch_kernel(cwc, monkeypatch):
    monkeypatch.setenv("GW_SINGLE_MODE", "batch")
    for name in ("circuit2", "circuit5_poseidon", "circuit9_authV2"):
        assert cwc.calc_witness_wtns(util.golden_inputs(name), util.golden_graph(name)) == util.golden_wtns(name)


def test_multi_gpu_shards_host_api(cwc):
    """gw_calc_witness_batch(..., n_gpus = all visible): contiguous shards, one host thread per GPU, no collective.
    With one visible GPU this still exercises the sharding entry point (n_gpus = 1)."""
    n = cwc.device_count()
    name = "circuit7_poseidon4"
    data = util.golden_graph(name)
    nodes, wit, imap = po.deserialize_graph(data)
    g = cwc.Graph(data)
    rng = np.random.default_rng(77)
    B = 5000 + 13
    vals = util.random_field_batch(rng, (B, g.n_inputs))
    vals[:, 0, :] = 0
    vals[:, 0, 0] = 1
    inp = vals.view(np.uint8).reshape(B, g.n_inputs, 32)
    one = g.calc_witness_batch(inp, n_gpus=1)
    many = g.calc_witness_batch(inp, n_gpus=n)
    assert (one == many).all()
    for b in sorted({0, B // n - 1, min(B // n, B - 1), B - 1}):
        assert util.unpack_u256(many[b].tobytes()) == po.evaluate(nodes, util.limbs_to_ints(vals[b]), wit)
    print(f"sharded over {n} GPU(s)")
    # gw_calc_witness_batch_on: the same call on an explicit device range (one process per GPU passes its own device)
    out = np.empty_like(one)
    g.calc_witness_batch_ptr(inp.ctypes.data, B, out.ctypes.data, n_gpus=1, first_device=n - 1)
    assert (out == one).all()
    with pytest.raises(cwc.WitnessCalcError, match="device range"):
        g.calc_witness_batch_ptr(inp.ctypes.data, B, out.ctypes.data, n_gpus=1, first_device=n)
    with pytest.raises(cwc.WitnessCalcError, match="exceeds"):
        g.calc_witness_batch(inp, n_gpus=n + 1)


def test_batch_wtns_framing_select_and_batch_cli(cwc, tmp_path):
    """JSON Lines -> packed inputs -> .wtns images framed by the pitched D2H copy (byte-identical to the
    single-witness entry point); selected-signals graphs return exactly those witness positions; the batch CLI
    writes the same files."""
    import json
    import subprocess
    name = "circuit7_poseidon4"
    data = util.golden_graph(name)
    nodes, wit, imap = po.deserialize_graph(data)
    g = cwc.Graph(data)
    rnd = random.Random(11)
    recs = [json.loads(util.golden_inputs(name))] + [{"a": [str(rnd.randrange(po.M)) for _ in range(4)]} for _ in range(300)]
    text = "\n".join(json.dumps(r) for r in recs)
    inp = g.parse_inputs_batch(text)
    fsz = g.wtns_file_size()
    for pitch in (fsz, fsz + 52):
        files = g.calc_witness_batch_wtns(inp, pitch)
        assert files[0, :fsz].tobytes() == util.golden_wtns(name)
        for i in (1, 150, 300):
            assert files[i, :fsz].tobytes() == cwc.calc_witness_wtns(json.dumps(recs[i]), data)
        assert (files[:, fsz:] == 0).all()
    full = g.calc_witness_batch(inp)
    pos = [5, 0, g.n_witness - 1, 5, 17]
    sel = g.select(pos)
    part = sel.calc_witness_batch(inp)
    assert part.shape == (301, 5, 32) and (part == full[:, pos, :]).all()
    assert sel.calc_witness_wtns(json.dumps(recs[3]))[76:] == full[3, pos, :].tobytes()
    # batch CLI
    gp, ip, od = tmp_path / "g.bin", tmp_path / "in.jsonl", tmp_path / "out"
    gp.write_bytes(data); ip.write_text(text)
    cli = cwc.CLI_PATH + "-batch"
    r = subprocess.run([cli, str(gp), str(ip), str(od)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "301 input sets parsed" in r.stdout and "Witnesses generated in:" in r.stdout
    assert (od / "00000000.wtns").read_bytes() == util.golden_wtns(name)
    assert (od / "00000300.wtns").read_bytes() == files[300, :fsz].tobytes()
    r = subprocess.run([cli], capture_output=True, text=True)
    assert r.returncode == 1 and "Usage:" in r.stderr
