"""GPU parity at BASELINE.json's full batch sizes (configs 2, 3 and a slice of 4), through the C ABI's device-buffer
entry point: every row (or a dense sample) bit-exact against the C restatement of the reference, plus the
size-independent properties the circuits offer (SHA-256 digests against hashlib, Num2Bits recomposition, purity of
duplicated rows, a fold over ALL rows compared between two differently shaped launches)."""
import hashlib
import importlib

import numpy as np
import pytest

from tests import util
from tests.util import po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    torch = pytest.importorskip("torch")
    cwc = importlib.import_module("circom-witnesscalc_b200")
    assert cwc.device_count() >= 1 and torch.cuda.is_available(), "no CUDA device visible"
    from oracle import cref
    return torch, cwc, cref


def _run_device(torch, g, host_in):
    """host uint8 [B, I, 32] -> device witness tensor [B, W*32] via gw_calc_witness_batch_device"""
    B = host_in.shape[0]
    d_in = torch.from_numpy(host_in.reshape(B, -1)).cuda()
    d_out = torch.empty((B, g.n_witness * 32), dtype=torch.uint8, device="cuda")
    g.calc_witness_batch_device(0, d_in.data_ptr(), B, d_out.data_ptr(), None, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return d_out


def _fold(torch, d_out):
    """order-independent checksum of checksums: per-row int64 sums, then their sum (wrapping)"""
    v = d_out.view(torch.int64)
    return int(v.sum(dim=1).sum().item())


@pytest.mark.parametrize("name,seed", [("circuit6_num2bits", 6), ("circuit7_poseidon4", 7)])
def test_config2_65536_sets_every_row(env, name, seed):
    torch, cwc, cref = env
    data = util.golden_graph(name)
    g = cwc.Graph(data)
    B = 65536
    rng = np.random.Generator(np.random.PCG64(seed))
    vals = util.random_field_batch(rng, (B, g.n_inputs))
    vals[:, 0, :] = 0
    vals[:, 0, 0] = 1
    host_in = vals.view(np.uint8).reshape(B, g.n_inputs, 32)
    d_out = _run_device(torch, g, host_in)
    got = d_out.cpu().numpy().reshape(B, g.n_witness, 32)
    want = cref.CGraph(data).evaluate_batch(host_in, 8)
    assert (got == want).all()
    if name == "circuit6_num2bits":
        # witness = [1, out, in, 256 bits ...]: the bit signals are bits (Shr then Band 1, bitify.circom:32)
        bits = got[:, 3:, :]
        assert (bits[:, :, 1:] == 0).all() and (bits[:, :, 0] <= 1).all()
    # a differently shaped launch (ragged: not a multiple of the CTA size) gives the same rows
    d2 = _run_device(torch, g, host_in[:B - 77])
    assert _fold(torch, d2) == _fold(torch, d_out[:B - 77])


def test_config3_sha256_16384_sets(env):
    torch, cwc, cref = env
    name = "circuit8_sha256_512"
    data = util.golden_graph(name)
    g = cwc.Graph(data)
    B = 16384
    rng = np.random.Generator(np.random.PCG64(8))
    host_in = np.zeros((B, g.n_inputs, 32), dtype=np.uint8)
    host_in[:, :, 0] = rng.integers(0, 2, size=(B, g.n_inputs), dtype=np.uint8)
    host_in[:, 0, 0] = 1
    host_in[B - 1] = host_in[3]                                        # purity: same inputs, same witness
    d_out = _run_device(torch, g, host_in)
    assert bool((d_out[B - 1] == d_out[3]).all())
    # digest = witness[1..257] (sha256.circom:77-79), message = the 512 input bits, big-endian bit order per byte
    rows = list(range(0, B, B // 64))[:64]
    sample = d_out[rows].cpu().numpy().reshape(len(rows), g.n_witness, 32)
    for k, b in enumerate(rows):
        msg_bits = host_in[b, 1:513, 0]
        msg = bytes(int("".join(str(int(x)) for x in msg_bits[8 * i:8 * i + 8]), 2) for i in range(64))
        digest_bits = sample[k, 1:257, 0]
        assert (sample[k, 1:257, 1:] == 0).all()
        got = bytes(int("".join(str(int(x)) for x in digest_bits[8 * i:8 * i + 8]), 2) for i in range(32))
        assert got == hashlib.sha256(msg).digest(), b
    # dense sample against the C oracle, all witness positions
    rows2 = list(range(5, B, B // 256))[:256]
    want = cref.CGraph(data).evaluate_batch(host_in[rows2], 8)
    got2 = d_out[rows2].cpu().numpy().reshape(len(rows2), g.n_witness, 32)
    assert (got2 == want).all()
    # every value of a SHA-256 witness on bit inputs is a bit
    v = d_out.view(torch.int64).view(B, g.n_witness, 4)
    assert int(v[:, :, 1:].abs().sum().item()) == 0 and int(v[:, :, 0].max().item()) <= 1 and int(v[:, :, 0].min().item()) >= 0


def test_config4_authv2_slice_and_launch_shapes(env):
    """authV2: row 0 = the reference's own (valid) fixture -> the committed golden .wtns; random rows against the
    C oracle; the same rows through launches of different shapes (one full wave of 148 x 32 witnesses, a ragged
    one, the host-buffer entry point) give the same bytes."""
    torch, cwc, cref = env
    name = "circuit9_authV2"
    data = util.golden_graph(name)
    nodes, wit, imap = po.deserialize_graph(data)
    g = cwc.Graph(data)
    B = 148 * 32
    rng = np.random.Generator(np.random.PCG64(9))
    vals = util.random_field_batch(rng, (B, g.n_inputs))
    for key in ("authClaimNonRevMtpNoAux", "gistMtpNoAux"):
        off, ln = g.input_signals[key]
        vals[:, off:off + ln, :] = 0
        vals[:, off:off + ln, 0] = rng.integers(0, 2, size=(B, ln), dtype=np.uint64)
    vals[:, 0, :] = 0
    vals[:, 0, 0] = 1
    buf = po.build_input_buffer(nodes, imap, po.deserialize_inputs(util.golden_inputs(name)))
    vals[0] = np.frombuffer(util.pack_u256(buf), dtype=np.uint64).reshape(g.n_inputs, 4)
    host_in = vals.view(np.uint8).reshape(B, g.n_inputs, 32)
    d_out = _run_device(torch, g, host_in)
    row0 = d_out[0].cpu().numpy().tobytes()
    assert po.wtns_from_witness(util.unpack_u256(row0)) == util.golden_wtns(name)
    rows = [1, 31, 32, 33, 1000, 2047, 2048, B - 1] + list(range(7, B, 311))
    want = cref.CGraph(data).evaluate_batch(host_in[rows], 8)
    got = d_out[rows].cpu().numpy().reshape(len(rows), g.n_witness, 32)
    assert (got == want).all()
    d2 = _run_device(torch, g, host_in[:1000])
    assert bool((d2 == d_out[:1000]).all())
    h = g.calc_witness_batch(host_in[990:1000 + 27])
    assert (h.reshape(37, -1) == d_out[990:1027].cpu().numpy()).all()
