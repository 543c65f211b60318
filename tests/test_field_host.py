"""csrc/field.cuh (the algorithms the CUDA kernels run: Barrett multiplication, the Montgomery helpers and
reduction, shifts, bitwise ops, signed comparisons, long division, pow, inversion) against Python integers, twice:
the plain host path, and the DEVICE formulations (PTX carry chains on even/odd accumulators) compiled for the host
with an emulated carry flag (-DGW_EMULATE_PTX).  The real PTX is covered on the GPU by test_gpu_parity.py."""
import ctypes
import random

import pytest

from tests import util
from tests.util import M

A8 = ctypes.c_uint32 * 8
A16 = ctypes.c_uint32 * 16


def enc(x):
    return A8(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)])


def dec(a):
    return sum(int(v) << (32 * i) for i, v in enumerate(a))


def call2(f, a, b):
    r = A8()
    f(enc(a), enc(b), r)
    return dec(r)


def call1(f, a):
    r = A8()
    f(enc(a), r)
    return dec(r)


@pytest.mark.parametrize("ptx", [False, True])
def test_field_and_integer_ops_random_and_edges(ptx):
    L = util.field_host_lib(ptx)
    assert bool(L.t_emulates_ptx()) == ptx
    rnd = random.Random(1)
    R = (1 << 256) % M
    Rinv = pow(R, -1, M)
    for t in range(6000):
        a, b = util.random_value(rnd), util.random_value(rnd)
        assert call2(L.t_mul, a, b) == a * b % M
        r16 = A16(); L.t_mul_wide(enc(a), enc(b), r16); assert dec(r16) == a * b
        r16 = A16(); L.t_sqr_wide(enc(a), r16); assert dec(r16) == a * a
        assert call1(L.t_sqr, a) == a * a % M
        r8 = A8(); L.t_mul_lo(enc(a), enc(b), r8); assert dec(r8) == (a * b) % (1 << 256)
        assert call2(L.t_mont_mul, a, b) == a * b * Rinv % M
        assert call1(L.t_to_mont, a) == a * R % M and call1(L.t_from_mont, a) == a * Rinv % M
        assert call2(L.t_add, a, b) == (a + b) % M and call2(L.t_sub, a, b) == (a - b) % M
        assert call1(L.t_neg, a) == (-a) % M
        x = rnd.randrange(1 << 256)
        assert call1(L.t_reduce, x) == x % M
        sh = rnd.choice([0, 1, 31, 32, 33, 63, 64, 100, 128, 200, 252, 253, 254, 255, 256, M - 1, rnd.randrange(300)])
        assert call2(L.t_shr, a, sh) == (a if sh == 0 else 0 if sh >= 254 else a >> sh)
        r = A8(); ov = L.t_shl(enc(a), enc(sh), r)
        full = a if sh == 0 else 0 if sh >= 254 else (a << sh) & ((1 << 256) - 1)
        if full < M:
            assert ov == 0 and dec(r) == full
        else:
            e2 = (a << sh) & ((1 << 254) - 1)
            assert ov == 1 and dec(r) == (e2 - M if e2 >= M else e2)
        for w, f in enumerate([lambda x, y: x & y, lambda x, y: x | y, lambda x, y: x ^ y]):
            r = A8(); eq = L.t_bitop(enc(a), enc(b), w, r); d = f(a, b)
            assert eq == (d == M) and dec(r) == (d - M if d >= M else d)
        na, nb = a > (M >> 1), b > (M >> 1)
        for w in range(4):
            exp = [na, nb, na, nb][w] if na != nb else [a < b, a > b, a <= b, a >= b][w]
            assert bool(L.t_cmp(enc(a), enc(b), w)) == exp
        nb_ = (~a) & ((1 << 254) - 1)
        assert call1(L.t_bnot, a) == (nb_ - M if nb_ >= M else nb_)
        if t < 150:
            if b:
                q, rr = A8(), A8(); L.t_divrem(enc(a), enc(b), q, rr); assert dec(q) == a // b and dec(rr) == a % b
            assert call2(L.t_pow, a, b) == pow(a, b, M)
            assert call1(L.t_inv, a) == (pow(a, -1, M) if a else 0)


@pytest.mark.parametrize("ptx", [False, True])
def test_barrett_worst_cases(ptx):
    L = util.field_host_lib(ptx)
    for a in ((1 << 256) - 1, (1 << 256) - (1 << 32), 0xFFFFFFFF, ((1 << 256) - 1) // 3, 1 << 255, 0):
        r16 = A16(); L.t_sqr_wide(enc(a), r16); assert dec(r16) == a * a
    for a in (M - 1, M - 2, (M >> 1) + 1, 1 << 253, (1 << 253) + 1):
        for b in (M - 1, M - 2, (M >> 1), (1 << 253) - 1, 3):
            assert call2(L.t_mul, a, b) == a * b % M


@pytest.mark.parametrize("ptx", [False, True])
def test_montgomery_reduction_of_accumulated_terms(ptx):
    """OP_DOT: P = sum of terms (value x pre-scaled constant, +-value << 256, constant), then ONE reduction
    P * 2^-256 mod M with the number of conditional subtractions the plan compiler derives from the term mix."""
    L = util.field_host_lib(ptx)
    rnd = random.Random(2)
    R = 1 << 256
    Rinv = pow(R, -1, M)
    edge = [0, 1, M - 1, M - 2, (1 << 253), (1 << 224) - 1, 0xFFFFFFFF, (1 << 32), (M >> 1)]
    pick = lambda: rnd.choice(edge) if rnd.random() < 0.4 else rnd.randrange(M)
    # raw 512-bit values: anything whose reduced value stays below 2^ncs * M and 2^256
    for t in range(3000):
        ncs = rnd.choice([1, 2, 3])
        lim = min((1 << ncs) * M, R) - M        # (P + mM) / R < P / R + M must stay below the bound
        P = rnd.randrange(lim * R) if rnd.random() < 0.7 else rnd.choice([0, 1, R - 1, R, lim * R - 1, (lim - 1) * R, (1 << 32) - 1,
                                                                              ((1 << 256) - 1) << 32, (lim * R - 1) & ~((1 << 224) - 1)])
        P %= lim * R
        r = A8(); L.t_mont_reduce(A16(*[(P >> (32 * i)) & 0xFFFFFFFF for i in range(16)]), ncs, r)
        assert dec(r) == P * Rinv % M, (t, ncs, hex(P))
    # term by term, like the kernel
    for t in range(1500):
        n_mac, n_hi = rnd.randrange(0, 9), rnd.randrange(0, 4)
        if n_mac + n_hi == 0:
            continue
        bound = 1 + 0.18906 * n_mac + n_hi
        if bound > 5.25:
            continue
        ncs = 1 if bound <= 2 else 2 if bound <= 4 else 3
        kinds, xs, cs, want = [], [], [], 0
        def term(k, x, c):
            kinds.append(k); xs.extend(enc(x)); cs.extend(enc(c))
        for _ in range(n_mac):
            x, c = pick(), pick()
            term(0, x, c * R % M); want += x * c
        for _ in range(n_hi):
            x = pick()
            if rnd.random() < 0.5:
                term(1, x, 0); want += x
            else:
                term(2, x, 0); want -= x
        if rnd.random() < 0.5:
            c = pick(); term(3, 0, c * R % M); want += c
        rnd.shuffle(order := list(range(len(kinds))))
        kinds2 = [kinds[i] for i in order]
        xs2 = [v for i in order for v in xs[8 * i:8 * i + 8]]
        cs2 = [v for i in order for v in cs[8 * i:8 * i + 8]]
        n = len(kinds2)
        r = A8(); L.t_dot_eval((ctypes.c_uint32 * n)(*kinds2), (ctypes.c_uint32 * (8 * n))(*xs2), (ctypes.c_uint32 * (8 * n))(*cs2), n, ncs, r)
        assert dec(r) == want % M, (t, n_mac, n_hi)
