"""Host path of csrc/field.cuh (the algorithms the CUDA kernels run: Barrett multiplication, the
Montgomery helpers, shifts, bitwise ops, signed comparisons, long division, pow, inversion) against
Python integers.  The device path (PTX carry chains) is covered on the GPU by test_gpu_parity.py."""
import ctypes
import random

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


def test_field_and_integer_ops_random_and_edges():
    L = util.field_host_lib()
    rnd = random.Random(1)
    R = (1 << 256) % M
    Rinv = pow(R, -1, M)
    for t in range(6000):
        a, b = util.random_value(rnd), util.random_value(rnd)
        assert call2(L.t_mul, a, b) == a * b % M
        r16 = A16(); L.t_mul_wide(enc(a), enc(b), r16); assert dec(r16) == a * b
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


def test_barrett_worst_cases():
    L = util.field_host_lib()
    for a in (M - 1, M - 2, (M >> 1) + 1, 1 << 253, (1 << 253) + 1):
        for b in (M - 1, M - 2, (M >> 1), (1 << 253) - 1, 3):
            assert call2(L.t_mul, a, b) == a * b % M
