// Modular inversion mod M by Bernstein-Yang "safegcd" divsteps (eprint 2019/266), constant-time
// variant, on signed 30-bit limbs: 20 rounds of 30 divsteps.  Every lane of a warp executes exactly
// the same instruction sequence (no data-dependent branches), which is what a SIMT evaluator needs;
// Fermat's a^(M-2) costs ~380 field multiplications, this costs about 25 multiplications' worth of
// issue slots, almost all of them on the ALU pipe instead of the multiplier.
//
// Used for Operation::Div (/root/reference/src/graph.rs:109: b == 0 -> 0, else a * b^-1); the
// reference gets inversion from ark-ff's Fp::inverse (binary extended Euclid, data-dependent loops).
//
// Invariants (f, g, d, e as in the paper, x the input):  d*x = f (mod M), e*x = g (mod M).
// Start f = M, g = x, d = 0, e = 1.  After >= 590 divsteps g = 0 and f = +-gcd = +-1, so x^-1 = +-d.
// For x = 0: g stays 0, d stays 0, the result is 0, which is the value Div needs.
#pragma once
#include "field.cuh"

namespace gw {

struct s30 { int32_t v[9]; };   // value = sum v[i] * 2^(30 i)

GW_HD constexpr int32_t MOD30(int i) {
  return i == 0 ? 0x30000001 : i == 1 ? 0x0f87d64f : i == 2 ? 0x1b970914 : i == 3 ? 0x0cfa121e : i == 4 ? 0x01585d28 :
         i == 5 ? 0x0116da06 : i == 6 ? 0x1a029b85 : i == 7 ? 0x139cb84c : 0x3064;
}
static const uint32_t MOD_INV30 = 0x10000001u;   // M^-1 mod 2^30
static const int32_t MASK30 = 0x3FFFFFFF;

GW_HD s30 s30_from_u256(const uint32_t* a) {
  s30 r;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    const int bit = 30 * i, w = bit >> 5, s = bit & 31;
    uint32_t lo = a[w] >> s;
    uint32_t hi = (s > 2 && w + 1 < 8) ? (a[w + 1] << (32 - s)) : 0u;
    r.v[i] = (int32_t)((lo | hi) & (uint32_t)MASK30);
  }
  return r;
}
// limbs must be normalised to [0, 2^30) and the value < 2^256
GW_HD void s30_to_u256(uint32_t* a, const s30& r) {
#pragma unroll
  for (int w = 0; w < 8; w++) {
    const int bit = 32 * w, i = bit / 30, s = bit - 30 * i;          // word w starts inside limb i at offset s
    uint32_t lo = (uint32_t)r.v[i] >> s;
    uint32_t hi = (i + 1 < 9) ? ((uint32_t)r.v[i + 1] << (30 - s)) : 0u;
    uint32_t hi2 = (s > 28 && i + 2 < 9) ? ((uint32_t)r.v[i + 2] << (60 - s)) : 0u;
    a[w] = lo | hi | hi2;
  }
}

// 30 divsteps on the low 30 bits; returns the new zeta and the transition matrix t = (u v; q r)
// with t * (f, g) = 2^30 * (f', g').
GW_HD int32_t divsteps_30(int32_t zeta, uint32_t f0, uint32_t g0, int32_t* t) {
  uint32_t u = 1, v = 0, q = 0, r = 1;
  uint32_t f = f0, g = g0;
#pragma unroll 6
  for (int i = 0; i < 30; i++) {
    uint32_t c1 = (uint32_t)(zeta >> 31);          // all ones if zeta < 0
    uint32_t c2 = (uint32_t)0 - (g & 1u);          // all ones if g is odd
    uint32_t x = (f ^ c1) - c1, y = (u ^ c1) - c1, z = (v ^ c1) - c1;   // conditionally negated f, u, v
    g += x & c2; q += y & c2; r += z & c2;
    c1 &= c2;                                      // zeta < 0 and g odd: swap roles
    zeta = (int32_t)(((uint32_t)zeta ^ c1) - 1u);
    f += g & c1; u += q & c1; v += r & c1;
    g >>= 1; u <<= 1; v <<= 1;
  }
  t[0] = (int32_t)u; t[1] = (int32_t)v; t[2] = (int32_t)q; t[3] = (int32_t)r;
  return zeta;
}

// (d, e) <- t * (d, e) / 2^30 mod M, keeping both in (-2M, M)
GW_HD void update_de_30(s30& d, s30& e, const int32_t* t) {
  const int32_t u = t[0], v = t[1], q = t[2], r = t[3];
  const int32_t sd = d.v[8] >> 31, se = e.v[8] >> 31;
  int32_t md = (u & sd) + (v & se), me = (q & sd) + (r & se);
  int32_t di = d.v[0], ei = e.v[0];
  int64_t cd = (int64_t)u * di + (int64_t)v * ei;
  int64_t ce = (int64_t)q * di + (int64_t)r * ei;
  md -= (int32_t)((MOD_INV30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)MASK30);
  me -= (int32_t)((MOD_INV30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)MASK30);
  cd += (int64_t)MOD30(0) * md;
  ce += (int64_t)MOD30(0) * me;
  cd >>= 30; ce >>= 30;                            // the low 30 bits are zero by construction
#pragma unroll
  for (int i = 1; i < 9; i++) {
    di = d.v[i]; ei = e.v[i];
    cd += (int64_t)u * di + (int64_t)v * ei;
    ce += (int64_t)q * di + (int64_t)r * ei;
    cd += (int64_t)MOD30(i) * md;
    ce += (int64_t)MOD30(i) * me;
    d.v[i - 1] = (int32_t)cd & MASK30; cd >>= 30;
    e.v[i - 1] = (int32_t)ce & MASK30; ce >>= 30;
  }
  d.v[8] = (int32_t)cd; e.v[8] = (int32_t)ce;
}

// (f, g) <- t * (f, g) / 2^30 (exact)
GW_HD void update_fg_30(s30& f, s30& g, const int32_t* t) {
  const int32_t u = t[0], v = t[1], q = t[2], r = t[3];
  int32_t fi = f.v[0], gi = g.v[0];
  int64_t cf = (int64_t)u * fi + (int64_t)v * gi;
  int64_t cg = (int64_t)q * fi + (int64_t)r * gi;
  cf >>= 30; cg >>= 30;
#pragma unroll
  for (int i = 1; i < 9; i++) {
    fi = f.v[i]; gi = g.v[i];
    cf += (int64_t)u * fi + (int64_t)v * gi;
    cg += (int64_t)q * fi + (int64_t)r * gi;
    f.v[i - 1] = (int32_t)cf & MASK30; cf >>= 30;
    g.v[i - 1] = (int32_t)cg & MASK30; cg >>= 30;
  }
  f.v[8] = (int32_t)cf; g.v[8] = (int32_t)cg;
}

// d in (-2M, M) -> [0, M), negated first if sign < 0
GW_HD void normalize_30(s30& r, int32_t sign) {
  int32_t c = 0;
  int32_t cond_add = r.v[8] >> 31;
  const int32_t cond_neg = sign >> 31;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    int32_t x = r.v[i] + (MOD30(i) & cond_add);
    x = (x ^ cond_neg) - cond_neg;
    x += c;
    c = (i < 8) ? (x >> 30) : 0;
    r.v[i] = (i < 8) ? (x & MASK30) : x;
  }
  cond_add = r.v[8] >> 31;
  c = 0;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    int32_t x = r.v[i] + (MOD30(i) & cond_add) + c;
    c = (i < 8) ? (x >> 30) : 0;
    r.v[i] = (i < 8) ? (x & MASK30) : x;
  }
  // one more conditional subtraction covers a value in [M, 2M) after negation of a value in (-2M, -M]
  int32_t t[9];
  c = 0;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    int32_t x = r.v[i] - MOD30(i) + c;
    c = (i < 8) ? (x >> 30) : 0;
    t[i] = (i < 8) ? (x & MASK30) : x;
  }
  const int32_t keep = t[8] >> 31;                  // all ones if r - M < 0: keep r
#pragma unroll
  for (int i = 0; i < 9; i++) r.v[i] = (r.v[i] & keep) | (t[i] & ~keep);
}

// x^-1 mod M for x in [0, M); 0 -> 0
GW_HD_NOINLINE fe fe_inv(const fe& x) {
  s30 f, g = s30_from_u256(x.l), d, e;
#pragma unroll
  for (int i = 0; i < 9; i++) { f.v[i] = MOD30(i); d.v[i] = 0; e.v[i] = 0; }
  e.v[0] = 1;
  int32_t zeta = -1;
#pragma unroll 1
  for (int it = 0; it < 20; it++) {
    int32_t t[4];
    zeta = divsteps_30(zeta, (uint32_t)f.v[0] | ((uint32_t)f.v[1] << 30), (uint32_t)g.v[0] | ((uint32_t)g.v[1] << 30), t);
    update_de_30(d, e, t);
    update_fg_30(f, g, t);
  }
  normalize_30(d, f.v[8]);
  fe r;
  s30_to_u256(r.l, d);
  return r;
}

}  // namespace gw
