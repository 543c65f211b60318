// One graph operation on canonical operands: the per-node semantics of the reference's
// Operation::eval_fr (/root/reference/src/graph.rs:102-144), UnoOperation::eval_fr (:188-197) and
// TresOperation::eval_fr (:221-225), shared by the CUDA kernels (device path) and by the host-side
// plan simulator used in the CPU unit tests (tests/csrc/plan_host_sim.cpp).
#pragma once
#include "field.cuh"
#include "inv_safegcd.cuh"
#include "isa.h"

namespace gw {

// ---- narrow values (isa.h: F_NARROW) -------------------------------------------------------------------
// two's-complement int64 in limbs 0..1
GW_HD int64_t narrow_of(const fe& a) { return (int64_t)((uint64_t)a.l[0] | ((uint64_t)a.l[1] << 32)); }
// canonical field element of a narrow value: v >= 0 -> v, v < 0 -> M + v
GW_HD fe fe_from_narrow(int64_t v) {
  const uint32_t ext = v < 0 ? 0xFFFFFFFFu : 0u;
  uint32_t t[8], m[8];
  t[0] = (uint32_t)(uint64_t)v; t[1] = (uint32_t)((uint64_t)v >> 32);
#pragma unroll
  for (int i = 2; i < 8; i++) t[i] = ext;
#pragma unroll
  for (int i = 0; i < 8; i++) m[i] = MOD_L(i) & ext;
  fe r;
  u256_add(r.l, t, m);           // the carry out of bit 255 cancels the sign extension
  return r;
}
// One graph operation on narrow operands.  The plan compiler only emits a narrow instruction when the true
// result fits (so wrapping int64 arithmetic is exact) and, for the bitwise ops and shifts, when the operands
// are provably non-negative (so the int64 bits ARE the canonical bits the reference works on).
GW_HD int64_t narrow_exec(uint32_t op, int64_t a, int64_t b, int64_t c) {
  const uint64_t ua = (uint64_t)a, ub = (uint64_t)b;
  switch (op) {
    case OP_MUL: return (int64_t)(ua * ub);                                  // graph.rs:105
    case OP_SQR: return (int64_t)(ua * ua);
    case OP_ADD: return (int64_t)(ua + ub);                                  // graph.rs:110
    case OP_SUB: return (int64_t)(ua - ub);                                  // graph.rs:111
    case OP_NEG: return (int64_t)(0 - ua);                                   // graph.rs:190-194
    case OP_EQ: return a == b;                                               // graph.rs:122-125
    case OP_NEQ: return a != b;                                              // graph.rs:126-129
    case OP_LT: return a < b;                                                // graph.rs:130-133: signed readings
    case OP_GT: return a > b;
    case OP_LEQ: return a <= b;
    case OP_GEQ: return a >= b;
    case OP_LAND: return (a != 0) && (b != 0);                               // graph.rs:134
    case OP_LOR: return (a != 0) || (b != 0);                                // graph.rs:135
    case OP_SHR: return ub >= 64 ? 0 : (int64_t)(ua >> (ub & 63));           // graph.rs:637-672, a < 2^62
    case OP_SHL: return ub >= 64 ? 0 : (int64_t)(ua << (ub & 63));           // graph.rs:621-635, result < 2^62
    case OP_BOR: return (int64_t)(ua | ub);                                  // graph.rs:689-702
    case OP_BAND: return (int64_t)(ua & ub);                                 // graph.rs:674-687
    case OP_BXOR: return (int64_t)(ua ^ ub);                                 // graph.rs:704-717
    case OP_NZ1: return a == 0 ? 1 : a;
    case OP_TERN: return a != 0 ? b : c;                                     // graph.rs:221-225
    default: return 0;
  }
}
// narrow OP_SHRAND: (a >> k) & c, a >= 0
GW_HD int64_t narrow_shr_and(int64_t a, uint32_t k, int64_t c) { return k >= 64 ? 0 : (int64_t)(((uint64_t)a >> k) & (uint64_t)c); }
// narrow OP_DOT term
GW_HD int64_t narrow_dot_term(int64_t acc, uint32_t kind, int64_t x, int64_t c) {
  const uint64_t u = (uint64_t)acc;
  return (int64_t)(kind == T_MAC ? u + (uint64_t)x * (uint64_t)c : kind == T_ADDHI ? u + (uint64_t)x : kind == T_SUBHI ? u - (uint64_t)x : u + (uint64_t)c);
}

// op: an Opcode that is not a data-movement op.  C is only read for OP_TERN.  st collects StatusBits.
GW_HD fe alu_exec(uint32_t op, const fe& A, fe Bv, const fe& C, uint32_t& st) {
  fe R;
  if (op == OP_DIV) { Bv = fe_inv(Bv); op = OP_MUL; }                        // graph.rs:109 (b == 0 -> 0)
  else if (op == OP_POW) { R = fe_pow(A, Bv); st |= ST_POW; }
  switch (op) {
    case OP_MUL: R = fe_mul(A, Bv); break;                                   // graph.rs:105
    case OP_SQR: R = fe_sqr(A); break;                                       // Mul(a, a)
    case OP_ADD: R = fe_add(A, Bv); break;                                   // graph.rs:110
    case OP_SUB: R = fe_sub(A, Bv); break;                                   // graph.rs:111
    case OP_POW: break;
    case OP_IDIV:                                                            // graph.rs:112-116
    case OP_MOD: {                                                           // graph.rs:117-121
      fe q, r;
      bool z = fe_is_zero(Bv);
      fe d = Bv; if (z) d = fe_small(1);
      u256_divrem(A, d, &q, &r);
      R = z ? fe_zero() : (op == OP_IDIV ? q : r);
      break;
    }
    case OP_EQ: R = fe_small(u256_eq(A.l, Bv.l) ? 1u : 0u); break;          // graph.rs:122-125
    case OP_NEQ: R = fe_small(u256_eq(A.l, Bv.l) ? 0u : 1u); break;         // graph.rs:126-129
    case OP_LT: R = fe_small(fe_cmp(A, Bv, 0) ? 1u : 0u); break;            // graph.rs:130-133
    case OP_GT: R = fe_small(fe_cmp(A, Bv, 1) ? 1u : 0u); break;
    case OP_LEQ: R = fe_small(fe_cmp(A, Bv, 2) ? 1u : 0u); break;
    case OP_GEQ: R = fe_small(fe_cmp(A, Bv, 3) ? 1u : 0u); break;
    case OP_LAND: R = fe_small((!fe_is_zero(A) && !fe_is_zero(Bv)) ? 1u : 0u); break;   // graph.rs:134
    case OP_LOR: R = fe_small((!fe_is_zero(A) || !fe_is_zero(Bv)) ? 1u : 0u); break;    // graph.rs:135
    case OP_SHL: { bool ov; R = fe_shl(A, Bv, &ov); if (ov) st |= ST_SHL_OVERFLOW; break; }
    case OP_SHR: R = fe_shr(A, Bv); break;
    case OP_BOR: { bool e; R = fe_bitop(A, Bv, 1, &e); if (e) st |= ST_BITWISE_EQ_M; break; }
    case OP_BAND: { bool e; R = fe_bitop(A, Bv, 0, &e); break; }
    case OP_BXOR: { bool e; R = fe_bitop(A, Bv, 2, &e); if (e) st |= ST_BITWISE_EQ_M; break; }
    case OP_NEG: R = fe_neg(A); break;                                       // graph.rs:190-194
    case OP_ID: R = A; st |= ST_ID; break;
    case OP_LNOT: R = fe_small(fe_is_zero(A) ? 1u : 0u); st |= ST_LNOT_BNOT; break;
    case OP_BNOT: R = fe_bnot(A); st |= ST_LNOT_BNOT; break;
    case OP_INV: R = fe_inv(A); break;                                       // plan compiler only (batched Div)
    case OP_NZ1: { bool z = fe_is_zero(A); R = A; R.l[0] = z ? 1u : A.l[0]; break; }
    case OP_WIDEN: R = fe_from_narrow(narrow_of(A)); break;                  // plan compiler only (narrow -> wide)
    case OP_TERN: {                                                          // graph.rs:221-225
      bool z = fe_is_zero(A);
#pragma unroll
      for (int i = 0; i < 8; i++) R.l[i] = z ? C.l[i] : Bv.l[i];
      break;
    }
    default: R = fe_zero(); break;
  }
  return R;
}

GW_HD bool op_has_b(uint32_t op) { return op < 32 || op == OP_TERN; }

// OP_SHRAND: (a >> k) & c for a shift amount 0 < k < 254 known at plan time (graph.rs:637-672 then :674-687; the
// result of the AND is <= c < M, so bit_and's reduction never fires)
GW_HD fe fe_shr_and(const fe& a, uint32_t k, const fe& c) {
  fe r = u256_shr(a, k);
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] &= c.l[i];
  return r;
}

}  // namespace gw
