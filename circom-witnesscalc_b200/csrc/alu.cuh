// One graph operation on canonical operands: the per-node semantics of the reference's
// Operation::eval_fr (/root/reference/src/graph.rs:102-144), UnoOperation::eval_fr (:188-197) and
// TresOperation::eval_fr (:221-225), shared by the CUDA kernels (device path) and by the host-side
// plan simulator used in the CPU unit tests (tests/csrc/plan_host_sim.cpp).
#pragma once
#include "field.cuh"
#include "inv_safegcd.cuh"
#include "isa.h"

namespace gw {

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
