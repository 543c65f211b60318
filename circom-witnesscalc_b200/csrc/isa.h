// Device "ISA" of the batch witness evaluator: what plan.cpp emits and engine.cu executes.
//
// The program is a sequence of 16-byte slots (uint4).  An instruction is one header slot,
//   x : opcode[7:0] | flags[15:8] | dst register[31:16]   (dst = 0xFFFF: value is not kept)
//   y : operand a   (register index, constant-table index, input index or spill slot)
//   z : operand b
//   w : operand c (TernCond) or witness position (when F_OUT is set)
// followed, for OP_DOT only, by ceil(n_terms / 2) slots holding two 8-byte terms each.
//
// Opcodes 0..19 are the reference's DuoOp numbers (protos/messages.proto:5-26), so a graph
// Op(op,a,b) maps 1:1; 32.. are unary ops (32 + UnoOp number for the reference's Neg/Id), the rest are
// ternary / data-movement / fused ops of this design.
#pragma once
#include <stdint.h>

namespace gw {

enum Opcode : uint32_t {
  OP_MUL = 0, OP_DIV = 1, OP_ADD = 2, OP_SUB = 3, OP_POW = 4, OP_IDIV = 5, OP_MOD = 6,
  OP_EQ = 7, OP_NEQ = 8, OP_LT = 9, OP_GT = 10, OP_LEQ = 11, OP_GEQ = 12, OP_LAND = 13,
  OP_LOR = 14, OP_SHL = 15, OP_SHR = 16, OP_BOR = 17, OP_BAND = 18, OP_BXOR = 19,
  OP_NEG = 32, OP_ID = 33, OP_LNOT = 34, OP_BNOT = 35,
  OP_INV = 36,       // dst <- a^-1 mod M (0 -> 0); only produced by the plan compiler (batched Div)
  OP_NZ1 = 37,       // dst <- (a == 0) ? 1 : a;    only produced by the plan compiler (batched Div)
  OP_WIDEN = 38,     // dst <- canonical 256-bit form of the NARROW value a; only produced by the plan compiler
  OP_TERN = 40,
  OP_INPUT = 48,     // dst <- inputs[w][a] mod M
  OP_SPILL_ST = 49,  // spill[b] <- reg a
  OP_SPILL_LD = 50,  // dst <- spill[a]
  OP_OUT = 51,       // witness[w] <- a (register or constant)
  OP_SQR = 52,       // dst <- a*a   (Mul with both operands the same node)
  OP_DOT = 53,       // dst <- sum of terms mod M with ONE Montgomery reduction; y = n_terms | n_cond_sub << 8
  OP_SHRAND = 54,    // dst <- (a >> k) & const;  z = k[7:0] | constant index << 8  (Num2Bits: Band(Shr(x, k), 1))
  OP_POW5 = 55,      // dst <- a^5 with a^2, a^4 also stored as witness values (Poseidon S-box: x2 = x*x; x4 = x2*x2; x5 = x4*x);
                     // .y = register of a | d4 << 16, .z = witness position of a^2 (NO_POS: none), a^4 goes to position
                     // .z + d4 (d4 = 0xFFFF: none), .w = position of a^5 (F_OUT).  Throughput plan only.
  OP_POW4 = 56,      // dst <- a^4, a^2 stored at witness position .z (NO_POS: none), a^4 at .w (F_OUT).  Latency plans only.
  OP_MULADD = 57,    // dst <- a * b + c (mod M); .w is operand c, so a witness store is a separate OP_OUT (like TernCond).
                     // Latency plans only: S-box to S-box links, plan.cpp rewrite_sbox_links.
  OP_NOP = 63,
};

enum Flags : uint32_t {
  F_A_CONST = 1u << 8,
  F_B_CONST = 1u << 9,
  F_C_CONST = 1u << 10,
  F_OUT = 1u << 11,      // also store the result to witness position .w
  F_NARROW = 1u << 12,   // narrow instruction: operands and result are signed 64-bit integers (see below)
};

// Narrow values.  The plan compiler proves, by interval arithmetic over the graph (sound for EVERY input:
// inputs themselves are never assumed small), that some values always lie in (-2^62, 2^62) when read as
// signed field elements (x > M/2 means x - M, the reading of the reference's comparisons, graph.rs:723-769).
// Such a value is computed by a narrow instruction (F_NARROW) on a two's-complement int64 kept in limbs 0..1
// of its register (limbs 2..7 are NOT written); add/sub/mul wrap mod 2^64, which is exact because the true
// result fits.  A register written by a wide instruction whose value is provably in [0, 2^62) can be read by
// a narrow instruction directly (canonical == zero-extended); a narrow register read by a wide instruction
// goes through OP_WIDEN first.  Witness stores of narrow results convert to canonical (v < 0 -> M + v).
// With F_NARROW: OP_DOT terms are (reg * c64 | +reg | -reg | c64) with plain (not pre-scaled) int64 constants
// in limbs 0..1 of the constant table entry; OP_SPILL_ST / OP_SPILL_LD move 8 bytes; OP_OUT converts.

// OP_DOT terms.  A term is (lo, hi): lo = kind[3:0] | register << 16, hi = constant-table index.
// The accumulator P is a 512-bit integer; the result is P * 2^-256 mod M (Montgomery reduction), so
// constants are stored pre-multiplied by 2^256 mod M and plain values enter at bit 256.
enum TermKind : uint32_t {
  T_MAC = 0,     // P += reg * const'            const' = c * 2^256 mod M   (or (M - c) * 2^256 for a subtracted term)
  T_ADDHI = 1,   // P += reg << 256              (+ value)
  T_SUBHI = 2,   // P += (M - reg) << 256        (- value)
  T_CONST = 3,   // P += const'                  const' = c * 2^256 mod M
};
static const uint32_t DOT_MAX_TERMS = 16;

static const uint32_t NO_DST = 0xFFFFu;
static const uint32_t NO_POS = 0xFFFFFFFFu;

struct Instr { uint32_t x, y, z, w; };

static inline Instr make_instr(uint32_t op, uint32_t flags, uint32_t dst, uint32_t a, uint32_t b, uint32_t c) {
  Instr i; i.x = (op & 0xFFu) | (flags & 0xFF00u) | (dst << 16); i.y = a; i.z = b; i.w = c; return i;
}
// number of slots of the instruction whose header is `h`
static inline uint32_t instr_slots(const Instr& h) {
  return ((h.x & 0xFFu) == OP_DOT) ? 1u + (((h.y & 0xFFu) + 1u) >> 1) : 1u;
}

// per-witness status bits (the reference panics / is unimplemented in these cases)
enum StatusBits : uint32_t {
  ST_SHL_OVERFLOW = 1u, ST_BITWISE_EQ_M = 2u, ST_POW = 4u, ST_ID = 8u, ST_LNOT_BNOT = 16u,
};

}  // namespace gw
