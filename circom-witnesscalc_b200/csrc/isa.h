// Device "ISA" of the batch witness evaluator: what plan.cpp emits and eval_kernels.cu executes.
//
// One instruction = 16 bytes (uint4):
//   x : opcode[7:0] | flags[15:8] | dst register[31:16]   (dst = 0xFFFF: value is not kept)
//   y : operand a   (register index, constant-table index, input index or spill slot)
//   z : operand b
//   w : operand c (TernCond) or witness position (when F_OUT is set)
//
// Opcodes 0..19 are the reference's DuoOp numbers (protos/messages.proto:5-26), so a graph
// Op(op,a,b) maps 1:1; the rest are unary / ternary / data-movement ops of this design.
#pragma once
#include <stdint.h>

namespace gw {

enum Opcode : uint32_t {
  OP_MUL = 0, OP_DIV = 1, OP_ADD = 2, OP_SUB = 3, OP_POW = 4, OP_IDIV = 5, OP_MOD = 6,
  OP_EQ = 7, OP_NEQ = 8, OP_LT = 9, OP_GT = 10, OP_LEQ = 11, OP_GEQ = 12, OP_LAND = 13,
  OP_LOR = 14, OP_SHL = 15, OP_SHR = 16, OP_BOR = 17, OP_BAND = 18, OP_BXOR = 19,
  OP_NEG = 32, OP_ID = 33, OP_LNOT = 34, OP_BNOT = 35,
  OP_TERN = 40,
  OP_INPUT = 48,     // dst <- inputs[w][a] mod M
  OP_SPILL_ST = 49,  // spill[b] <- reg a
  OP_SPILL_LD = 50,  // dst <- spill[a]
  OP_OUT = 51,       // witness[w] <- a (register or constant)
  OP_SQR = 52,       // dst <- a*a   (Mul with both operands the same node)
  OP_NOP = 63,
};

enum Flags : uint32_t {
  F_A_CONST = 1u << 8,
  F_B_CONST = 1u << 9,
  F_C_CONST = 1u << 10,
  F_OUT = 1u << 11,      // also store the result to witness position .w
  F_PAIR = 1u << 12,     // MUL/SQR only: the next slot is an independent MUL/SQR issued together (both
                         // read their operands before either writes); never straddles a 32-slot block
};

static const uint32_t NO_DST = 0xFFFFu;

struct Instr { uint32_t x, y, z, w; };

static inline Instr make_instr(uint32_t op, uint32_t flags, uint32_t dst, uint32_t a, uint32_t b, uint32_t c) {
  Instr i; i.x = (op & 0xFFu) | (flags & 0xFF00u) | (dst << 16); i.y = a; i.z = b; i.w = c; return i;
}

// per-witness status bits (the reference panics / is unimplemented in these cases)
enum StatusBits : uint32_t {
  ST_SHL_OVERFLOW = 1u, ST_BITWISE_EQ_M = 2u, ST_POW = 4u, ST_ID = 8u, ST_LNOT_BNOT = 16u,
};

}  // namespace gw
