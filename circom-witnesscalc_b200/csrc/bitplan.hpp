// Bit-sliced plan for Boolean graphs (SHA-256, Num2Bits/BinSum towers on bit inputs).
//
// The reference evaluates every node of such a graph as a 256-bit field operation with three Montgomery
// conversions around each bitwise op (/root/reference/src/graph.rs:674-717, :637-672); the throughput plan of
// plan.hpp still spends one interpreted instruction per node and input set.  When every value of the graph is,
// for every assignment of BITS to the inputs, either a bit, a small non-negative integer or a function of a few
// bits, the whole graph is a Boolean circuit, and 32 input sets fit the 32 bits of one machine word:
//
//   * typing (compile_bit_plan): a forward pass gives every live node one of
//       CONST                      a constant,
//       TT(support, table)         a function of <= 6 bit-valued nodes given by its table of FIELD values, computed with
//                                  the exact op semantics of alu.cuh (this is what proves circomlib's polynomial gates
//                                  a*(1-2b-2c+4bc)+b+c-2bc bit-valued: interval arithmetic loses the correlation),
//       BV                         a non-negative integer below M held as bit planes (BinSum's `lin`, the word arithmetic
//                                  of sha256compression_function.circom); sums stay a bit heap until somebody looks at
//                                  their bits and are then compressed with full adders (carry-save, not ripple per term);
//     anything else makes the graph ineligible (the throughput plan runs it as before).
//   * the result is a DAG of 3-input look-up tables over bit planes; a plane is ONE 32-bit word per warp:
//     bit k = the value for input set 32 g + k of the warp's group g.  The DAG is levelised and packed into steps of
//     32 independent LUTs: in a step every LANE executes its own LUT on the group's planes in shared memory
//     (lane = instruction, word = 32 input sets), so the kernel needs no inter-warp synchronisation at all.
//   * the typing holds under the CONTRACT that the inputs are bits.  The kernel checks it for every input set while it
//     packs the planes; the input sets that break it are collected and evaluated by the generic kernel afterwards, so the
//     result is exact for every input (graph.rs semantics), only slower for those sets.
//   * witness values are bits: the eval kernel leaves [group][position] plane words in HBM and an expansion kernel
//     writes the 32-byte canonical rows with 512-byte coalesced stores -- the only HBM-heavy part of the path.
#pragma once
#include <string>
#include <vector>

#include "graph.hpp"

namespace gw {

struct BitOp { uint32_t x, y, z, w; };   // x: lut[7:0]; y: slot a | slot b << 16; z: slot c | dst slot << 16; w: witness position / BIT_NO_POS
static const uint32_t BIT_NO_SLOT = 0xFFFFu;
static const uint32_t BIT_NO_POS = 0xFFFFFFFFu;
static const uint32_t BIT_SLOT_ZERO = 0, BIT_SLOT_ONES = 1;   // fixed planes: all zeros, all ones
static const uint32_t BIT_CONTRACT = 0xFFFFFFFFu;

struct BitPlanOptions {
  uint32_t max_support = 6;      // table size limit of the TT domain (2^6 field values)
  uint32_t max_slots = 12000;    // plane slots per warp (4 B each in shared memory)
  bool merge_luts = true;        // compose single-use LUTs into their reader when the support stays <= 3
};

struct BitPlan {
  bool eligible = false;
  std::string reason;            // why not, when not eligible
  uint32_t n_inputs = 0, n_witness = 0;
  // triples (input index, plane slot or BIT_NO_SLOT, bit): bit = BIT_CONTRACT: the input must be 0 or 1 (checked per input
  // set) and is its own plane; otherwise the plane is bit `bit` of the input reduced mod M (an input that is only ever
  // taken apart with Shr/Band -- Num2Bits of a field element -- carries no contract).  Sorted by input index.
  std::vector<uint32_t> inputs;
  bool has_field_inputs = false;
  std::vector<BitOp> code;       // n_steps x 32 LUT instructions, one per lane (padding: dst = BIT_NO_SLOT, w = BIT_NO_POS)
  uint32_t n_steps = 0, n_slots = 2;
  std::vector<int32_t> const_of_pos;   // per witness position: -1 = a bit plane, -2 = a wide value (below), else index into const_vals
  std::vector<U256> const_vals;
  // witness values that are integers of several bits (Bits2Num sums, field inputs passed through): their planes are
  // stored behind the W position planes of a group, at W + base + bit.  Triples (position, base, number of planes).
  std::vector<uint32_t> wide;
  uint32_t plane_stride = 0;     // plane words per group: n_witness + all wide planes
  // statistics
  uint64_t n_luts = 0, n_levels = 0, n_nodes_bit = 0, n_nodes_tt = 0, n_nodes_bv = 0, n_full_adders = 0, n_merged = 0;
};

BitPlan compile_bit_plan(const Graph& g, const BitPlanOptions& opt);

// value of a 3-input look-up table on three words (bit k of the result = lut[a_k | b_k << 1 | c_k << 2])
static inline uint32_t bit_lut3(uint32_t lut, uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r = 0;
  for (uint32_t k = 0; k < 8; k++)
    if ((lut >> k) & 1u) r |= ((k & 1u) ? a : ~a) & ((k & 2u) ? b : ~b) & ((k & 4u) ? c : ~c);
  return r;
}

}  // namespace gw
