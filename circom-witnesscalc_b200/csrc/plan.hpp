// Graph -> device program ("plan").
//
// This is the load-time half of what replaces /root/reference/src/graph.rs:367-391 (evaluate): the
// reference walks Vec<Node> and keeps one Fr per node (N x 32 B per witness).  For a batch that does
// not fit (authV2: 194 k nodes -> 6 MB per witness), so the node list is compiled once into a linear
// instruction stream over a small per-witness register file (shared memory on the device) with
// Belady-style spilling to a witness-major spill area in HBM.  Dead nodes are dropped, constants go
// to a deduplicated table, witness positions are attached to the defining instruction.
//
// Two graph-level rewrites run first (both preserve every node value bit for bit):
//   * Div batching: Div nodes on the same "division level" are independent, so groups of up to
//     `div_batch` of them share ONE modular inversion (Montgomery's trick: 4k-3 multiplications + 1
//     inversion instead of k inversions), with b == 0 -> 0 kept exact (graph.rs:109).
//   * Linear-combination fusion: trees of Add/Sub whose inner nodes are single-use temporaries and whose
//     leaves are Mul(value, constant) become one OP_DOT: products are accumulated in 512 bits and
//     reduced once (the reference's build-circuit emits `lc += c * x` chains).
#pragma once
#include "graph.hpp"
#include "isa.h"

namespace gw {

struct PlanOptions {
  uint32_t n_regs = 12;      // per-witness registers kept in shared memory (12 x 32 B x 512 threads = 192 KB)
  uint32_t div_batch = 8;    // max Div nodes sharing one inversion (1 = off)
  bool fuse_dot = true;      // fuse linear combinations into OP_DOT
  uint32_t max_terms = 8;    // max terms of one OP_DOT (<= DOT_MAX_TERMS, <= n_regs - 3)
  bool narrow = true;        // narrow typing: provably small values are computed on int64 (isa.h: F_NARROW)
};

struct PlanStats {
  uint64_t graph_nodes = 0, graph_ops = 0;   // ops = Op + UnoOp + TresOp nodes of the file (node-ops/s metric)
  uint64_t live_ops = 0;                     // ops of the file reachable from the witness
  uint64_t instrs = 0, slots = 0, spill_st = 0, spill_ld = 0, outs = 0;
  uint64_t op_count[64] = {0};               // emitted instructions by opcode
  uint64_t dot_terms[4] = {0, 0, 0, 0};      // emitted OP_DOT terms by TermKind
  uint64_t inversions = 0;                   // modular inversions per witness (OP_DIV + OP_INV)
  uint64_t div_nodes = 0;                    // live Div nodes of the graph
  uint64_t mul_nodes = 0;                    // live Mul nodes of the graph
  uint32_t max_live = 0;                     // peak number of simultaneously live values
  uint64_t narrow_instrs = 0;                // emitted narrow (int64) instructions, spill moves excluded
};

struct Plan {
  std::vector<Instr> code;   // slots
  std::vector<U256> consts;
  uint32_t n_regs = 0, n_spill = 0;
  uint32_t n_spill_narrow = 0;   // 8-byte spill slots of narrow values (a pool of their own: [slot][thread] uint2)
  uint32_t n_inputs = 0;     // I: length of the input buffer incl. slot 0
  uint32_t n_witness = 0;    // W
  PlanStats stats;
};

Plan compile_plan(const Graph& g, const PlanOptions& opt);

// Single-witness ("latency") mode: the same instruction format, but instructions are grouped into
// dependency levels; all instructions of a level are independent and are executed by different threads
// of ONE CTA, with a CTA barrier between levels.  Operands name slots of one shared-memory value file
// (slots are recycled only at level boundaries, so there are no hazards inside a level).
struct LatencyPlan {
  std::vector<Instr> code;            // level after level
  std::vector<uint32_t> level_count;  // instructions per level
  std::vector<U256> consts;
  uint32_t n_slots = 0, n_inputs = 0, n_witness = 0;
  uint32_t max_level_width = 0;
};
LatencyPlan compile_latency_plan(const Graph& g, uint32_t max_slots);

}  // namespace gw
