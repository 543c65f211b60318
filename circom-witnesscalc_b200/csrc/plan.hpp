// Graph -> device program ("plan").
//
// This is the load-time half of what replaces /root/reference/src/graph.rs:367-391 (evaluate): the
// reference walks Vec<Node> and keeps one Fr per node (N x 32 B per witness).  For a batch that does
// not fit (authV2: 194 k nodes -> 6 MB per witness), so the node list is compiled once into a linear
// instruction stream over a small per-witness register file (shared memory on the device) with
// Belady-style spilling to a witness-major spill area in HBM.  Dead nodes are dropped, constants go
// to a deduplicated table, witness positions are attached to the defining instruction.
#pragma once
#include "graph.hpp"
#include "isa.h"

namespace gw {

struct PlanOptions {
  uint32_t n_regs = 12;      // per-witness registers kept in shared memory (12 x 32 B x 128 threads = 48 KB per CTA: 4 CTAs/SM)
  bool pair_muls = true;     // schedule independent multiplications next to each other and issue them as pairs
  uint32_t pair_window = 24; // how far ahead (in nodes) a partner is searched
};

struct PlanStats {
  uint64_t graph_nodes = 0, graph_ops = 0;   // ops = Op + UnoOp + TresOp nodes of the file (node-ops/s metric)
  uint64_t live_ops = 0;                     // ops reachable from the witness
  uint64_t instrs = 0, spill_st = 0, spill_ld = 0, outs = 0, mul_pairs = 0;
  uint64_t op_count[64] = {0};               // executed instructions by opcode
  uint32_t max_live = 0;                     // peak number of simultaneously live values
};

struct Plan {
  std::vector<Instr> code;
  std::vector<U256> consts;
  uint32_t n_regs = 0, n_spill = 0;
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
