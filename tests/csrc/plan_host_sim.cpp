// TEST-ONLY host simulator of the device plan: executes the instruction stream that plan.cpp emits
// with the host path of csrc/alu.cuh, one witness at a time.  It exists so that the graph codec, the
// inputs parser, the register allocator / spill logic and the op semantics can be checked against the
// oracle on a machine without a GPU.  It is NOT part of libcircom_witnesscalc.so (the product has no
// CPU fallback); the same alu_exec() runs on the device inside eval_batch_kernel.
#include <string.h>

#include <memory>

#include "../../circom-witnesscalc_b200/csrc/alu.cuh"
#include "../../circom-witnesscalc_b200/csrc/graph.hpp"
#include "../../circom-witnesscalc_b200/csrc/inputs.hpp"
#include "../../circom-witnesscalc_b200/csrc/plan.hpp"
#include "../../circom-witnesscalc_b200/csrc/wtns.hpp"

using namespace gw;

struct SimGraph { Graph g; Plan plan; std::string err; };

static fe to_fe(const U256& v) { fe r; memcpy(r.l, v.l, 32); return r; }

extern "C" {

SimGraph* sim_load2(const uint8_t* data, size_t len, uint32_t n_regs, int pair, char* err, size_t errlen);
SimGraph* sim_load(const uint8_t* data, size_t len, uint32_t n_regs, char* err, size_t errlen) {
  return sim_load2(data, len, n_regs, 1, err, errlen);
}
SimGraph* sim_load2(const uint8_t* data, size_t len, uint32_t n_regs, int pair, char* err, size_t errlen) {
  try {
    std::unique_ptr<SimGraph> s(new SimGraph());
    s->g = deserialize_witnesscalc_graph(data, len);
    PlanOptions o; o.n_regs = n_regs; o.pair_muls = pair != 0;
    s->plan = compile_plan(s->g, o);
    return s.release();
  } catch (const std::exception& e) { if (err && errlen) { strncpy(err, e.what(), errlen - 1); err[errlen - 1] = 0; } return nullptr; }
}
void sim_free(SimGraph* s) { delete s; }
void sim_info(SimGraph* s, uint64_t* out) {
  out[0] = s->g.nodes.size(); out[1] = s->plan.n_inputs; out[2] = s->plan.n_witness; out[3] = s->plan.code.size();
  out[4] = s->plan.n_regs; out[5] = s->plan.n_spill; out[6] = s->plan.stats.spill_ld; out[7] = s->plan.stats.spill_st;
  out[8] = s->plan.stats.max_live; out[9] = s->plan.consts.size(); out[10] = s->plan.stats.live_ops; out[11] = s->plan.stats.graph_ops;
  out[12] = s->plan.stats.mul_pairs;
}
// inputs: I x 32 B, witness: W x 32 B; returns status bits, or -1 on a malformed plan
int64_t sim_eval(SimGraph* s, const uint8_t* inputs, uint8_t* witness) {
  const Plan& p = s->plan;
  std::vector<fe> rf(p.n_regs, fe_zero()), spill(p.n_spill, fe_zero());
  uint32_t st = 0;
  for (size_t pc = 0; pc < p.code.size(); pc++) {
    const Instr& ins = p.code[pc];
    uint32_t op = ins.x & 0xFF, dst = ins.x >> 16;
    if (op == OP_NOP) continue;
    if (ins.x & F_PAIR) {                       // pair semantics: read all four operands, then write
      if (pc + 1 >= p.code.size() || (pc & 31) == 31) return -1;
      const Instr& in2 = p.code[++pc];
      auto ld = [&](uint32_t idx, bool is_const, fe* o) { if (is_const) { if (idx >= p.consts.size()) return false; *o = to_fe(p.consts[idx]); } else { if (idx >= p.n_regs) return false; *o = rf[idx]; } return true; };
      fe A1, B1, A2, B2, R1, R2;
      uint32_t op2 = in2.x & 0xFF, dst2 = in2.x >> 16;
      if ((op != OP_MUL && op != OP_SQR) || (op2 != OP_MUL && op2 != OP_SQR)) return -1;
      if (!ld(ins.y, ins.x & F_A_CONST, &A1) || !ld(in2.y, in2.x & F_A_CONST, &A2)) return -1;
      if (op == OP_SQR) B1 = A1; else if (!ld(ins.z, ins.x & F_B_CONST, &B1)) return -1;
      if (op2 == OP_SQR) B2 = A2; else if (!ld(in2.z, in2.x & F_B_CONST, &B2)) return -1;
      fe_mul2(A1, B1, A2, B2, R1, R2);
      if (dst != NO_DST) { if (dst >= p.n_regs) return -1; rf[dst] = R1; }
      if (dst2 != NO_DST) { if (dst2 >= p.n_regs) return -1; rf[dst2] = R2; }
      if (ins.x & F_OUT) { if (ins.w >= p.n_witness) return -1; memcpy(witness + 32 * (size_t)ins.w, R1.l, 32); }
      if (in2.x & F_OUT) { if (in2.w >= p.n_witness) return -1; memcpy(witness + 32 * (size_t)in2.w, R2.l, 32); }
      continue;
    }
    fe A = fe_zero(), B = fe_zero(), C = fe_zero(), R;
    if (op == OP_SPILL_ST) { if (ins.y >= p.n_regs || ins.z >= p.n_spill) return -1; spill[ins.z] = rf[ins.y]; continue; }
    if (op == OP_SPILL_LD) { if (dst >= p.n_regs || ins.y >= p.n_spill) return -1; rf[dst] = spill[ins.y]; continue; }
    if (op == OP_INPUT) {
      if (ins.y >= p.n_inputs) return -1;
      fe v; memcpy(v.l, inputs + 32 * (size_t)ins.y, 32);
      R = fe_reduce256(v);
    } else {
      auto load = [&](uint32_t idx, bool is_const, fe* o) { if (is_const) { if (idx >= p.consts.size()) return false; *o = to_fe(p.consts[idx]); } else { if (idx >= p.n_regs) return false; *o = rf[idx]; } return true; };
      if (!load(ins.y, ins.x & F_A_CONST, &A)) return -1;
      if (op == OP_OUT) { if (ins.w >= p.n_witness) return -1; memcpy(witness + 32 * (size_t)ins.w, A.l, 32); continue; }
      if (op_has_b(op) && !load(ins.z, ins.x & F_B_CONST, &B)) return -1;
      if (op == OP_TERN && !load(ins.w, ins.x & F_C_CONST, &C)) return -1;
      R = alu_exec(op, A, B, C, st);
    }
    if (dst != NO_DST) { if (dst >= p.n_regs) return -1; rf[dst] = R; }
    if (ins.x & F_OUT) { if (ins.w >= p.n_witness) return -1; memcpy(witness + 32 * (size_t)ins.w, R.l, 32); }
  }
  return st;
}
// inputs JSON -> I x 32 B buffer; returns 0 ok / 1 error (message in err)
int sim_inputs(SimGraph* s, const char* json, uint8_t* buf, char* err, size_t errlen) {
  try {
    InputList in = deserialize_inputs(json, strlen(json));
    std::vector<U256> b = build_inputs_buffer(s->g, in);
    memcpy(buf, b.data(), b.size() * 32);
    return 0;
  } catch (const std::exception& e) { if (err && errlen) { strncpy(err, e.what(), errlen - 1); err[errlen - 1] = 0; } return 1; }
}
// re-serialise the parsed graph (codec round trip); returns the size, copies up to cap bytes
size_t sim_reserialize(SimGraph* s, uint8_t* out, size_t cap) {
  std::vector<uint8_t> v = serialize_witnesscalc_graph(s->g);
  if (out && cap >= v.size()) memcpy(out, v.data(), v.size());
  return v.size();
}
void sim_wtns_header(uint32_t n, uint8_t* dst) { wtns_write_header(dst, n); }

// latency plan: level by level; inside a level every instruction reads its operands before ANY
// instruction of the level writes (the kernel runs them concurrently), which exposes slot hazards.
// out5: n_levels, n_slots, n_instrs, max level width, status
int64_t sim_eval_latency(SimGraph* s, const uint8_t* inputs, uint8_t* witness, uint64_t* out5) {
  LatencyPlan lp;
  try { lp = compile_latency_plan(s->g, 7264); } catch (const std::exception&) { return -2; }
  std::vector<fe> slots(lp.n_slots, fe_zero());
  uint32_t st = 0;
  size_t pos = 0;
  struct W { uint32_t dst; fe v; };
  for (uint32_t cnt : lp.level_count) {
    std::vector<W> writes;
    for (size_t k = pos; k < pos + cnt; k++) {
      const Instr& ins = lp.code[k];
      uint32_t op = ins.x & 0xFF, dst = ins.x >> 16;
      if (op == OP_NOP) continue;
      auto load = [&](uint32_t idx, bool is_const, fe* o) { if (is_const) { if (idx >= lp.consts.size()) return false; *o = to_fe(lp.consts[idx]); } else { if (idx >= lp.n_slots) return false; *o = slots[idx]; } return true; };
      fe A = fe_zero(), B = fe_zero(), C = fe_zero(), R;
      if (op == OP_INPUT) { if (ins.y >= lp.n_inputs) return -1; fe v; memcpy(v.l, inputs + 32 * (size_t)ins.y, 32); R = fe_reduce256(v); }
      else {
        if (!load(ins.y, ins.x & F_A_CONST, &A)) return -1;
        if (op == OP_OUT) { if (ins.w >= lp.n_witness) return -1; memcpy(witness + 32 * (size_t)ins.w, A.l, 32); continue; }
        if (op_has_b(op) && !load(ins.z, ins.x & F_B_CONST, &B)) return -1;
        if (op == OP_TERN && !load(ins.w, ins.x & F_C_CONST, &C)) return -1;
        R = alu_exec(op, A, B, C, st);
      }
      if (dst != NO_DST) { if (dst >= lp.n_slots) return -1; writes.push_back({dst, R}); }
      if (ins.x & F_OUT) { if (ins.w >= lp.n_witness) return -1; memcpy(witness + 32 * (size_t)ins.w, R.l, 32); }
    }
    for (const W& w : writes) slots[w.dst] = w.v;
    pos += cnt;
  }
  out5[0] = lp.level_count.size(); out5[1] = lp.n_slots; out5[2] = lp.code.size(); out5[3] = lp.max_level_width; out5[4] = st;
  return 0;
}
}
