#include "plan.hpp"

#include <string.h>

#include <algorithm>
#include <map>

namespace gw {

namespace {

struct Allocator {
  const std::vector<uint32_t>& use_start;
  const std::vector<uint32_t>& use_list;
  std::vector<uint32_t> use_ptr;           // per value: index of its next unconsumed use
  std::vector<int32_t> reg_of, spill_of;   // per value
  std::vector<int32_t> reg_val;            // per register: value or -1
  std::vector<uint32_t> free_regs, free_spill;
  uint32_t n_spill = 0;
  Plan& plan;

  Allocator(size_t n, uint32_t n_regs, const std::vector<uint32_t>& us, const std::vector<uint32_t>& ul, Plan& p)
      : use_start(us), use_list(ul), use_ptr(n), reg_of(n, -1), spill_of(n, -1), reg_val(n_regs, -1), plan(p) {
    for (size_t i = 0; i < n; i++) use_ptr[i] = use_start[i];
    for (uint32_t r = n_regs; r-- > 0;) free_regs.push_back(r);
  }
  uint32_t next_use(uint32_t v) const { return use_ptr[v] < use_start[v + 1] ? use_list[use_ptr[v]] : 0xFFFFFFFFu; }
  bool has_uses(uint32_t v) const { return use_ptr[v] < use_start[v + 1]; }
  void emit(const Instr& in) {
    plan.code.push_back(in);
    plan.stats.op_count[in.x & 0x3F]++;
  }
  // a free register; `pin` = registers that must stay resident for the instruction being built
  uint32_t alloc_reg(const uint32_t* pin, int n_pin) {
    if (!free_regs.empty()) { uint32_t r = free_regs.back(); free_regs.pop_back(); return r; }
    int best = -1; uint32_t best_use = 0;
    for (uint32_t r = 0; r < reg_val.size(); r++) {
      bool pinned = false;
      for (int k = 0; k < n_pin; k++) pinned |= (pin[k] == r);
      if (pinned) continue;
      uint32_t nu = next_use((uint32_t)reg_val[r]);
      if (best < 0 || nu > best_use) { best = (int)r; best_use = nu; }
    }
    if (best < 0) throw Error("plan: register file too small for one instruction");
    uint32_t v = (uint32_t)reg_val[best];
    if (spill_of[v] < 0) {                 // first eviction of this value: write it out once (SSA: never changes)
      uint32_t s;
      if (!free_spill.empty()) { s = free_spill.back(); free_spill.pop_back(); }
      else s = n_spill++;
      spill_of[v] = (int32_t)s;
      emit(make_instr(OP_SPILL_ST, 0, NO_DST, (uint32_t)best, s, 0));
      plan.stats.spill_st++;
    }
    reg_of[v] = -1;
    reg_val[best] = -1;
    return (uint32_t)best;
  }
  void bind(uint32_t v, uint32_t r) { reg_of[v] = (int32_t)r; reg_val[r] = (int32_t)v; }
  void release(uint32_t v) {               // value is dead
    if (reg_of[v] >= 0) { reg_val[reg_of[v]] = -1; free_regs.push_back((uint32_t)reg_of[v]); reg_of[v] = -1; }
    if (spill_of[v] >= 0) { free_spill.push_back((uint32_t)spill_of[v]); spill_of[v] = -1; }
  }
};

}  // namespace

Plan compile_plan(const Graph& g, const PlanOptions& opt) {
  const size_t N = g.nodes.size();
  if (opt.n_regs < 4 || opt.n_regs > 4096) throw Error("plan: n_regs out of range");
  Plan plan;
  plan.n_regs = opt.n_regs;
  plan.n_inputs = g.inputs_size;
  plan.n_witness = (uint32_t)g.witness_signals.size();
  plan.stats.graph_nodes = N;
  plan.stats.graph_ops = g.n_ops();

  // liveness from the witness backwards (dead nodes are unobservable: evaluate() is pure, graph.rs:367)
  std::vector<uint8_t> needed(N, 0);
  for (uint32_t s : g.witness_signals) needed[s] = 1;
  for (size_t i = N; i-- > 0;) {
    if (!needed[i]) continue;
    const Node& nd = g.nodes[i];
    if (nd.kind >= N_UNO) needed[nd.a] = 1;
    if (nd.kind >= N_DUO) needed[nd.b] = 1;
    if (nd.kind == N_TRES) needed[nd.c] = 1;
  }

  // constants: N_CONST nodes, and Input(0) which get_inputs_buffer forces to 1 (lib.rs:177-181)
  std::vector<int32_t> const_of(N, -1);
  std::map<U256, uint32_t> cix;
  auto intern = [&](const U256& v) {
    auto it = cix.find(v);
    if (it == cix.end()) { it = cix.emplace(v, (uint32_t)plan.consts.size()).first; plan.consts.push_back(v); }
    return (int32_t)it->second;
  };
  for (size_t i = 0; i < N; i++) {
    if (!needed[i]) continue;
    const Node& nd = g.nodes[i];
    if (nd.kind == N_CONST) const_of[i] = intern(g.constants.at(nd.a));
    else if (nd.kind == N_INPUT && nd.a == 0) const_of[i] = intern(u256_from_u64(1));
    else if (nd.kind == N_INPUT && nd.a >= g.inputs_size) throw Error("plan: input index out of range");
  }

  // witness positions per node (CSR)
  std::vector<uint32_t> out_start(N + 1, 0), out_list(g.witness_signals.size());
  for (uint32_t s : g.witness_signals) out_start[s + 1]++;
  for (size_t i = 0; i < N; i++) out_start[i + 1] += out_start[i];
  {
    std::vector<uint32_t> fill(out_start.begin(), out_start.end() - 1);
    for (uint32_t j = 0; j < g.witness_signals.size(); j++) out_list[fill[g.witness_signals[j]]++] = j;
  }

  // use lists of non-constant values, in instruction order (CSR)
  auto operands = [&](const Node& nd, uint32_t* ops) {
    int n = 0;
    if (nd.kind >= N_UNO) ops[n++] = nd.a;
    if (nd.kind >= N_DUO) ops[n++] = nd.b;
    if (nd.kind == N_TRES) ops[n++] = nd.c;
    return n;
  };
  std::vector<uint32_t> use_start(N + 1, 0);
  for (size_t i = 0; i < N; i++) {
    if (!needed[i]) continue;
    uint32_t ops[3]; int n = operands(g.nodes[i], ops);
    for (int k = 0; k < n; k++) if (const_of[ops[k]] < 0) use_start[ops[k] + 1]++;
  }
  for (size_t i = 0; i < N; i++) use_start[i + 1] += use_start[i];
  std::vector<uint32_t> use_list(use_start[N]);
  {
    std::vector<uint32_t> fill(use_start.begin(), use_start.end() - 1);
    for (size_t i = 0; i < N; i++) {
      if (!needed[i]) continue;
      uint32_t ops[3]; int n = operands(g.nodes[i], ops);
      for (int k = 0; k < n; k++) if (const_of[ops[k]] < 0) use_list[fill[ops[k]]++] = (uint32_t)i;
    }
  }

  Allocator al(N, opt.n_regs, use_start, use_list, plan);
  plan.code.reserve(N + N / 4);
  uint32_t live = 0;

  for (size_t i = 0; i < N; i++) {
    if (!needed[i]) continue;
    const Node& nd = g.nodes[i];
    const uint32_t n_out = out_start[i + 1] - out_start[i];
    const uint32_t* outs = &out_list[out_start[i]];

    if (const_of[i] >= 0) {                 // constants never occupy a register
      for (uint32_t k = 0; k < n_out; k++) { al.emit(make_instr(OP_OUT, F_A_CONST, NO_DST, (uint32_t)const_of[i], 0, outs[k])); plan.stats.outs++; }
      continue;
    }
    if (nd.kind == N_CONST) continue;

    uint32_t ops[3]; int n_ops = operands(nd, ops);
    uint32_t enc[3] = {0, 0, 0}; uint32_t flags = 0;
    uint32_t pinned[3]; int n_pin = 0;
    if (nd.kind >= N_UNO) plan.stats.live_ops++;
    // make the operands resident
    for (int k = 0; k < n_ops; k++) {
      uint32_t x = ops[k];
      if (const_of[x] >= 0) { enc[k] = (uint32_t)const_of[x]; flags |= (F_A_CONST << k); continue; }
      if (al.reg_of[x] < 0) {
        if (al.spill_of[x] < 0) throw Error("plan: operand neither resident nor spilled");
        uint32_t r = al.alloc_reg(pinned, n_pin);
        al.emit(make_instr(OP_SPILL_LD, 0, r, (uint32_t)al.spill_of[x], 0, 0));
        plan.stats.spill_ld++;
        al.bind(x, r);
      }
      enc[k] = (uint32_t)al.reg_of[x];
      pinned[n_pin++] = enc[k];
    }
    // consume the uses; operands that die here give their register back before dst is chosen
    for (int k = 0; k < n_ops; k++) {
      uint32_t x = ops[k];
      if (const_of[x] >= 0) continue;
      while (al.use_ptr[x] < use_start[x + 1] && use_list[al.use_ptr[x]] <= i) al.use_ptr[x]++;
    }
    for (int k = 0; k < n_ops; k++) {
      uint32_t x = ops[k];
      if (const_of[x] >= 0 || al.has_uses(x)) continue;
      if (al.reg_of[x] >= 0 || al.spill_of[x] >= 0) { al.release(x); live--; }
    }
    const bool has_uses = al.has_uses((uint32_t)i);
    const bool out_inline = n_out >= 1 && nd.kind != N_TRES;       // .w is operand c for TernCond
    const bool need_reg = has_uses || n_out > (out_inline ? 1u : 0u);
    uint32_t dst = NO_DST;
    if (need_reg) { dst = al.alloc_reg(nullptr, 0); al.bind((uint32_t)i, dst); live++; plan.stats.max_live = std::max(plan.stats.max_live, live); }
    if (!need_reg && n_out == 0) continue;   // cannot happen for needed nodes, kept for safety

    uint32_t op;
    if (nd.kind == N_INPUT) { op = OP_INPUT; enc[0] = nd.a; }
    else if (nd.kind == N_UNO) op = OP_NEG + nd.op;
    else if (nd.kind == N_TRES) op = OP_TERN;
    else op = (nd.op == OP_MUL && nd.a == nd.b && !(flags & F_A_CONST)) ? (uint32_t)OP_SQR : nd.op;
    uint32_t w = nd.kind == N_TRES ? enc[2] : (out_inline ? outs[0] : 0);
    if (out_inline) { flags |= F_OUT; plan.stats.outs++; }
    al.emit(make_instr(op, flags, dst, enc[0], enc[1], w));
    for (uint32_t k = out_inline ? 1u : 0u; k < n_out; k++) { al.emit(make_instr(OP_OUT, 0, NO_DST, dst, 0, outs[k])); plan.stats.outs++; }
    if (need_reg && !has_uses) { al.release((uint32_t)i); live--; }
  }
  plan.n_spill = al.n_spill;
  plan.stats.instrs = plan.code.size();
  if (plan.consts.empty()) plan.consts.push_back(u256_from_u64(0));
  return plan;
}

}  // namespace gw
