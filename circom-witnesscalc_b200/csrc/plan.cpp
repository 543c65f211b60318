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

  auto operands = [&](const Node& nd, uint32_t* ops) {
    int n = 0;
    if (nd.kind >= N_UNO) ops[n++] = nd.a;
    if (nd.kind >= N_DUO) ops[n++] = nd.b;
    if (nd.kind == N_TRES) ops[n++] = nd.c;
    return n;
  };

  // schedule: file order (any topological order is valid, graph.rs:343-356), except that an
  // independent multiplication found within a small window is pulled up next to a multiplication so
  // that the two can be issued as one pair (instruction-level parallelism inside a thread).
  std::vector<uint32_t> order; order.reserve(N);
  std::vector<int32_t> partner(N, -1);       // for the first node of a pair: the second one
  {
    std::vector<uint8_t> scheduled(N, 0);
    auto is_mul = [&](size_t i) { const Node& nd = g.nodes[i]; return needed[i] && nd.kind == N_DUO && nd.op == OP_MUL && const_of[i] < 0; };
    auto ready = [&](size_t j, size_t first) {
      const Node& nd = g.nodes[j];
      for (uint32_t x : {nd.a, nd.b}) {
        if (x == first) return false;
        if (const_of[x] < 0 && !scheduled[x]) return false;
      }
      return true;
    };
    for (size_t i = 0; i < N; i++) {
      if (!needed[i] || scheduled[i]) continue;
      scheduled[i] = 1; order.push_back((uint32_t)i);
      if (!opt.pair_muls || !is_mul(i)) continue;
      size_t end = std::min(N, i + 1 + (size_t)opt.pair_window);
      for (size_t j = i + 1; j < end; j++) {
        if (scheduled[j] || !is_mul(j) || !ready(j, i)) continue;
        scheduled[j] = 1; order.push_back((uint32_t)j); partner[i] = (int32_t)j;
        plan.stats.mul_pairs++;
        break;
      }
    }
  }
  std::vector<uint32_t> pos(N, 0);
  for (size_t p = 0; p < order.size(); p++) pos[order[p]] = (uint32_t)p;

  // use lists of non-constant values as schedule positions of their consumers (CSR, ascending)
  std::vector<uint32_t> use_start(N + 1, 0);
  for (uint32_t i : order) {
    uint32_t ops[3]; int n = operands(g.nodes[i], ops);
    for (int k = 0; k < n; k++) if (const_of[ops[k]] < 0) use_start[ops[k] + 1]++;
  }
  for (size_t i = 0; i < N; i++) use_start[i + 1] += use_start[i];
  std::vector<uint32_t> use_list(use_start[N]);
  {
    std::vector<uint32_t> fill(use_start.begin(), use_start.end() - 1);
    for (uint32_t i : order) {
      uint32_t ops[3]; int n = operands(g.nodes[i], ops);
      for (int k = 0; k < n; k++) if (const_of[ops[k]] < 0) use_list[fill[ops[k]]++] = pos[i];
    }
  }

  Allocator al(N, opt.n_regs, use_start, use_list, plan);
  plan.code.reserve(N + N / 4);
  uint32_t live = 0;

  struct Enc { uint32_t op, flags, dst, a, b, w; bool need_reg, has_uses, out_inline; };
  // phase 1 of one node: operands resident (registers in `pinned` stay put); returns encodings of operands
  auto load_operands = [&](uint32_t i, uint32_t* enc, uint32_t& flags, uint32_t* pinned, int& n_pin) {
    const Node& nd = g.nodes[i];
    uint32_t ops[3]; int n_ops = operands(nd, ops);
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
  };
  // phase 2: consume the uses up to schedule position `upto`; operands that die give their register back
  auto retire_operands = [&](uint32_t i, uint32_t upto) {
    const Node& nd = g.nodes[i];
    uint32_t ops[3]; int n_ops = operands(nd, ops);
    for (int k = 0; k < n_ops; k++) {
      uint32_t x = ops[k];
      if (const_of[x] >= 0) continue;
      while (al.use_ptr[x] < use_start[x + 1] && use_list[al.use_ptr[x]] <= upto) al.use_ptr[x]++;
    }
    for (int k = 0; k < n_ops; k++) {
      uint32_t x = ops[k];
      if (const_of[x] >= 0 || al.has_uses(x)) continue;
      if (al.reg_of[x] >= 0 || al.spill_of[x] >= 0) { al.release(x); live--; }
    }
  };

  for (size_t p = 0; p < order.size(); p++) {
    const uint32_t i = order[p];
    const Node& nd = g.nodes[i];
    const uint32_t n_out = out_start[i + 1] - out_start[i];
    const uint32_t* outs = &out_list[out_start[i]];

    if (const_of[i] >= 0) {                 // constants never occupy a register
      for (uint32_t k = 0; k < n_out; k++) { al.emit(make_instr(OP_OUT, F_A_CONST, NO_DST, (uint32_t)const_of[i], 0, outs[k])); plan.stats.outs++; }
      continue;
    }
    if (nd.kind == N_CONST) continue;
    if (nd.kind >= N_UNO) plan.stats.live_ops++;

    const int n_nodes_here = partner[i] >= 0 ? 2 : 1;
    uint32_t ids[2] = {i, partner[i] >= 0 ? (uint32_t)partner[i] : 0u};
    if (n_nodes_here == 2) { plan.stats.live_ops++; p++; }       // the partner is order[p + 1]
    uint32_t enc[2][3] = {{0, 0, 0}, {0, 0, 0}}; uint32_t flags[2] = {0, 0};
    uint32_t pinned[8]; int n_pin = 0;
    for (int s = 0; s < n_nodes_here; s++) load_operands(ids[s], enc[s], flags[s], pinned, n_pin);
    for (int s = 0; s < n_nodes_here; s++) retire_operands(ids[s], pos[ids[n_nodes_here - 1]]);

    uint32_t dsts[2] = {NO_DST, NO_DST}; bool need_reg[2], has_uses[2], out_inline[2];
    uint32_t dpin[2]; int n_dpin = 0;
    for (int s = 0; s < n_nodes_here; s++) {
      const uint32_t id = ids[s];
      const uint32_t no = out_start[id + 1] - out_start[id];
      has_uses[s] = al.has_uses(id);
      out_inline[s] = no >= 1 && g.nodes[id].kind != N_TRES;       // .w is operand c for TernCond
      need_reg[s] = has_uses[s] || no > (out_inline[s] ? 1u : 0u);
      if (need_reg[s]) {
        dsts[s] = al.alloc_reg(dpin, n_dpin); al.bind(id, dsts[s]); dpin[n_dpin++] = dsts[s];
        live++; plan.stats.max_live = std::max(plan.stats.max_live, live);
      }
    }
    if (n_nodes_here == 2 && (plan.code.size() & 31) == 31) al.emit(make_instr(OP_NOP, 0, NO_DST, 0, 0, 0));
    for (int s = 0; s < n_nodes_here; s++) {
      const uint32_t id = ids[s];
      const Node& n2 = g.nodes[id];
      const uint32_t* o2 = &out_list[out_start[id]];
      uint32_t op;
      if (n2.kind == N_INPUT) { op = OP_INPUT; enc[s][0] = n2.a; }
      else if (n2.kind == N_UNO) op = OP_NEG + n2.op;
      else if (n2.kind == N_TRES) op = OP_TERN;
      else op = (n2.op == OP_MUL && n2.a == n2.b && !(flags[s] & F_A_CONST)) ? (uint32_t)OP_SQR : n2.op;
      uint32_t w = n2.kind == N_TRES ? enc[s][2] : (out_inline[s] ? o2[0] : 0);
      if (out_inline[s]) { flags[s] |= F_OUT; plan.stats.outs++; }
      if (n_nodes_here == 2 && s == 0) flags[s] |= F_PAIR;
      al.emit(make_instr(op, flags[s], dsts[s], enc[s][0], enc[s][1], w));
    }
    for (int s = 0; s < n_nodes_here; s++) {
      const uint32_t id = ids[s];
      const uint32_t no = out_start[id + 1] - out_start[id];
      const uint32_t* o2 = &out_list[out_start[id]];
      for (uint32_t k = out_inline[s] ? 1u : 0u; k < no; k++) { al.emit(make_instr(OP_OUT, 0, NO_DST, dsts[s], 0, o2[k])); plan.stats.outs++; }
      if (need_reg[s] && !has_uses[s]) { al.release(id); live--; }
    }
  }
  plan.n_spill = al.n_spill;
  plan.stats.instrs = plan.code.size();
  if (plan.consts.empty()) plan.consts.push_back(u256_from_u64(0));
  return plan;
}

}  // namespace gw

namespace gw {

LatencyPlan compile_latency_plan(const Graph& g, uint32_t max_slots) {
  const size_t N = g.nodes.size();
  LatencyPlan lp;
  lp.n_inputs = g.inputs_size;
  lp.n_witness = (uint32_t)g.witness_signals.size();

  std::vector<uint8_t> needed(N, 0);
  for (uint32_t s : g.witness_signals) needed[s] = 1;
  for (size_t i = N; i-- > 0;) {
    if (!needed[i]) continue;
    const Node& nd = g.nodes[i];
    if (nd.kind >= N_UNO) needed[nd.a] = 1;
    if (nd.kind >= N_DUO) needed[nd.b] = 1;
    if (nd.kind == N_TRES) needed[nd.c] = 1;
  }
  std::vector<int32_t> const_of(N, -1);
  std::map<U256, uint32_t> cix;
  auto intern = [&](const U256& v) {
    auto it = cix.find(v);
    if (it == cix.end()) { it = cix.emplace(v, (uint32_t)lp.consts.size()).first; lp.consts.push_back(v); }
    return (int32_t)it->second;
  };
  for (size_t i = 0; i < N; i++) {
    if (!needed[i]) continue;
    const Node& nd = g.nodes[i];
    if (nd.kind == N_CONST) const_of[i] = intern(g.constants.at(nd.a));
    else if (nd.kind == N_INPUT && nd.a == 0) const_of[i] = intern(u256_from_u64(1));
    else if (nd.kind == N_INPUT && nd.a >= g.inputs_size) throw Error("plan: input index out of range");
  }
  auto operands = [&](const Node& nd, uint32_t* ops) {
    int n = 0;
    if (nd.kind >= N_UNO) ops[n++] = nd.a;
    if (nd.kind >= N_DUO) ops[n++] = nd.b;
    if (nd.kind == N_TRES) ops[n++] = nd.c;
    return n;
  };
  // dependency levels (inputs are level 0), last level at which every value is read
  std::vector<uint32_t> level(N, 0), last_use(N, 0);
  uint32_t n_levels = 1;
  for (size_t i = 0; i < N; i++) {
    if (!needed[i] || const_of[i] >= 0) continue;
    const Node& nd = g.nodes[i];
    uint32_t ops[3]; int n = operands(nd, ops);
    uint32_t lv = 0;
    for (int k = 0; k < n; k++) if (const_of[ops[k]] < 0) lv = std::max(lv, level[ops[k]] + 1);
    if (nd.kind >= N_UNO && lv == 0) lv = 1;            // constant-only operands: still after the input level
    level[i] = lv;
    for (int k = 0; k < n; k++) if (const_of[ops[k]] < 0) last_use[ops[k]] = std::max(last_use[ops[k]], lv);
    n_levels = std::max(n_levels, lv + 1);
  }
  // witness positions per node
  std::vector<uint32_t> out_start(N + 1, 0), out_list(g.witness_signals.size());
  for (uint32_t s : g.witness_signals) out_start[s + 1]++;
  for (size_t i = 0; i < N; i++) out_start[i + 1] += out_start[i];
  {
    std::vector<uint32_t> fill(out_start.begin(), out_start.end() - 1);
    for (uint32_t j = 0; j < g.witness_signals.size(); j++) out_list[fill[g.witness_signals[j]]++] = j;
  }
  // extra OUT instructions (second and later witness positions, TernCond results) run one level later
  for (size_t i = 0; i < N; i++) {
    if (!needed[i] || const_of[i] >= 0) continue;
    uint32_t n_out = out_start[i + 1] - out_start[i];
    bool inline_out = n_out >= 1 && g.nodes[i].kind != N_TRES;
    if (n_out > (inline_out ? 1u : 0u)) { last_use[i] = std::max(last_use[i], level[i] + 1); n_levels = std::max(n_levels, level[i] + 2); }
  }
  // bucket nodes by level, inside a level by opcode (keeps warps uniform)
  std::vector<std::vector<uint32_t>> by_level(n_levels);
  for (size_t i = 0; i < N; i++) if (needed[i] && const_of[i] < 0) by_level[level[i]].push_back((uint32_t)i);
  auto opkey = [&](uint32_t i) { const Node& nd = g.nodes[i]; return (uint32_t)nd.kind * 64u + nd.op; };
  for (auto& v : by_level) std::stable_sort(v.begin(), v.end(), [&](uint32_t a, uint32_t b) { return opkey(a) < opkey(b); });

  std::vector<int32_t> slot_of(N, -1);
  std::vector<uint32_t> free_slots;
  std::vector<std::vector<uint32_t>> dying(n_levels + 1);       // values whose last read happens at this level
  uint32_t n_slots = 0;
  std::vector<Instr> pending_outs;                              // OUT instructions for the next level
  for (uint32_t L = 0; L < n_levels; L++) {
    uint32_t count = 0;
    for (const Instr& in : pending_outs) { lp.code.push_back(in); count++; }
    pending_outs.clear();
    if (L == 0) {
      for (size_t i = 0; i < N; i++) {
        if (!needed[i] || const_of[i] < 0) continue;
        for (uint32_t k = out_start[i]; k < out_start[i + 1]; k++) { lp.code.push_back(make_instr(OP_OUT, F_A_CONST, NO_DST, (uint32_t)const_of[i], 0, out_list[k])); count++; }
      }
    }
    for (uint32_t i : by_level[L]) {
      const Node& nd = g.nodes[i];
      uint32_t ops[3]; int n = operands(nd, ops);
      uint32_t enc[3] = {0, 0, 0}, flags = 0;
      for (int k = 0; k < n; k++) {
        if (const_of[ops[k]] >= 0) { enc[k] = (uint32_t)const_of[ops[k]]; flags |= (F_A_CONST << k); }
        else { if (slot_of[ops[k]] < 0) throw Error("latency plan: operand without a slot"); enc[k] = (uint32_t)slot_of[ops[k]]; }
      }
      const uint32_t n_out = out_start[i + 1] - out_start[i];
      const uint32_t* outs = &out_list[out_start[i]];
      const bool inline_out = n_out >= 1 && nd.kind != N_TRES;
      uint32_t dst = NO_DST;
      if (last_use[i] > L) {
        if (!free_slots.empty()) { dst = free_slots.back(); free_slots.pop_back(); }
        else { dst = n_slots++; if (n_slots > max_slots) throw Error("latency plan: graph is too wide for the shared-memory value file"); }
        slot_of[i] = (int32_t)dst;
        dying[last_use[i]].push_back(i);
      }
      uint32_t op;
      if (nd.kind == N_INPUT) { op = OP_INPUT; enc[0] = nd.a; }
      else if (nd.kind == N_UNO) op = OP_NEG + nd.op;
      else if (nd.kind == N_TRES) op = OP_TERN;
      else op = (nd.op == OP_MUL && nd.a == nd.b && !(flags & F_A_CONST)) ? (uint32_t)OP_SQR : nd.op;
      uint32_t w = nd.kind == N_TRES ? enc[2] : (inline_out ? outs[0] : 0);
      if (inline_out) flags |= F_OUT;
      lp.code.push_back(make_instr(op, flags, dst, enc[0], enc[1], w));
      count++;
      for (uint32_t k = inline_out ? 1u : 0u; k < n_out; k++) pending_outs.push_back(make_instr(OP_OUT, 0, NO_DST, dst, 0, outs[k]));
    }
    // slots read for the last time in this level become reusable from the next level on
    for (uint32_t v : dying[L]) { free_slots.push_back((uint32_t)slot_of[v]); }
    lp.level_count.push_back(count);
    lp.max_level_width = std::max(lp.max_level_width, count);
  }
  if (!pending_outs.empty()) {
    for (const Instr& in : pending_outs) lp.code.push_back(in);
    lp.level_count.push_back((uint32_t)pending_outs.size());
  }
  lp.n_slots = std::max(n_slots, 1u);
  if (lp.consts.empty()) lp.consts.push_back(u256_from_u64(0));
  return lp;
}

}  // namespace gw
