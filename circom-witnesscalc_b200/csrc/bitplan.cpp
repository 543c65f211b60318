#include "bitplan.hpp"

#include <string.h>

#include <algorithm>
#include <array>
#include <map>

#include "alu.cuh"

namespace gw {

namespace {

inline int n_operands(const Node& nd) { return nd.kind == N_TRES ? 3 : nd.kind == N_DUO ? 2 : nd.kind == N_UNO ? 1 : 0; }
inline uint32_t operand(const Node& nd, int k) { return k == 0 ? nd.a : k == 1 ? nd.b : nd.c; }
inline fe to_fe(const U256& v) { fe r; memcpy(r.l, v.l, 32); return r; }
inline U256 to_u256(const fe& v) { U256 r; memcpy(r.l, v.l, 32); return r; }

// Exact upper bounds of bit-vector values: 256-bit unsigned integers.  A bit-vector value must stay below M, so that
// the integer the planes hold IS the field element the reference computes (no reduction ever happens).
struct Big {
  U256 v;
  Big() { memset(v.l, 0, 32); }
  explicit Big(uint64_t x) { memset(v.l, 0, 32); v.l[0] = (uint32_t)x; v.l[1] = (uint32_t)(x >> 32); }
  explicit Big(const U256& x) : v(x) {}
  bool is_zero() const { for (int k = 0; k < 8; k++) if (v.l[k]) return false; return true; }
  bool is_one() const { if (v.l[0] != 1) return false; for (int k = 1; k < 8; k++) if (v.l[k]) return false; return true; }
  int bits() const { for (int k = 7; k >= 0; k--) if (v.l[k]) { int n = 0; uint32_t x = v.l[k]; while (x) { n++; x >>= 1; } return 32 * k + n; } return 0; }
  bool bit(int p) const { return p < 256 && ((v.l[p >> 5] >> (p & 31)) & 1u); }
  bool operator<(const Big& o) const { return v < o.v; }
  bool below_modulus() const { return v < BN254_M; }
  // false on overflow past 256 bits
  static bool add(const Big& a, const Big& b, Big* r) {
    uint64_t c = 0;
    for (int k = 0; k < 8; k++) { c += (uint64_t)a.v.l[k] + b.v.l[k]; r->v.l[k] = (uint32_t)c; c >>= 32; }
    return c == 0;
  }
  static bool shl(const Big& a, uint32_t k, Big* r) {
    if (a.is_zero()) { *r = Big(); return true; }
    if ((uint32_t)a.bits() + k > 256) return false;
    *r = Big();
    for (int p = 0; p < a.bits(); p++) if (a.bit(p)) r->v.l[(p + (int)k) >> 5] |= 1u << ((p + (int)k) & 31);
    return true;
  }
  static Big shr(const Big& a, uint32_t k) {
    Big r;
    for (int p = (int)k; p < a.bits(); p++) if (a.bit(p)) r.v.l[(p - (int)k) >> 5] |= 1u << ((p - (int)k) & 31);
    return r;
  }
  static bool mul(const Big& a, const Big& b, Big* r) {
    if (a.bits() + b.bits() > 256) return false;            // conservative: the product certainly fits below
    uint32_t P[16];
    u256_mul_wide(P, a.v.l, b.v.l);
    for (int k = 8; k < 16; k++) if (P[k]) return false;
    memcpy(r->v.l, P, 32);
    return true;
  }
  static Big ones(int n) { Big r; for (int p = 0; p < n && p < 256; p++) r.v.l[p >> 5] |= 1u << (p & 31); return r; }
};

// ---- truth tables over <= 6 variables held in 64 bits: bit idx = f(x_0 = idx & 1, x_1 = idx >> 1 & 1, ..) --------------
inline uint64_t tt_cofactor(uint64_t t, int n, int k, int val) {     // fix variable k: table over the remaining n - 1 variables
  uint64_t r = 0;
  for (int idx = 0; idx < (1 << (n - 1)); idx++) {
    const int lo = idx & ((1 << k) - 1), hi = idx >> k;
    const int full = lo | (val << k) | (hi << (k + 1));
    r |= ((t >> full) & 1ull) << idx;
  }
  return r;
}
inline bool tt_depends(uint64_t t, int n, int k) { return tt_cofactor(t, n, k, 0) != tt_cofactor(t, n, k, 1); }
// identify variable k with variable j (j < k): keep the assignments where they agree, drop variable k
inline uint64_t tt_merge(uint64_t t, int n, int j, int k) {
  uint64_t r = 0;
  for (int idx = 0; idx < (1 << (n - 1)); idx++) {
    const int lo = idx & ((1 << k) - 1), hi = idx >> k;
    const int vj = (idx >> j) & 1;
    const int full = lo | (vj << k) | (hi << (k + 1));
    r |= ((t >> full) & 1ull) << idx;
  }
  return r;
}

// ---- the LUT DAG ---------------------------------------------------------------------------------------------------
struct BGraph {
  enum { K_ZERO = 0, K_ONE = 1, K_INPUT = 2, K_LUT = 3 };
  struct BN { uint8_t kind, n, lut; uint32_t in[3]; };       // K_INPUT: in[0] = input index; K_LUT: n inputs, table of 2^n bits
  std::vector<BN> nodes;
  std::map<std::array<uint32_t, 4>, uint32_t> cse;
  std::map<std::pair<uint32_t, uint32_t>, uint32_t> input_ids;   // (input index, bit or BIT_CONTRACT) -> node
  BGraph() {
    nodes.push_back(BN{K_ZERO, 0, 0, {0, 0, 0}});
    nodes.push_back(BN{K_ONE, 0, 0, {0, 0, 0}});
  }
  uint32_t input(uint32_t idx, uint32_t bit) {
    auto key = std::make_pair(idx, bit);
    auto it = input_ids.find(key);
    if (it != input_ids.end()) return it->second;
    nodes.push_back(BN{K_INPUT, 0, 0, {idx, bit, 0}});
    return input_ids[key] = (uint32_t)nodes.size() - 1;
  }
  // f(leaves) given by `table` (2^n bits, n <= 6): constants, repeated and irrelevant leaves are removed, what is
  // left becomes one LUT (n <= 3) or a Shannon tree of multiplexers over the last leaf
  uint32_t func(std::vector<uint32_t> lv, uint64_t t) {
    int n = (int)lv.size();
    for (int k = n - 1; k >= 0; k--) {
      bool drop = false;
      if (lv[k] <= 1) { t = tt_cofactor(t, n, k, (int)lv[k]); drop = true; }
      else {
        int j = -1;
        for (int q = 0; q < k; q++) if (lv[q] == lv[k]) j = q;
        if (j >= 0) { t = tt_merge(t, n, j, k); drop = true; }
        else if (!tt_depends(t, n, k)) { t = tt_cofactor(t, n, k, 0); drop = true; }
      }
      if (drop) { lv.erase(lv.begin() + k); n--; }
    }
    if (n < 64) t &= (n >= 6) ? ~0ull : ((1ull << (1 << n)) - 1);
    if (n == 0) return (uint32_t)(t & 1);
    if (n == 1 && t == 2) return lv[0];
    if (n <= 3) {
      // canonical order of the inputs: ascending ids
      int perm[3] = {0, 1, 2};
      for (int x = 1; x < n; x++) for (int y = x; y > 0 && lv[(size_t)perm[y]] < lv[(size_t)perm[y - 1]]; y--) std::swap(perm[y], perm[y - 1]);
      uint32_t nt = 0;
      for (int idx = 0; idx < (1 << n); idx++) {
        int old = 0;
        for (int q = 0; q < n; q++) if ((idx >> q) & 1) old |= 1 << perm[q];
        nt |= (uint32_t)((t >> old) & 1ull) << idx;
      }
      std::array<uint32_t, 4> key = {nt | ((uint32_t)n << 8), 0, 0, 0};
      for (int q = 0; q < n; q++) key[1 + q] = lv[perm[q]];
      auto it = cse.find(key);
      if (it != cse.end()) return it->second;
      BN b{K_LUT, (uint8_t)n, (uint8_t)nt, {0, 0, 0}};
      for (int q = 0; q < n; q++) b.in[q] = key[1 + q];
      nodes.push_back(b);
      return cse[key] = (uint32_t)nodes.size() - 1;
    }
    std::vector<uint32_t> rest(lv.begin(), lv.end() - 1);
    const uint32_t f0 = func(rest, tt_cofactor(t, n, n - 1, 0)), f1 = func(rest, tt_cofactor(t, n, n - 1, 1));
    return func({lv[n - 1], f1, f0}, 0xD8);                  // s ? f1 : f0
  }
  uint32_t lut2(uint32_t t4, uint32_t a, uint32_t b) { return func({a, b}, t4); }
  uint32_t lut3(uint32_t t8, uint32_t a, uint32_t b, uint32_t c) { return func({a, b, c}, t8); }
};

struct TT { std::vector<uint32_t> sup; std::vector<U256> tab; };
struct BV { std::vector<std::vector<uint32_t>> cols; Big hi; bool compressed = true; };
enum { V_NONE = 0, V_CONST = 1, V_TT = 2, V_BV = 3 };
struct Val { uint8_t k = V_NONE; int32_t plane = -1; int32_t tt = -1, bv = -1; U256 c; };

struct Fail { std::string why; };

}  // namespace

BitPlan compile_bit_plan(const Graph& g, const BitPlanOptions& opt) {
  BitPlan bp;
  bp.n_inputs = g.inputs_size;
  bp.n_witness = (uint32_t)g.witness_signals.size();
  const size_t N = g.nodes.size();
  const uint32_t K = std::min<uint32_t>(std::max<uint32_t>(opt.max_support, 3), 6);
  // liveness and use counts
  std::vector<uint8_t> needed(N, 0);
  std::vector<uint32_t> uses(N, 0);
  for (uint32_t s : g.witness_signals) { needed[s] = 1; uses[s]++; }
  for (size_t i = N; i-- > 0;) {
    if (!needed[i]) continue;
    const Node& nd = g.nodes[i];
    for (int k = 0; k < n_operands(nd); k++) { needed[operand(nd, k)] = 1; uses[operand(nd, k)]++; }
  }
  bool any_input = false;
  for (size_t i = 0; i < N; i++) if (needed[i] && g.nodes[i].kind == N_INPUT && g.nodes[i].a != 0) any_input = true;
  // An input whose every reader is Shr(input, constant) (Num2Bits: (in >> i) & 1) is a FIELD input: any value is fine, its
  // planes are the bits of the value reduced mod M.  Every other input is under the bit contract.
  std::vector<uint8_t> field_input(N, 0);
  for (size_t i = 0; i < N; i++) field_input[i] = needed[i] && g.nodes[i].kind == N_INPUT && g.nodes[i].a != 0;
  {
    bool any_shift = false;
    std::vector<uint8_t> shifted(N, 0);
    for (size_t i = 0; i < N; i++) {
      if (!needed[i]) continue;
      const Node& nd = g.nodes[i];
      for (int k = 0; k < n_operands(nd); k++) {
        const uint32_t o = operand(nd, k);
        if (!field_input[o]) continue;
        if (nd.kind == N_DUO && nd.op == OP_SHR && k == 0 && g.nodes[nd.b].kind == N_CONST) { shifted[o] = 1; any_shift = true; }
        else field_input[o] = 0;
      }
    }
    (void)any_shift; (void)shifted;                          // an input nobody reads (a witness signal only) is a field input too
  }
  if (!any_input || bp.n_witness == 0) { bp.reason = "no live inputs or empty witness"; return bp; }

  BGraph bg;
  std::vector<Val> val(N);
  std::vector<TT> tts;
  std::vector<BV> bvs;
  const U256 ZERO = u256_from_u64(0), ONE = u256_from_u64(1);
  auto is_bit_const = [&](const U256& c) { return c == ZERO || c == ONE; };

  auto set_const = [&](size_t i, const U256& c) { val[i] = Val(); val[i].k = V_CONST; val[i].c = c; };
  auto set_leaf = [&](size_t i, uint32_t plane) {          // a bit-valued node known only by its plane
    if (plane <= 1) { set_const(i, plane ? ONE : ZERO); return; }
    TT t; t.sup = {(uint32_t)i}; t.tab = {ZERO, ONE};
    tts.push_back(std::move(t));
    val[i] = Val(); val[i].k = V_TT; val[i].tt = (int32_t)tts.size() - 1; val[i].plane = (int32_t)plane;
  };
  // compress a bit heap to one plane per column with full adders (carry-save: ~ one adder per surplus bit)
  auto compress = [&](BV& b) {
    if (b.compressed) return;
    const size_t width = (size_t)b.hi.bits();
    if (b.cols.size() < width) b.cols.resize(width);
    for (size_t p = 0; p < b.cols.size(); p++) {
      std::vector<uint32_t> q;
      uint32_t ones = 0;
      for (uint32_t x : b.cols[p]) { if (x == 1) ones++; else if (x != 0) q.push_back(x); }
      // constant ones of a column: pairs carry into the next column
      if (ones >= 2 && p + 1 < b.cols.size()) for (uint32_t k = 0; k < ones / 2; k++) b.cols[p + 1].push_back(1);
      if (ones & 1) q.push_back(1);
      size_t head = 0;
      while (q.size() - head >= 3) {
        const uint32_t a = q[head], x = q[head + 1], c = q[head + 2];
        head += 3;
        q.push_back(bg.lut3(0x96, a, x, c));
        if (p + 1 < width) { if (p + 1 >= b.cols.size()) b.cols.resize(p + 2); b.cols[p + 1].push_back(bg.lut3(0xE8, a, x, c)); }
        bp.n_full_adders++;
      }
      if (q.size() - head == 2) {
        const uint32_t a = q[head], x = q[head + 1];
        if (p + 1 < width) { if (p + 1 >= b.cols.size()) b.cols.resize(p + 2); b.cols[p + 1].push_back(bg.lut2(0x8, a, x)); }
        b.cols[p] = {bg.lut2(0x6, a, x)};
      } else if (q.size() - head == 1) b.cols[p] = {q[head]};
      else b.cols[p].clear();
    }
    b.cols.resize(width);
    b.compressed = true;
  };
  auto plane_of_col = [&](const BV& b, size_t p) -> uint32_t { return (p < b.cols.size() && !b.cols[p].empty()) ? b.cols[p][0] : 0u; };
  // BV view of a node (memoised in val[i].bv)
  auto to_bv = [&](uint32_t i) -> int32_t {
    Val& v = val[i];
    if (v.bv >= 0) return v.bv;
    BV b;
    if (v.k == V_CONST) {
      b.hi = Big(v.c);
      if (!b.hi.below_modulus()) return -1;
      for (int p = 0; p < b.hi.bits(); p++) { b.cols.emplace_back(); if (b.hi.bit(p)) b.cols.back().push_back(1); }
    } else if (v.k == V_TT) {
      if (v.plane >= 0) { b.hi = Big(1); b.cols = {{(uint32_t)v.plane}}; }
      else {
        const TT& t = tts[(size_t)v.tt];
        Big mx;
        for (size_t a = 0; a < t.tab.size(); a++) { const Big x(t.tab[a]); if (!x.below_modulus()) return -1; if (mx < x) mx = x; }
        std::vector<uint32_t> leaves;
        for (uint32_t s : t.sup) leaves.push_back((uint32_t)val[s].plane);
        b.hi = mx;
        for (int p = 0; p < mx.bits(); p++) {
          uint64_t bits = 0;
          for (size_t a = 0; a < t.tab.size(); a++) bits |= (uint64_t)Big(t.tab[a]).bit(p) << a;
          b.cols.push_back({bg.func(leaves, bits)});
        }
      }
    } else return -1;
    bvs.push_back(std::move(b));
    return v.bv = (int32_t)bvs.size() - 1;
  };

  try {
    for (size_t i = 0; i < N; i++) {
      if (!needed[i]) continue;
      const Node& nd = g.nodes[i];
      if (nd.kind == N_CONST) { set_const(i, g.constants.at(nd.a)); continue; }
      if (nd.kind == N_INPUT) {
        if (nd.a == 0) { set_const(i, ONE); continue; }          // get_inputs_buffer forces slot 0 to 1 (lib.rs:177-181)
        if (nd.a >= g.inputs_size) throw Fail{"input index out of range"};
        if (field_input[i]) {
          // the 254 bits of the input reduced mod M (Fr::new, graph.rs:376): a bit vector below M, no contract
          BV b;
          b.hi = Big(BN254_M);
          b.hi.v.l[0] -= 1;                                      // M - 1 (M is odd)
          for (uint32_t p = 0; p < 254; p++) b.cols.push_back({bg.input(nd.a, p)});
          bvs.push_back(std::move(b));
          val[i] = Val(); val[i].k = V_BV; val[i].bv = (int32_t)bvs.size() - 1;
          bp.has_field_inputs = true;
          continue;
        }
        set_leaf(i, bg.input(nd.a, BIT_CONTRACT));               // CONTRACT: the input is a bit (checked per input set on the device)
        continue;
      }
      const int no = n_operands(nd);
      const uint32_t opc = nd.kind == N_UNO ? OP_NEG + nd.op : nd.kind == N_TRES ? (uint32_t)OP_TERN : (uint32_t)nd.op;
      uint32_t o[3] = {0, 0, 0};
      for (int k = 0; k < no; k++) o[k] = operand(nd, k);

      // ---- 1. table domain --------------------------------------------------------------------------------------------
      {
        struct View { std::vector<uint32_t> sup; std::vector<U256> tab; };
        View vw[3];
        // A bit-valued operand can be seen through its own table (its support joins ours) or cut off at its plane (it
        // becomes one variable).  Variables must be INDEPENDENT for the table to be exact on every reachable assignment: a
        // variable x whose own definition reads other variables of the union is substituted by that definition
        // (Xor3: mid = b*c next to b and c).  Tried in this order: cut / uncut with at most 3 variables (one LUT), then
        // the same with the table limit K.
        auto substitute = [&](View& v, uint32_t x) {
          const TT& dx = tts[(size_t)val[x].tt];
          std::vector<uint32_t> ns;
          for (uint32_t s : v.sup) if (s != x) ns.push_back(s);
          for (uint32_t s : dx.sup) if (std::find(ns.begin(), ns.end(), s) == ns.end()) ns.push_back(s);
          std::sort(ns.begin(), ns.end());
          std::vector<U256> nt((size_t)1 << ns.size());
          auto bit_at = [&](size_t a, uint32_t s) { return (a >> (size_t)(std::lower_bound(ns.begin(), ns.end(), s) - ns.begin())) & 1u; };
          for (size_t a = 0; a < nt.size(); a++) {
            size_t dxi = 0;
            for (size_t q = 0; q < dx.sup.size(); q++) dxi |= bit_at(a, dx.sup[q]) << q;
            const size_t xv = dx.tab[dxi] == ONE ? 1 : 0;
            size_t old = 0;
            for (size_t q = 0; q < v.sup.size(); q++) old |= (v.sup[q] == x ? xv : bit_at(a, v.sup[q])) << q;
            nt[a] = v.tab[old];
          }
          v.sup.swap(ns); v.tab.swap(nt);
        };
        // every viable attempt is evaluated; the first whose table is constant or bit-valued wins (cut views can miss that:
        // siblings cut off at their planes lose what they have in common), otherwise the most inlined table is kept
        std::vector<uint32_t> sup;
        bool have_virtual = false;
        TT virt;
        for (int attempt = 0; attempt < 4 && val[i].k == V_NONE; attempt++) {
          const bool cutting = attempt < 2;                  // 0: cut, <= 3; 1: cut, siblings, <= K; 2: uncut, <= 3; 3: everything inlined that fits K
          const bool siblings = (attempt & 1) != 0;
          const uint32_t limit = siblings ? K : 3u;
          bool viewable = true;
          for (int k = 0; k < no; k++) {
            const Val& v = val[o[k]];
            if (v.k == V_CONST) vw[k] = View{{}, {v.c}};
            else if (v.k == V_TT && !(cutting && v.plane >= 0)) vw[k] = View{tts[(size_t)v.tt].sup, tts[(size_t)v.tt].tab};
            else if (v.plane >= 0) vw[k] = View{{o[k]}, {ZERO, ONE}};
            else viewable = false;
          }
          if (!viewable) break;
          for (int guard = 0; guard < 64; guard++) {
            sup.clear();
            for (int k = 0; k < no; k++) for (uint32_t s : vw[k].sup) if (std::find(sup.begin(), sup.end(), s) == sup.end()) sup.push_back(s);
            // a variable whose definition shares variables with the rest of the union
            uint32_t dep = 0xFFFFFFFFu;
            for (uint32_t x : sup) {
              if (val[x].k != V_TT) continue;
              const TT& dx = tts[(size_t)val[x].tt];
              if (dx.sup.size() == 1 && dx.sup[0] == x) continue;
              size_t shared = 0, fresh = 0;
              for (uint32_t s : dx.sup) { if (std::find(sup.begin(), sup.end(), s) != sup.end()) shared++; else fresh++; }
              if (siblings && shared == 0) {
                // ... or with the definition of another variable (x = a*b next to y = (1-a)*b)
                for (uint32_t y : sup) {
                  if (y == x || val[y].k != V_TT) continue;
                  const TT& dy = tts[(size_t)val[y].tt];
                  if (dy.sup.size() == 1 && dy.sup[0] == y) continue;
                  for (uint32_t s : dx.sup) if (std::find(dy.sup.begin(), dy.sup.end(), s) != dy.sup.end()) shared++;
                }
              }
              if (attempt == 3) shared = 1;                   // last resort: look through every definition that fits
              if (shared > 0 && sup.size() - 1 + fresh <= K) { dep = x; break; }
            }
            if (dep == 0xFFFFFFFFu) break;
            for (int k = 0; k < no; k++) if (std::find(vw[k].sup.begin(), vw[k].sup.end(), dep) != vw[k].sup.end()) substitute(vw[k], dep);
          }
          if (sup.size() > limit) continue;
          std::sort(sup.begin(), sup.end());
          const size_t n = sup.size();
          std::vector<U256> tab((size_t)1 << n);
          uint32_t st = 0;
          for (size_t a = 0; a < tab.size(); a++) {
            fe x[3] = {fe_zero(), fe_zero(), fe_zero()};
            for (int k = 0; k < no; k++) {
              size_t idx = 0;
              const std::vector<uint32_t>& s = vw[k].sup;
              for (size_t q = 0; q < s.size(); q++) {
                const size_t pos = (size_t)(std::lower_bound(sup.begin(), sup.end(), s[q]) - sup.begin());
                if ((a >> pos) & 1) idx |= (size_t)1 << q;
              }
              x[k] = to_fe(vw[k].tab[idx]);
            }
            tab[a] = to_u256(alu_exec(opc, x[0], x[1], x[2], st));
          }
          if (st) throw Fail{"an operation the reference leaves undefined (Shl overflow, Bor/Bxor == M, Pow, Id, Lnot, Bnot) can occur"};
          // drop the variables the table does not depend on
          std::vector<uint32_t> sup2 = sup;
          for (size_t k = sup2.size(); k-- > 0;) {
            bool dep = false;
            for (size_t a = 0; a < tab.size() && !dep; a++) if (!((a >> k) & 1) && !(tab[a] == tab[a | ((size_t)1 << k)])) dep = true;
            if (dep) continue;
            std::vector<U256> t2(tab.size() / 2);
            for (size_t a = 0; a < t2.size(); a++) { const size_t lo = a & (((size_t)1 << k) - 1), hi = a >> k; t2[a] = tab[lo | (hi << (k + 1))]; }
            tab.swap(t2);
            sup2.erase(sup2.begin() + (long)k);
          }
          if (sup2.empty()) { set_const(i, tab[0]); bp.n_nodes_tt++; break; }
          bool bits = true;
          for (const U256& t : tab) bits &= is_bit_const(t);
          if (!bits) {
            // not (provably) a bit in this view: remember the first (smallest) table in case no view proves it
            if (!have_virtual) { virt.sup = sup2; virt.tab = tab; have_virtual = true; }
            continue;
          }
          std::vector<uint32_t> leaves;
          uint64_t tb = 0;
          for (uint32_t s : sup2) leaves.push_back((uint32_t)val[s].plane);
          for (size_t a = 0; a < tab.size(); a++) if (tab[a] == ONE) tb |= 1ull << a;
          const int32_t plane = (int32_t)bg.func(leaves, tb);
          bp.n_nodes_bit++;
          if (plane <= 1) { set_const(i, plane ? ONE : ZERO); break; }
          TT t; t.sup = sup2; t.tab = tab;
          tts.push_back(std::move(t));
          val[i] = Val(); val[i].k = V_TT; val[i].tt = (int32_t)tts.size() - 1; val[i].plane = plane;
        }
        if (val[i].k == V_NONE && have_virtual) {
          // a table of small integers (a sum of a few bits) is integer logic; a table of field-sized entries is what gives
          // field arithmetic away (see the eligibility test at the end)
          int width = 0;
          for (const U256& t : virt.tab) width = std::max(width, Big(t).bits());
          if (width <= 64) bp.n_nodes_bv++; else bp.n_nodes_tt++;
          tts.push_back(std::move(virt));
          val[i] = Val(); val[i].k = V_TT; val[i].tt = (int32_t)tts.size() - 1; val[i].plane = -1;
        }
        if (val[i].k != V_NONE) continue;
      }

      // ---- 2. bit-vector domain ------------------------------------------------------------------------------------------
      if (nd.kind != N_DUO) throw Fail{"a unary or ternary operation on values that are not functions of a few bits"};
      int32_t ia = to_bv(o[0]), ib = to_bv(o[1]);
      if (ia < 0 || ib < 0) throw Fail{std::string("node ") + std::to_string(i) + ": operand of op " + std::to_string(opc) + " is not a small non-negative integer under the bit contract"};
      BV r;
      // a constant shift amount; anything >= 2^62 reads as "at least 254" (the reference's cut-off)
      auto shift_const = [&](const Val& v, uint64_t* k) { if (v.k != V_CONST) return false; const Big c(v.c); *k = c.bits() > 16 ? 1000 : c.v.l[0]; return true; };
      switch (opc) {
        case OP_ADD: {
          if (!Big::add(bvs[(size_t)ia].hi, bvs[(size_t)ib].hi, &r.hi) || !r.hi.below_modulus()) throw Fail{"a sum of integers can reach the modulus"};
          r.compressed = false;
          for (int side = 0; side < 2; side++) {
            const uint32_t on = o[side];
            BV& s = bvs[(size_t)(side ? ib : ia)];
            // an uncompressed heap with this Add as its only reader is merged as it is; anything else is compressed once
            if (!s.compressed && !(uses[on] == 1)) compress(s);
            if (r.cols.size() < s.cols.size()) r.cols.resize(s.cols.size());
            for (size_t p = 0; p < s.cols.size(); p++) for (uint32_t x : s.cols[p]) if (x != 0) r.cols[p].push_back(x);
          }
          break;
        }
        case OP_MUL: {
          const bool ca = val[o[0]].k == V_CONST, cb = val[o[1]].k == V_CONST;
          BV& x = bvs[(size_t)(ca ? ib : ia)];
          BV& y = bvs[(size_t)(ca ? ia : ib)];
          if (!Big::mul(x.hi, y.hi, &r.hi) || !r.hi.below_modulus()) throw Fail{"a product of integers can reach the modulus"};
          r.compressed = false;
          compress(x);
          if (ca || cb) {
            const Big c = y.hi;                                 // the constant's value
            for (int s = 0; s < c.bits(); s++) {
              if (!c.bit(s)) continue;
              if (r.cols.size() < x.cols.size() + (size_t)s) r.cols.resize(x.cols.size() + (size_t)s);
              for (size_t p = 0; p < x.cols.size(); p++) { const uint32_t pl = plane_of_col(x, p); if (pl) r.cols[p + (size_t)s].push_back(pl); }
            }
          } else {
            compress(y);
            if (x.cols.size() * y.cols.size() > 1024) throw Fail{"product of two wide integers"};
            r.cols.resize(x.cols.size() + y.cols.size());
            for (size_t p = 0; p < x.cols.size(); p++) for (size_t q = 0; q < y.cols.size(); q++) {
              const uint32_t pl = bg.lut2(0x8, plane_of_col(x, p), plane_of_col(y, q));
              if (pl) r.cols[p + q].push_back(pl);
            }
          }
          break;
        }
        case OP_SHL: {
          uint64_t k;
          if (!shift_const(val[o[1]], &k)) throw Fail{"shift by a non-constant amount"};
          const BV& x = bvs[(size_t)ia];
          if (k >= 254) break;                                 // graph.rs:621-635: b >= 254 -> 0 (r.hi stays 0)
          if (!Big::shl(x.hi, (uint32_t)k, &r.hi) || !r.hi.below_modulus()) throw Fail{"a left shift can reach the modulus"};
          r.compressed = x.compressed;
          r.cols.assign((size_t)k, {});
          r.cols.insert(r.cols.end(), x.cols.begin(), x.cols.end());
          break;
        }
        case OP_SHR: {
          uint64_t k;
          if (!shift_const(val[o[1]], &k)) throw Fail{"shift by a non-constant amount"};
          BV& x = bvs[(size_t)ia];
          compress(x);
          r.hi = k >= 254 ? Big() : Big::shr(x.hi, (uint32_t)k);   // graph.rs:637-672 (b >= 254 -> 0)
          for (size_t p = (size_t)std::min<uint64_t>(k, x.cols.size()); p < x.cols.size(); p++) r.cols.push_back(x.cols[p]);
          break;
        }
        case OP_BAND: case OP_BOR: case OP_BXOR: {
          BV& x = bvs[(size_t)ia];
          BV& y = bvs[(size_t)ib];
          compress(x); compress(y);
          const size_t w = opc == OP_BAND ? std::min(x.cols.size(), y.cols.size()) : std::max(x.cols.size(), y.cols.size());
          const uint32_t t4 = opc == OP_BAND ? 0x8u : opc == OP_BOR ? 0xEu : 0x6u;
          for (size_t p = 0; p < w; p++) { const uint32_t pl = bg.lut2(t4, plane_of_col(x, p), plane_of_col(y, p)); r.cols.push_back(pl ? std::vector<uint32_t>{pl} : std::vector<uint32_t>{}); }
          if (opc == OP_BAND) r.hi = x.hi < y.hi ? x.hi : y.hi;
          else {
            r.hi = Big::ones(std::max(x.hi.bits(), y.hi.bits()));
            if (!r.hi.below_modulus()) throw Fail{"a bitwise or/xor can reach the modulus"};     // then bit_or/bit_xor never reduce (graph.rs:689-717)
          }
          break;
        }
        default: throw Fail{std::string("op ") + std::to_string(opc) + " on integers that are not functions of a few bits"};
      }
      bp.n_nodes_bv++;
      if (r.hi.is_zero()) { set_const(i, ZERO); continue; }
      if (r.hi.is_one()) {                                          // a bit again (Band(x >> k, 1)): from here on a table leaf
        compress(r);
        set_leaf(i, plane_of_col(r, 0));
        continue;
      }
      // constant after all (every plane constant)?
      bvs.push_back(std::move(r));
      val[i] = Val(); val[i].k = V_BV; val[i].bv = (int32_t)bvs.size() - 1;
    }

    // ---- witness positions: bits or constants ---------------------------------------------------------------------------
    std::vector<uint32_t> out_plane(bp.n_witness, 0);
    std::vector<uint32_t> wide_planes;                       // LUT DAG node per wide plane, stored at plane index W + k
    bp.const_of_pos.assign(bp.n_witness, -1);
    std::map<U256, int32_t> cix;
    auto const_out = [&](uint32_t j, const U256& c) {
      auto it = cix.find(c);
      if (it == cix.end()) { it = cix.emplace(c, (int32_t)bp.const_vals.size()).first; bp.const_vals.push_back(c); }
      bp.const_of_pos[j] = it->second;
    };
    for (uint32_t j = 0; j < bp.n_witness; j++) {
      const Val& v = val[g.witness_signals[j]];
      if (v.k == V_CONST) const_out(j, v.c);
      else if (v.plane >= 2) out_plane[j] = (uint32_t)v.plane;
      else if (v.plane >= 0) const_out(j, v.plane ? ONE : ZERO);
      else if (v.k == V_BV || to_bv(g.witness_signals[j]) >= 0) {
        // an integer of several bits (a bit heap, or a table of integers below M): its planes go behind the position
        // planes, the expansion assembles the value
        BV& b = bvs[(size_t)val[g.witness_signals[j]].bv];
        compress(b);
        bp.const_of_pos[j] = -2;
        bp.wide.push_back(j); bp.wide.push_back((uint32_t)wide_planes.size()); bp.wide.push_back((uint32_t)b.cols.size());
        for (size_t p = 0; p < b.cols.size(); p++) wide_planes.push_back(plane_of_col(b, p));
      }
      else throw Fail{"a witness signal is neither a bit, an integer below M nor a constant under the bit contract"};
    }
    bp.plane_stride = bp.n_witness + (uint32_t)wide_planes.size();

    // ---- dead LUT elimination, optional merging of single-use LUTs into their reader -----------------------------------
    const size_t NB = bg.nodes.size();
    std::vector<uint32_t> fan(NB, 0);
    auto count_fanout = [&]() {
      std::fill(fan.begin(), fan.end(), 0);
      std::vector<uint8_t> live(NB, 0);
      for (uint32_t j = 0; j < bp.n_witness; j++) if (bp.const_of_pos[j] == -1) { live[out_plane[j]] = 1; fan[out_plane[j]]++; }
      for (uint32_t pl : wide_planes) { live[pl] = 1; fan[pl]++; }
      for (size_t b = NB; b-- > 2;) {
        if (!live[b] || bg.nodes[b].kind != BGraph::K_LUT) continue;
        for (int q = 0; q < bg.nodes[b].n; q++) { live[bg.nodes[b].in[q]] = 1; fan[bg.nodes[b].in[q]]++; }
      }
      return live;
    };
    std::vector<uint8_t> live = count_fanout();
    if (opt.merge_luts) {
      // reader f(.., m, ..) with m = h(..) read by nobody else: compose when the union of the supports has <= 3 planes
      for (size_t b = 2; b < NB; b++) {
        if (!live[b] || bg.nodes[b].kind != BGraph::K_LUT) continue;
        for (bool again = true; again;) {
          again = false;
          BGraph::BN& f = bg.nodes[b];
          for (int q = 0; q < f.n && !again; q++) {
            const uint32_t m = f.in[q];
            if (bg.nodes[m].kind != BGraph::K_LUT || fan[m] != 1) continue;
            const BGraph::BN& h = bg.nodes[m];
            std::vector<uint32_t> sup;
            for (int r = 0; r < f.n; r++) if (r != q) sup.push_back(f.in[r]);
            for (int r = 0; r < h.n; r++) if (std::find(sup.begin(), sup.end(), h.in[r]) == sup.end()) sup.push_back(h.in[r]);
            if (sup.size() > 3) continue;
            uint32_t nt = 0;
            for (uint32_t a = 0; a < (1u << sup.size()); a++) {
              auto bit_of = [&](uint32_t id) { for (size_t r = 0; r < sup.size(); r++) if (sup[r] == id) return (a >> r) & 1u; return 0u; };
              uint32_t hi = 0;
              for (int r = 0; r < h.n; r++) hi |= bit_of(h.in[r]) << r;
              const uint32_t hv = (h.lut >> hi) & 1u;
              uint32_t fi = 0;
              for (int r = 0; r < f.n; r++) fi |= (r == q ? hv : bit_of(f.in[r])) << r;
              nt |= ((f.lut >> fi) & 1u) << a;
            }
            // readers: f drops its old inputs and reads `sup`; h is dead (fan only steers this heuristic: liveness is recomputed)
            for (int r = 0; r < f.n; r++) if (r != q) fan[f.in[r]]--;
            fan[m] = 0;
            for (uint32_t s : sup) fan[s]++;
            for (int r = 0; r < h.n; r++) fan[h.in[r]]--;
            f.n = (uint8_t)sup.size(); f.lut = (uint8_t)nt;
            for (size_t r = 0; r < 3; r++) f.in[r] = r < sup.size() ? sup[r] : 0;
            bp.n_merged++;
            again = true;
          }
        }
      }
      live = count_fanout();
    }

    // ---- levels, steps of 32 independent LUTs, plane slots -----------------------------------------------------------------
    // extra copies: a plane that sits at several witness positions, or an input plane that is a witness signal itself
    struct Emit { uint32_t node; uint32_t pos; bool copy; };
    std::vector<std::vector<uint32_t>> pos_of(NB);
    for (uint32_t j = 0; j < bp.n_witness; j++) if (bp.const_of_pos[j] == -1) pos_of[out_plane[j]].push_back(j);
    for (size_t k = 0; k < wide_planes.size(); k++) pos_of[wide_planes[k]].push_back(bp.n_witness + (uint32_t)k);
    std::vector<uint32_t> step_of(NB, 0);                      // step in which the plane is written; inputs: 0 (prologue), LUTs: >= 1
    std::vector<uint32_t> level(NB, 0);
    std::vector<uint32_t> fill(2, 0);                          // LUTs per step (index = step)
    std::vector<std::vector<Emit>> steps(2);
    uint32_t first_free = 1;
    auto place = [&](uint32_t min_step, Emit e) {
      uint32_t s = std::max(min_step, first_free);
      while (true) {
        if (s >= fill.size()) { fill.resize(s + 1, 0); steps.resize(s + 1); }
        if (fill[s] < 32) break;
        s++;
      }
      fill[s]++; steps[s].push_back(e);
      while (first_free < fill.size() && fill[first_free] >= 32) first_free++;
      return s;
    };
    std::vector<uint32_t> order;
    for (size_t b = 2; b < NB; b++) {
      if (!live[b]) continue;
      if (bg.nodes[b].kind == BGraph::K_LUT) {
        uint32_t lv = 0;
        for (int q = 0; q < bg.nodes[b].n; q++) lv = std::max(lv, level[bg.nodes[b].in[q]]);
        level[b] = lv + 1;
        bp.n_levels = std::max<uint64_t>(bp.n_levels, level[b]);
        order.push_back((uint32_t)b);
      }
    }
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return level[a] < level[b]; });
    std::vector<uint32_t> last_read(NB, 0);                     // last step that reads the plane
    for (uint32_t b : order) {
      uint32_t ms = 1;
      for (int q = 0; q < bg.nodes[b].n; q++) ms = std::max(ms, step_of[bg.nodes[b].in[q]] + 1);
      const std::vector<uint32_t>& ps = pos_of[b];
      step_of[b] = place(ms, Emit{b, ps.empty() ? BIT_NO_POS : ps[0], false});
      for (int q = 0; q < bg.nodes[b].n; q++) last_read[bg.nodes[b].in[q]] = std::max(last_read[bg.nodes[b].in[q]], step_of[b]);
      bp.n_luts++;
    }
    for (size_t b = 0; b < NB; b++) {
      if (!live[b] && b >= 2) continue;
      const std::vector<uint32_t>& ps = pos_of[b];
      const bool is_lut = bg.nodes[b].kind == BGraph::K_LUT;
      for (size_t k = is_lut ? 1 : 0; k < ps.size(); k++) {
        const uint32_t s = place(step_of[b] + 1, Emit{(uint32_t)b, ps[k], true});
        last_read[b] = std::max(last_read[b], s);
      }
    }
    bp.n_steps = (uint32_t)steps.size();
    // slots: a plane's slot is free for LUTs of steps AFTER its last reader (lanes of one step are not ordered)
    std::vector<uint32_t> slot(NB, BIT_NO_SLOT);
    slot[0] = BIT_SLOT_ZERO; slot[1] = BIT_SLOT_ONES;
    uint32_t n_slots = 2;
    std::vector<std::vector<uint32_t>> dying(bp.n_steps + 1);
    std::vector<uint32_t> free_slots;
    auto take = [&](uint32_t b) {
      if (!free_slots.empty()) { slot[b] = free_slots.back(); free_slots.pop_back(); }
      else slot[b] = n_slots++;
      if (n_slots > std::min<uint32_t>(opt.max_slots, 0xFFFE)) throw Fail{"too many planes alive at once for the shared-memory plane file"};
      dying[last_read[b]].push_back(b);
    };
    // every input the typing looked at is checked against the contract, also those whose plane ended up unread
    for (auto& kv : bg.input_ids) {                          // ordered by (input index, bit)
      const bool used = live[kv.second] && last_read[kv.second] > 0;
      if (used) take(kv.second);
      if (!used && kv.first.second != BIT_CONTRACT) continue; // an unread bit of a field input needs no plane (and no check)
      bp.inputs.push_back(kv.first.first); bp.inputs.push_back(slot[kv.second]); bp.inputs.push_back(kv.first.second);
    }
    bp.code.assign((size_t)bp.n_steps * 32, BitOp{0, 0, BIT_NO_SLOT << 16, BIT_NO_POS});
    for (uint32_t s = 1; s < bp.n_steps; s++) {
      // operands first (they are all older), then the destinations of this step
      for (const Emit& e : steps[s]) if (!e.copy && last_read[e.node] > 0) take(e.node);
      uint32_t lane = 0;
      for (const Emit& e : steps[s]) {
        BitOp& op = bp.code[(size_t)s * 32 + lane++];
        const BGraph::BN& b = bg.nodes[e.node];
        if (e.copy) { op.x = 0xAA; op.y = slot[e.node]; op.z = BIT_NO_SLOT << 16; op.w = e.pos; continue; }
        // 8-bit table over (a, b, c); unused inputs read the zero plane
        uint32_t t8 = 0;
        for (uint32_t a = 0; a < 8; a++) t8 |= ((b.lut >> (a & ((1u << b.n) - 1))) & 1u) << a;
        uint32_t in[3] = {0, 0, 0};
        for (int q = 0; q < b.n; q++) { in[q] = slot[b.in[q]]; if (in[q] == BIT_NO_SLOT) throw Fail{"internal: operand without a slot"}; }
        op.x = t8; op.y = in[0] | (in[1] << 16); op.z = in[2] | (slot[e.node] << 16); op.w = e.pos;
      }
      for (uint32_t b : dying[s]) free_slots.push_back(slot[b]);
    }
    bp.n_slots = n_slots;
    // A graph of field arithmetic is "Boolean" too if one pretends that its inputs are bits -- every node is then a table
    // of field values over a few bits -- but nobody feeds such a graph bits.  Its typing gives it away: (almost) no node
    // is a bit or an integer, everything is a table (Poseidon(1): 600 tables, no bit), or 254-bit table entries get
    // synthesised into thousands of LUTs per operation.  Not worth a speculation.
    size_t live_ops = 0;
    for (size_t i = 0; i < N; i++) live_ops += needed[i] && g.nodes[i].kind >= N_UNO;
    if ((bp.n_nodes_bit + bp.n_nodes_bv) * 8 < bp.n_nodes_tt || bp.n_luts > 8 * live_ops + 64)
      throw Fail{"the operations of this graph are field arithmetic, not logic: its inputs are hardly meant to be bits"};
    bp.eligible = true;
  } catch (const Fail& f) {
    bp.eligible = false;
    bp.reason = f.why;
    bp.code.clear(); bp.inputs.clear();
  }
  return bp;
}

}  // namespace gw
