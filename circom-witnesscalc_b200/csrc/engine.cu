// Batch witness evaluator for sm_100a: the device half of what replaces
// /root/reference/src/graph.rs:367-391 (graph::evaluate).
//
// Execution model (throughput mode): one thread = one input set ("witness").  All 32 lanes of a
// warp run the SAME instruction of the plan for 32 different witnesses, so opcode dispatch is
// warp-uniform and there is no divergence.  Each warp streams the 16-byte instructions itself:
// lane l loads instruction base+l (one coalesced 512 B read), the next block of 32 is prefetched
// while the current one executes, and the instruction being executed is broadcast with shuffles, so
// warps never synchronise with each other.  Values live in a per-witness register file in shared
// memory laid out [register][half][thread] as uint4 (conflict-free LDS.128/STS.128); values that do
// not fit are spilled to HBM witness-major ([slot][half][thread], coalesced 128-bit accesses).
// Witness values are written canonical, 32 B each, at witness[w][position].
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <thread>

#include "engine.hpp"
#include "alu.cuh"

namespace gw {

struct KParams {
  const uint4* code; uint32_t n_slots;
  const uint4* consts;
  const uint4* inputs;     // [B][I][2]
  uint4* out;              // [B][W][2]
  uint4* spill;            // [n_spill][2][V][spill_threads]
  uint32_t* status;        // [B] or null
  unsigned long long B;
  uint32_t I, W;
  uint32_t n_tiles;
  unsigned long long spill_threads;
  unsigned long long out_wrap;   // profiling aid (GW_DEBUG_OUT_WRAP): witness rows wrap modulo this many rows; 0 = off
};

__device__ __forceinline__ fe fe_from(uint4 lo, uint4 hi) {
  fe r; r.l[0] = lo.x; r.l[1] = lo.y; r.l[2] = lo.z; r.l[3] = lo.w; r.l[4] = hi.x; r.l[5] = hi.y; r.l[6] = hi.z; r.l[7] = hi.w; return r;
}
__device__ __forceinline__ uint4 fe_lo(const fe& a) { return make_uint4(a.l[0], a.l[1], a.l[2], a.l[3]); }
__device__ __forceinline__ uint4 fe_hi(const fe& a) { return make_uint4(a.l[4], a.l[5], a.l[6], a.l[7]); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

static const int RING = 64;     // instruction slots staged per warp in shared memory (two blocks of 32)

// T threads per CTA, V input sets per thread.  Shared memory: [T/32 warps][RING] instruction slots, then the
// register file [n_regs][2 halves][V][T] of uint4.
template <int T, int V>
__global__ void __launch_bounds__(T) eval_batch_kernel(const KParams p) {
  extern __shared__ uint4 smem[];
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  uint4* ring = smem + (tid >> 5) * RING;
  uint4* rf = smem + (T / 32) * RING;
  const unsigned long long gthread = (unsigned long long)blockIdx.x * T + tid;
  const uint32_t n = p.n_slots;

  auto rf_load = [&](uint32_t r, int v) { return fe_from(rf[((r * 2) * V + v) * T + tid], rf[((r * 2 + 1) * V + v) * T + tid]); };
  auto rf_store = [&](uint32_t r, int v, const fe& x) { rf[((r * 2) * V + v) * T + tid] = fe_lo(x); rf[((r * 2 + 1) * V + v) * T + tid] = fe_hi(x); };
  auto const_load = [&](uint32_t c) { return fe_from(__ldg(p.consts + 2 * (size_t)c), __ldg(p.consts + 2 * (size_t)c + 1)); };

  for (uint32_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    const uint4* in[V]; uint4* out[V]; bool active[V]; unsigned long long wv[V];
    uint32_t st[V];
#pragma unroll
    for (int v = 0; v < V; v++) {
      wv[v] = ((unsigned long long)tile * V + v) * T + tid;
      active[v] = wv[v] < p.B;
      const unsigned long long wl = active[v] ? wv[v] : p.B - 1;
      in[v] = p.inputs + wl * p.I * 2;
      out[v] = p.out + (p.out_wrap ? wl % p.out_wrap : wl) * p.W * 2;
      st[v] = 0;
    }
    auto out_store = [&](uint32_t j, int v, const fe& x) { if (active[v]) { out[v][2 * (size_t)j] = fe_lo(x); out[v][2 * (size_t)j + 1] = fe_hi(x); } };

    // pull what the NEXT instruction reads from global memory (table constants, spilled values) towards L1 while
    // the current instruction computes; its header (and term slots) are already in the ring
    auto prefetch_one = [&](const uint4& h, uint32_t at) {
      const uint32_t o = h.x & 0xFFu;
      if (o == OP_DOT) {
        const uint32_t nt = h.y & 0xFFu;
#pragma unroll 1
        for (uint32_t t = 0; t < nt; t += 2) {
          const uint4 sl = ring[(at + 1 + (t >> 1)) & (RING - 1)];
          if ((sl.x & 0xFu) != T_ADDHI && (sl.x & 0xFu) != T_SUBHI) prefetch_l1(p.consts + 2 * (size_t)sl.y);
          if (t + 1 < nt && (sl.z & 0xFu) != T_ADDHI && (sl.z & 0xFu) != T_SUBHI) prefetch_l1(p.consts + 2 * (size_t)sl.w);
        }
      } else if (o == OP_SPILL_LD) {
#pragma unroll
        for (int v = 0; v < V; v++) {
          prefetch_l1(p.spill + (((size_t)h.y * 2) * V + v) * p.spill_threads + gthread);
          prefetch_l1(p.spill + (((size_t)h.y * 2 + 1) * V + v) * p.spill_threads + gthread);
        }
      } else if (o == OP_SHRAND) {
        prefetch_l1(p.consts + 2 * (size_t)(h.z >> 8));
      } else if (o < 48u) {
        if (h.x & F_A_CONST) prefetch_l1(p.consts + 2 * (size_t)h.y);
        if ((h.x & F_B_CONST) && o < 32u) prefetch_l1(p.consts + 2 * (size_t)h.z);
      }
    };

    auto prefetch_operands = [&](const uint4& h, uint32_t at) {
      prefetch_one(h, at);
      if ((V == 1) && (h.x & F_PAIR)) {
        const uint32_t at2 = at + (((h.x & 0xFFu) == OP_DOT) ? 1u + (((h.y & 0xFFu) + 1u) >> 1) : 1u);
        prefetch_one(ring[at2 & (RING - 1)], at2);
      }
    };

    // instruction ring: blocks 0 and 1 staged, block 2 in flight in nblk
    __syncwarp();
    ring[lane] = __ldg(p.code + min((uint32_t)lane, n - 1));
    ring[32 + lane] = __ldg(p.code + min(32u + lane, n - 1));
    uint4 nblk = __ldg(p.code + min(64u + lane, n - 1));
    __syncwarp();
    uint32_t pc = 0;
    uint4 nxt = ring[0];
    while (pc < n) {
      const uint4 ins = nxt;
      const uint32_t op = ins.x & 0xFFu;
      const uint32_t dst = ins.x >> 16;
      const uint32_t len = (op == OP_DOT) ? 1u + (((ins.y & 0xFFu) + 1u) >> 1) : 1u;
      uint32_t npc = pc + len;
      uint4 ins2 = ins;
      const bool paired = (V == 1) && (ins.x & F_PAIR);
      if (paired) {                                   // bundle: the partner follows immediately
        ins2 = ring[npc & (RING - 1)];
        npc += ((ins2.x & 0xFFu) == OP_DOT) ? 1u + (((ins2.y & 0xFFu) + 1u) >> 1) : 1u;
      }
      const bool cross = (npc >> 5) != (pc >> 5);
      if (!cross) { nxt = ring[npc & (RING - 1)]; if (npc < n) prefetch_operands(nxt, npc); }

      fe R[V];
      bool have_result = true;
      if (paired) {
        // ---- two independent instructions of the same class as one bundle (V == 1): their carry chains sit in the
        // same basic block, so the scheduler interleaves them; every operand is read before either result is stored
        fe R1, R2;
        if (op == OP_DOT) {
          const uint32_t nt1 = ins.y & 0xFFu, nt2 = ins2.y & 0xFFu;
          const uint32_t base1 = pc + 1, base2 = pc + len + 1;
          const uint32_t tmin = min(nt1, nt2);
          uint32_t P1[16], P2[16];
#pragma unroll
          for (int k = 0; k < 16; k++) { P1[k] = 0; P2[k] = 0; }
          auto term_at = [&](uint32_t base, uint32_t t, uint32_t& kind, uint32_t& reg, uint32_t& ci) {
            const uint4 sl = ring[(base + (t >> 1)) & (RING - 1)];
            const uint32_t lo = (t & 1) ? sl.z : sl.x;
            ci = (t & 1) ? sl.w : sl.y; kind = lo & 0xFu; reg = lo >> 16;
          };
          auto one_term = [&](uint32_t* P, uint32_t kind, uint32_t reg, uint32_t ci) {
            if (kind == T_MAC) { const fe c = const_load(ci); const fe x = rf_load(reg, 0); uint32_t Q[16]; u256_mul_wide(Q, x.l, c.l); u512_add(P, Q); }
            else if (kind == T_CONST) { const fe c = const_load(ci); u512_add256(P, c.l, 0); }
            else dot_term(P, kind, rf_load(reg, 0), fe_zero());
          };
#pragma unroll 1
          for (uint32_t t = 0; t < tmin; t++) {
            uint32_t k1, r1, c1, k2, r2, c2;
            term_at(base1, t, k1, r1, c1); term_at(base2, t, k2, r2, c2);
            if (k1 == T_MAC && k2 == T_MAC) {
              const fe cv1 = const_load(c1), cv2 = const_load(c2);
              const fe x1 = rf_load(r1, 0), x2 = rf_load(r2, 0);
              uint32_t Q1[16], Q2[16];
              u256_mul_wide(Q1, x1.l, cv1.l); u256_mul_wide(Q2, x2.l, cv2.l);
              u512_add(P1, Q1); u512_add(P2, Q2);
            } else { one_term(P1, k1, r1, c1); one_term(P2, k2, r2, c2); }
          }
#pragma unroll 1
          for (uint32_t t = tmin; t < nt1; t++) { uint32_t k1, r1, c1; term_at(base1, t, k1, r1, c1); one_term(P1, k1, r1, c1); }
#pragma unroll 1
          for (uint32_t t = tmin; t < nt2; t++) { uint32_t k2, r2, c2; term_at(base2, t, k2, r2, c2); one_term(P2, k2, r2, c2); }
          R1 = fe_mont_reduce_core(P1); R2 = fe_mont_reduce_core(P2);
          fe_cond_sub_n(R1, (int)((ins.y >> 8) & 0xFFu)); fe_cond_sub_n(R2, (int)((ins2.y >> 8) & 0xFFu));
        } else {                                       // MUL / SQR x MUL / SQR
          const uint32_t op2 = ins2.x & 0xFFu;
          const fe A1 = (ins.x & F_A_CONST) ? const_load(ins.y) : rf_load(ins.y, 0);
          const fe B1 = (op == OP_SQR) ? A1 : ((ins.x & F_B_CONST) ? const_load(ins.z) : rf_load(ins.z, 0));
          const fe A2 = (ins2.x & F_A_CONST) ? const_load(ins2.y) : rf_load(ins2.y, 0);
          const fe B2 = (op2 == OP_SQR) ? A2 : ((ins2.x & F_B_CONST) ? const_load(ins2.z) : rf_load(ins2.z, 0));
          uint32_t Pm1[16], Pm2[16];
          u256_mul_wide(Pm1, A1.l, B1.l); u256_mul_wide(Pm2, A2.l, B2.l);
          R1 = fe_barrett(Pm1); R2 = fe_barrett(Pm2);
        }
        const uint32_t dst2 = ins2.x >> 16;
        if (dst != NO_DST) rf_store(dst, 0, R1);
        if (dst2 != NO_DST) rf_store(dst2, 0, R2);
        if (ins.x & F_OUT) out_store(ins.w, 0, R1);
        if (ins2.x & F_OUT) out_store(ins2.w, 0, R2);
        have_result = false;
      } else if (op == OP_DOT) {
        // fused linear combination: 512-bit accumulators, one Montgomery reduction per input set
        const uint32_t nt = ins.y & 0xFFu, ncs = (ins.y >> 8) & 0xFFu;
        uint32_t P[V][16];
#pragma unroll
        for (int v = 0; v < V; v++) {
#pragma unroll
          for (int k = 0; k < 16; k++) P[v][k] = 0;
        }
#pragma unroll 1
        for (uint32_t t = 0; t < nt; t++) {
          const uint4 sl = ring[(pc + 1 + (t >> 1)) & (RING - 1)];
          const uint32_t lo = (t & 1) ? sl.z : sl.x, ci = (t & 1) ? sl.w : sl.y;
          const uint32_t kind = lo & 0xFu, reg = lo >> 16;
          if (kind == T_MAC) {
            const fe c = const_load(ci);
            uint32_t Q[V][16];
#pragma unroll
            for (int v = 0; v < V; v++) { const fe x = rf_load(reg, v); u256_mul_wide(Q[v], x.l, c.l); }
#pragma unroll
            for (int v = 0; v < V; v++) u512_add(P[v], Q[v]);
          } else if (kind == T_CONST) {
            const fe c = const_load(ci);
#pragma unroll
            for (int v = 0; v < V; v++) u512_add256(P[v], c.l, 0);
          } else {
#pragma unroll
            for (int v = 0; v < V; v++) dot_term(P[v], kind, rf_load(reg, v), fe_zero());
          }
        }
#pragma unroll
        for (int v = 0; v < V; v++) R[v] = fe_mont_reduce(P[v], (int)ncs);
      } else if (op == OP_MUL || op == OP_SQR) {
        fe A[V], Bv[V];
        if (ins.x & F_A_CONST) { const fe c = const_load(ins.y);
#pragma unroll
          for (int v = 0; v < V; v++) A[v] = c;
        } else {
#pragma unroll
          for (int v = 0; v < V; v++) A[v] = rf_load(ins.y, v);
        }
        if (op == OP_SQR) {
#pragma unroll
          for (int v = 0; v < V; v++) Bv[v] = A[v];
        } else if (ins.x & F_B_CONST) { const fe c = const_load(ins.z);
#pragma unroll
          for (int v = 0; v < V; v++) Bv[v] = c;
        } else {
#pragma unroll
          for (int v = 0; v < V; v++) Bv[v] = rf_load(ins.z, v);
        }
        uint32_t Pm[V][16];
#pragma unroll
        for (int v = 0; v < V; v++) u256_mul_wide(Pm[v], A[v].l, Bv[v].l);
#pragma unroll
        for (int v = 0; v < V; v++) R[v] = fe_barrett(Pm[v]);
      } else if (op == OP_ADD || op == OP_SUB) {
        fe A[V], Bv[V];
        if (ins.x & F_A_CONST) { const fe c = const_load(ins.y);
#pragma unroll
          for (int v = 0; v < V; v++) A[v] = c;
        } else {
#pragma unroll
          for (int v = 0; v < V; v++) A[v] = rf_load(ins.y, v);
        }
        if (ins.x & F_B_CONST) { const fe c = const_load(ins.z);
#pragma unroll
          for (int v = 0; v < V; v++) Bv[v] = c;
        } else {
#pragma unroll
          for (int v = 0; v < V; v++) Bv[v] = rf_load(ins.z, v);
        }
#pragma unroll
        for (int v = 0; v < V; v++) R[v] = (op == OP_ADD) ? fe_add(A[v], Bv[v]) : fe_sub(A[v], Bv[v]);
      } else if (op == OP_SPILL_ST) {
#pragma unroll
        for (int v = 0; v < V; v++) {
          const fe x = rf_load(ins.y, v);
          p.spill[(((size_t)ins.z * 2) * V + v) * p.spill_threads + gthread] = fe_lo(x);
          p.spill[(((size_t)ins.z * 2 + 1) * V + v) * p.spill_threads + gthread] = fe_hi(x);
        }
        have_result = false;
      } else if (op == OP_SPILL_LD) {
#pragma unroll
        for (int v = 0; v < V; v++)
          R[v] = fe_from(p.spill[(((size_t)ins.y * 2) * V + v) * p.spill_threads + gthread],
                         p.spill[(((size_t)ins.y * 2 + 1) * V + v) * p.spill_threads + gthread]);
      } else if (op == OP_OUT) {
        if (ins.x & F_A_CONST) { const fe c = const_load(ins.y);
#pragma unroll
          for (int v = 0; v < V; v++) out_store(ins.w, v, c);
        } else {
#pragma unroll
          for (int v = 0; v < V; v++) out_store(ins.w, v, rf_load(ins.y, v));
        }
        have_result = false;
      } else if (op == OP_INPUT) {
#pragma unroll
        for (int v = 0; v < V; v++) R[v] = fe_reduce256(fe_from(__ldg(in[v] + 2 * (size_t)ins.y), __ldg(in[v] + 2 * (size_t)ins.y + 1)));
      } else if (op == OP_SHRAND) {
        const fe c = const_load(ins.z >> 8);
#pragma unroll
        for (int v = 0; v < V; v++) R[v] = fe_shr_and(rf_load(ins.y, v), ins.z & 0xFFu, c);
      } else if (op == OP_INV || op == OP_DIV) {
        // modular inversion for all V input sets at once (interleaved safegcd chains); Div = a * b^-1, b == 0 -> 0
        fe X[V];
        const uint32_t src = (op == OP_INV) ? ins.y : ins.z;
        const uint32_t src_const = (op == OP_INV) ? (ins.x & F_A_CONST) : (ins.x & F_B_CONST);
        if (src_const) { const fe c = const_load(src);
#pragma unroll
          for (int v = 0; v < V; v++) X[v] = c;
        } else {
#pragma unroll
          for (int v = 0; v < V; v++) X[v] = rf_load(src, v);
        }
        fe_inv_batch<V>(X);
        if (op == OP_DIV) {
          uint32_t Pm[V][16];
          if (ins.x & F_A_CONST) { const fe c = const_load(ins.y);
#pragma unroll
            for (int v = 0; v < V; v++) u256_mul_wide(Pm[v], c.l, X[v].l);
          } else {
#pragma unroll
            for (int v = 0; v < V; v++) { const fe a = rf_load(ins.y, v); u256_mul_wide(Pm[v], a.l, X[v].l); }
          }
#pragma unroll
          for (int v = 0; v < V; v++) R[v] = fe_barrett(Pm[v]);
        } else {
#pragma unroll
          for (int v = 0; v < V; v++) R[v] = X[v];
        }
      } else if (op != OP_NOP) {
        // everything else (rare ops, out-of-line helpers), one input set after the other
#pragma unroll 1
        for (int v = 0; v < V; v++) {
          fe A = (ins.x & F_A_CONST) ? const_load(ins.y) : rf_load(ins.y, v), Bv = fe_zero(), C = fe_zero();
          if (op_has_b(op)) Bv = (ins.x & F_B_CONST) ? const_load(ins.z) : rf_load(ins.z, v);
          if (op == OP_TERN) C = (ins.x & F_C_CONST) ? const_load(ins.w) : rf_load(ins.w, v);
          uint32_t s = 0;
          const fe r = alu_exec(op, A, Bv, C, s);
#pragma unroll
          for (int u = 0; u < V; u++) if (u == v) { R[u] = r; st[u] |= s; }
        }
      } else {
        have_result = false;
      }
      if (have_result) {
        if (dst != NO_DST) {
#pragma unroll
          for (int v = 0; v < V; v++) rf_store(dst, v, R[v]);
        }
        if (ins.x & F_OUT) {
#pragma unroll
          for (int v = 0; v < V; v++) out_store(ins.w, v, R[v]);
        }
      }

      if (cross) {
        // block (pc >> 5) is consumed: stage block (pc >> 5) + 2 in its half, start loading the one after it
        __syncwarp();
        ring[((pc >> 5) & 1u) * 32u + lane] = nblk;
        nblk = __ldg(p.code + min(((pc >> 5) + 3u) * 32u + lane, n - 1));
        __syncwarp();
        nxt = ring[npc & (RING - 1)];
        if (npc < n) prefetch_operands(nxt, npc);
      }
      pc = npc;
    }
#pragma unroll
    for (int v = 0; v < V; v++) if (p.status != nullptr && active[v]) p.status[wv[v]] = st[v];
  }
}

// ---- single-witness latency mode ------------------------------------------------------------------------
// One CTA evaluates ONE witness: the instructions of a dependency level are spread over the threads
// (intra-level node parallelism), values live in one shared-memory slot file, a CTA barrier separates
// levels.  Each thread prefetches its instruction of the next level while it executes the current one.
struct LParams {
  const uint4* code; const uint32_t* level_count; uint32_t n_levels;
  const uint4* consts;
  const uint4* inputs;     // [I][2]
  uint4* out;              // [W][2]
  uint32_t* status;        // [1] or null
};

template <int T>
__global__ void __launch_bounds__(T) eval_latency_kernel(const LParams p) {
  extern __shared__ uint4 slots[];     // [n_slots][2]
  const uint32_t tid = threadIdx.x;
  auto slot_load = [&](uint32_t r) { return fe_from(slots[2 * r], slots[2 * r + 1]); };
  auto const_load = [&](uint32_t c) { return fe_from(__ldg(p.consts + 2 * (size_t)c), __ldg(p.consts + 2 * (size_t)c + 1)); };
  auto operand = [&](uint32_t is_const, uint32_t idx) { return is_const ? const_load(idx) : slot_load(idx); };
  uint32_t st = 0;
  auto exec = [&](const uint4 ins) {
    const uint32_t op = ins.x & 0xFFu, dst = ins.x >> 16;
    fe R;
    if (op == OP_NOP) return;
    if (op == OP_OUT) { fe v = operand(ins.x & F_A_CONST, ins.y); p.out[2 * (size_t)ins.w] = fe_lo(v); p.out[2 * (size_t)ins.w + 1] = fe_hi(v); return; }
    if (op == OP_INPUT) {
      R = fe_reduce256(fe_from(__ldg(p.inputs + 2 * (size_t)ins.y), __ldg(p.inputs + 2 * (size_t)ins.y + 1)));
    } else if (op == OP_MUL || op == OP_SQR) {
      fe A = operand(ins.x & F_A_CONST, ins.y);
      fe Bv = (op == OP_SQR) ? A : operand(ins.x & F_B_CONST, ins.z);
      R = fe_mul(A, Bv);
    } else if (op == OP_ADD || op == OP_SUB) {
      fe A = operand(ins.x & F_A_CONST, ins.y), Bv = operand(ins.x & F_B_CONST, ins.z);
      R = (op == OP_ADD) ? fe_add(A, Bv) : fe_sub(A, Bv);
    } else {
      fe A = operand(ins.x & F_A_CONST, ins.y), Bv = fe_zero(), C = fe_zero();
      if (op_has_b(op)) Bv = operand(ins.x & F_B_CONST, ins.z);
      if (op == OP_TERN) C = operand(ins.x & F_C_CONST, ins.w);
      R = alu_exec(op, A, Bv, C, st);
    }
    if (dst != NO_DST) { slots[2 * dst] = fe_lo(R); slots[2 * dst + 1] = fe_hi(R); }
    if (ins.x & F_OUT) { p.out[2 * (size_t)ins.w] = fe_lo(R); p.out[2 * (size_t)ins.w + 1] = fe_hi(R); }
  };
  const uint4 nop = make_uint4(OP_NOP, 0, 0, 0);
  const uint32_t nl = p.n_levels;
  uint32_t pos = 0;
  uint32_t c0 = __ldg(p.level_count), c1 = nl > 1 ? __ldg(p.level_count + 1) : 0, c2 = nl > 2 ? __ldg(p.level_count + 2) : 0;
  uint4 nxt = tid < c0 ? __ldg(p.code + tid) : nop;
  for (uint32_t L = 0; L < nl; L++) {
    const uint4 ins = nxt;
    const uint32_t c = c0, npos = pos + c;
    nxt = tid < c1 ? __ldg(p.code + npos + tid) : nop;                    // next level's instruction
    const uint32_t c3 = (L + 3 < nl) ? __ldg(p.level_count + L + 3) : 0;
    if (tid < c) exec(ins);
    for (uint32_t i = tid + T; i < c; i += T) exec(__ldg(p.code + pos + i));   // wide levels
    if (nxt.x & F_A_CONST) prefetch_l1(p.consts + 2 * (size_t)nxt.y);
    if ((nxt.x & F_B_CONST) && (nxt.x & 0xFFu) < 32) prefetch_l1(p.consts + 2 * (size_t)nxt.z);
    __syncthreads();
    pos = npos; c0 = c1; c1 = c2; c2 = c3;
  }
  if (p.status != nullptr && st) atomicOr(p.status, st);
}

// ---- integer-pipe microbenchmark (roofline denominator for multiplication-heavy graphs) ---------
// WHICH = 0: rows of (mad.lo.cc, madc.hi.cc) pairs exactly as in u256_mul_wide -> IMAD.WIDE.U32(.X)
//            with carry predicates; counts one op per 32x32+64 multiply-accumulate,
//         1: mad.lo.u32 (IMAD), 2: mad.hi.u32 (IMAD.HI.U32), 3: add.u32 (ALU pipe, for comparison).
// Every multiplicand comes from another accumulator, so nothing is loop-invariant.  32 ops per
// loop iteration and thread.  Reports executed ops per second over the whole chip.
template <int WHICH>
__global__ void imad_bench_kernel(uint32_t* out, int iters, uint32_t y) {
  uint32_t x = threadIdx.x * 2654435761u + 12345u + blockIdx.x;
  uint32_t s[16];
#pragma unroll
  for (int i = 0; i < 16; i++) s[i] = x + i * 131u;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (WHICH == 0) {
#pragma unroll
        for (int r = 0; r < 2; r++) {
          uint32_t* e = s + 8 * r;
          uint32_t m = s[8 * (1 - r) + u];
          asm volatile("mad.lo.cc.u32 %0, %8, %12, %0; madc.hi.cc.u32 %1, %8, %12, %1;"
                       "madc.lo.cc.u32 %2, %9, %12, %2; madc.hi.cc.u32 %3, %9, %12, %3;"
                       "madc.lo.cc.u32 %4, %10, %12, %4; madc.hi.cc.u32 %5, %10, %12, %5;"
                       "madc.lo.cc.u32 %6, %11, %12, %6; madc.hi.u32 %7, %11, %12, %7;"
                       : "+r"(e[0]), "+r"(e[1]), "+r"(e[2]), "+r"(e[3]), "+r"(e[4]), "+r"(e[5]), "+r"(e[6]), "+r"(e[7])
                       : "r"(y), "r"(y ^ 0x55u), "r"(y + 3u), "r"(y * 3u), "r"(m));
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; i++) {
          if (WHICH == 1) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(s[i]) : "r"(s[(i + 1) & 7]), "r"(y));
          else if (WHICH == 2) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(s[i]) : "r"(s[(i + 1) & 7]), "r"(y));
          else asm volatile("add.u32 %0, %0, %1;" : "+r"(s[i]) : "r"(s[(i + 1) & 7]));
        }
      }
    }
  }
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) acc ^= s[i];
  if (acc == 0x12345678u) out[0] = acc;
}

#define CUDA_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) throw Error(std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #x); } while (0)

double imad_microbench(int device, int which) {
  CUDA_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop; CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  uint32_t* d; CUDA_CHECK(cudaMalloc(&d, 4));
  const int iters = 4000, threads = 256, blocks = prop.multiProcessorCount * 8;
  cudaEvent_t e0, e1; CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1));
  auto go = [&](int n) {
    if (which == 0) imad_bench_kernel<0><<<blocks, threads>>>(d, n, 0x9e3779b9u);
    else if (which == 1) imad_bench_kernel<1><<<blocks, threads>>>(d, n, 0x9e3779b9u);
    else if (which == 2) imad_bench_kernel<2><<<blocks, threads>>>(d, n, 0x9e3779b9u);
    else imad_bench_kernel<3><<<blocks, threads>>>(d, n, 0x9e3779b9u);
  };
  go(50);
  CUDA_CHECK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    CUDA_CHECK(cudaEventRecord(e0));
    go(iters);
    CUDA_CHECK(cudaEventRecord(e1));
    CUDA_CHECK(cudaEventSynchronize(e1));
    float ms; CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    best = std::min(best, ms);
  }
  cudaFree(d); cudaEventDestroy(e0); cudaEventDestroy(e1);
  double ops = (double)blocks * threads * (double)iters * 32.0;
  return ops / (best * 1e-3);
}

// ---- engine ------------------------------------------------------------------------------------------
typedef void (*BatchKernel)(const KParams);
static BatchKernel batch_kernel(int T, int V) {
  if (T == 32 && V == 1) return eval_batch_kernel<32, 1>;
  if (T == 32 && V == 2) return eval_batch_kernel<32, 2>;
  if (T == 64 && V == 1) return eval_batch_kernel<64, 1>;
  if (T == 64 && V == 2) return eval_batch_kernel<64, 2>;
  throw Error("GW_THREADS must be 32 or 64 and GW_V 1 or 2");
}

struct Engine::Dev {
  int device = -1;
  int sms = 0, ctas_per_sm = 0;
  uint4* code = nullptr; uint4* consts = nullptr;
  uint4* spill = nullptr; size_t spill_threads = 0;
  // staging for the host-buffer API
  cudaStream_t stream[2] = {nullptr, nullptr};
  uint4* d_in[2] = {nullptr, nullptr}; uint4* d_out[2] = {nullptr, nullptr};
  uint32_t* d_status[2] = {nullptr, nullptr};
  size_t chunk = 0;
  // single-witness latency mode
  uint4* lat_code = nullptr; uint32_t* lat_levels = nullptr; uint4* lat_consts = nullptr;
  uint4* lat_in = nullptr; uint4* lat_out = nullptr; uint32_t* lat_status = nullptr;
  std::mutex mu;
};

static int env_int(const char* name, int dflt) { const char* s = getenv(name); return (s && *s) ? atoi(s) : dflt; }

Engine::Engine(const uint8_t* graph_data, size_t len) {
  graph = deserialize_witnesscalc_graph(graph_data, len);
  threads = env_int("GW_THREADS", 64);
  sets_per_thread = env_int("GW_V", 1);
  batch_kernel(threads, sets_per_thread);      // validates the combination
  PlanOptions opt; opt.n_regs = (uint32_t)env_int("GW_REGS", (int)opt.n_regs);
  opt.div_batch = (uint32_t)env_int("GW_DIV_BATCH", (int)opt.div_batch);
  opt.fuse_dot = env_int("GW_FUSE_DOT", 1) != 0;
  opt.max_terms = (uint32_t)env_int("GW_MAX_TERMS", (int)opt.max_terms);
  opt.pair = env_int("GW_PAIR", 0) != 0 && sets_per_thread == 1;   // bundles are the V == 1 source of ILP
  plan = compile_plan(graph, opt);
}

size_t Engine::smem_bytes() const {
  return (size_t)(threads / 32) * RING * 16 + (size_t)plan.n_regs * 32 * sets_per_thread * threads;
}

Engine::~Engine() {
  for (auto& kv : devs) {
    Dev* d = kv.second;
    cudaSetDevice(d->device);
    cudaFree(d->code); cudaFree(d->consts); cudaFree(d->spill);
    cudaFree(d->lat_code); cudaFree(d->lat_levels); cudaFree(d->lat_consts); cudaFree(d->lat_in); cudaFree(d->lat_out); cudaFree(d->lat_status);
    for (int i = 0; i < 2; i++) { cudaFree(d->d_in[i]); cudaFree(d->d_out[i]); cudaFree(d->d_status[i]); if (d->stream[i]) cudaStreamDestroy(d->stream[i]); }
    delete d;
  }
}

Engine::Dev* Engine::dev(int device) {
  std::lock_guard<std::mutex> lk(mu);
  auto it = devs.find(device);
  if (it != devs.end()) return it->second;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) throw Error("no CUDA device available: this library has no CPU fallback");
  if (device < 0 || device >= ndev) throw Error("CUDA device index out of range");
  CUDA_CHECK(cudaSetDevice(device));
  Dev* d = new Dev();
  d->device = device;
  cudaDeviceProp prop; CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  d->sms = prop.multiProcessorCount;
  const size_t smem = smem_bytes();
  if (smem > (size_t)prop.sharedMemPerBlockOptin) throw Error("register file does not fit shared memory: lower GW_REGS or GW_THREADS");
  BatchKernel k = batch_kernel(threads, sets_per_thread);
  CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d->ctas_per_sm, k, threads, smem));
  if (d->ctas_per_sm < 1) throw Error("kernel cannot be resident with this register-file size");
  CUDA_CHECK(cudaMalloc(&d->code, plan.code.size() * sizeof(Instr)));
  CUDA_CHECK(cudaMemcpy(d->code, plan.code.data(), plan.code.size() * sizeof(Instr), cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMalloc(&d->consts, plan.consts.size() * 32));
  CUDA_CHECK(cudaMemcpy(d->consts, plan.consts.data(), plan.consts.size() * 32, cudaMemcpyHostToDevice));
  d->spill_threads = (size_t)d->sms * d->ctas_per_sm * threads;
  if (plan.n_spill) CUDA_CHECK(cudaMalloc(&d->spill, (size_t)plan.n_spill * 32 * sets_per_thread * d->spill_threads));
  devs[device] = d;
  return d;
}

void Engine::launch(Dev* d, const void* d_inputs, size_t B, void* d_witness, uint32_t* d_status, void* stream) {
  if (B == 0 || plan.code.empty()) return;
  KParams p;
  p.code = d->code; p.n_slots = (uint32_t)plan.code.size(); p.consts = d->consts;
  p.inputs = (const uint4*)d_inputs; p.out = (uint4*)d_witness; p.spill = d->spill; p.status = d_status;
  p.B = B; p.I = plan.n_inputs; p.W = plan.n_witness;
  const size_t tile = (size_t)threads * sets_per_thread;
  size_t n_tiles = (B + tile - 1) / tile;
  if (n_tiles > 0xFFFFFFFFull) throw Error("batch too large");
  p.n_tiles = (uint32_t)n_tiles;
  p.spill_threads = d->spill_threads;
  p.out_wrap = (unsigned long long)env_int("GW_DEBUG_OUT_WRAP", 0);
  int grid = (int)std::min<size_t>(n_tiles, (size_t)d->sms * d->ctas_per_sm);
  batch_kernel(threads, sets_per_thread)<<<grid, threads, smem_bytes(), (cudaStream_t)stream>>>(p);
  CUDA_CHECK(cudaGetLastError());
}

void Engine::run_device(int device, const void* d_inputs, size_t B, void* d_witness, uint32_t* d_status, void* stream) {
  Dev* d = dev(device);
  CUDA_CHECK(cudaSetDevice(device));
  std::lock_guard<std::mutex> lk(d->mu);     // the spill area is shared by all launches on this device
  launch(d, d_inputs, B, d_witness, d_status, stream);
}

// host buffers: chunked, double-buffered H2D -> kernel -> D2H on two streams
void Engine::run_host_on(int device, const uint8_t* inputs, size_t B, uint8_t* witness, uint32_t* status) {
  Dev* d = dev(device);
  CUDA_CHECK(cudaSetDevice(device));
  std::lock_guard<std::mutex> lk(d->mu);
  const size_t in_b = (size_t)plan.n_inputs * 32, out_b = (size_t)plan.n_witness * 32;
  // chunk size: bounded by a device-memory budget per staging buffer (two of them), at most two full
  // waves of resident threads, and small enough to give the copy/compute pipeline >= 4 stages when the
  // batch is large.  GW_CHUNK_MB overrides the budget.
  size_t free_b = 0, total_b = 0;
  CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
  size_t budget = (size_t)env_int("GW_CHUNK_MB", 24576) << 20;
  budget = std::min(budget, (free_b + 2 * d->chunk * (in_b + out_b)) / 5);
  size_t chunk = std::min<size_t>(budget / std::max<size_t>(out_b + in_b, 1), 2 * d->spill_threads * sets_per_thread);
  if (B >= 4 * 2048) chunk = std::min(chunk, (B + 3) / 4);
  chunk = std::max<size_t>(std::min(chunk, B), 1);
  if (chunk > d->chunk) {
    for (int i = 0; i < 2; i++) {
      cudaFree(d->d_in[i]); cudaFree(d->d_out[i]); cudaFree(d->d_status[i]);
      CUDA_CHECK(cudaMalloc(&d->d_in[i], chunk * in_b));
      CUDA_CHECK(cudaMalloc(&d->d_out[i], std::max<size_t>(chunk * out_b, 32)));
      CUDA_CHECK(cudaMalloc(&d->d_status[i], chunk * 4));
      if (!d->stream[i]) CUDA_CHECK(cudaStreamCreateWithFlags(&d->stream[i], cudaStreamNonBlocking));
    }
    d->chunk = chunk;
  }
  // Both streams share the spill area, so kernels of consecutive chunks must not overlap: an event
  // chains kernel k+1 behind kernel k while the copies of the two streams overlap with it.
  cudaEvent_t kdone[2]; CUDA_CHECK(cudaEventCreateWithFlags(&kdone[0], cudaEventDisableTiming)); CUDA_CHECK(cudaEventCreateWithFlags(&kdone[1], cudaEventDisableTiming));
  int k = 0;
  for (size_t off = 0; off < B; off += chunk, k ^= 1) {
    size_t nb = std::min(chunk, B - off);
    cudaStream_t s = d->stream[k];
    CUDA_CHECK(cudaMemcpyAsync(d->d_in[k], inputs + off * in_b, nb * in_b, cudaMemcpyHostToDevice, s));
    if (off) CUDA_CHECK(cudaStreamWaitEvent(s, kdone[k ^ 1], 0));
    launch(d, d->d_in[k], nb, d->d_out[k], status ? d->d_status[k] : nullptr, s);
    CUDA_CHECK(cudaEventRecord(kdone[k], s));
    CUDA_CHECK(cudaMemcpyAsync(witness + off * out_b, d->d_out[k], nb * out_b, cudaMemcpyDeviceToHost, s));
    if (status) CUDA_CHECK(cudaMemcpyAsync(status + off, d->d_status[k], nb * 4, cudaMemcpyDeviceToHost, s));
  }
  CUDA_CHECK(cudaStreamSynchronize(d->stream[0]));
  CUDA_CHECK(cudaStreamSynchronize(d->stream[1]));
  cudaEventDestroy(kdone[0]); cudaEventDestroy(kdone[1]);
}

void Engine::run_host(const uint8_t* inputs, size_t B, uint8_t* witness, uint32_t* status, int n_gpus, int first_device) {
  if (B == 0) return;
  if (n_gpus <= 1) { run_host_on(first_device, inputs, B, witness, status); return; }
  // independent input sets: contiguous shards, one host thread per GPU, no collective
  const size_t in_b = (size_t)plan.n_inputs * 32, out_b = (size_t)plan.n_witness * 32;
  std::vector<std::thread> th;
  std::vector<std::string> errs(n_gpus);
  for (int g = 0; g < n_gpus; g++) {
    size_t lo = B * g / n_gpus, hi = B * (g + 1) / n_gpus;
    if (hi == lo) continue;
    th.emplace_back([=, &errs]() {
      try { run_host_on(first_device + g, inputs + lo * in_b, hi - lo, witness + lo * out_b, status ? status + lo : nullptr); }
      catch (const std::exception& e) { errs[g] = e.what(); }
    });
  }
  for (auto& t : th) t.join();
  for (auto& e : errs) if (!e.empty()) throw Error(e);
}

static const int LAT_THREADS = 128;

// one witness, host buffers: inputs I x 32 B, witness W x 32 B; returns the kernel time in ms if asked
void Engine::run_latency(int device, const uint8_t* inputs, uint8_t* witness, uint32_t* status, float* kernel_ms) {
  Dev* d = dev(device);
  CUDA_CHECK(cudaSetDevice(device));
  std::lock_guard<std::mutex> lk(d->mu);
  {
    std::lock_guard<std::mutex> lk2(mu);
    if (!lat_ready) {
      cudaDeviceProp prop; CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
      lat_plan = compile_latency_plan(graph, (uint32_t)(prop.sharedMemPerBlockOptin / 32));
      lat_ready = true;
    }
  }
  const LatencyPlan& lp = lat_plan;
  const size_t in_b = (size_t)lp.n_inputs * 32, out_b = std::max<size_t>((size_t)lp.n_witness * 32, 32);
  if (!d->lat_code) {
    CUDA_CHECK(cudaMalloc(&d->lat_code, std::max<size_t>(lp.code.size(), 1) * sizeof(Instr)));
    CUDA_CHECK(cudaMemcpy(d->lat_code, lp.code.data(), lp.code.size() * sizeof(Instr), cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMalloc(&d->lat_levels, std::max<size_t>(lp.level_count.size(), 1) * 4));
    CUDA_CHECK(cudaMemcpy(d->lat_levels, lp.level_count.data(), lp.level_count.size() * 4, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMalloc(&d->lat_consts, lp.consts.size() * 32));
    CUDA_CHECK(cudaMemcpy(d->lat_consts, lp.consts.data(), lp.consts.size() * 32, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMalloc(&d->lat_in, in_b));
    CUDA_CHECK(cudaMalloc(&d->lat_out, out_b));
    CUDA_CHECK(cudaMalloc(&d->lat_status, 4));
    CUDA_CHECK(cudaFuncSetAttribute(eval_latency_kernel<LAT_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)lp.n_slots * 32)));
  }
  if (lp.level_count.empty()) return;
  LParams p;
  p.code = d->lat_code; p.level_count = d->lat_levels; p.n_levels = (uint32_t)lp.level_count.size();
  p.consts = d->lat_consts; p.inputs = d->lat_in; p.out = d->lat_out; p.status = status ? d->lat_status : nullptr;
  CUDA_CHECK(cudaMemcpyAsync(d->lat_in, inputs, in_b, cudaMemcpyHostToDevice, 0));
  if (status) CUDA_CHECK(cudaMemsetAsync(d->lat_status, 0, 4, 0));
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (kernel_ms) { CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1)); CUDA_CHECK(cudaEventRecord(e0, 0)); }
  eval_latency_kernel<LAT_THREADS><<<1, LAT_THREADS, (size_t)lp.n_slots * 32, 0>>>(p);
  CUDA_CHECK(cudaGetLastError());
  if (kernel_ms) CUDA_CHECK(cudaEventRecord(e1, 0));
  CUDA_CHECK(cudaMemcpyAsync(witness, d->lat_out, (size_t)lp.n_witness * 32, cudaMemcpyDeviceToHost, 0));
  if (status) CUDA_CHECK(cudaMemcpyAsync(status, d->lat_status, 4, cudaMemcpyDeviceToHost, 0));
  CUDA_CHECK(cudaStreamSynchronize(0));
  if (kernel_ms) { CUDA_CHECK(cudaEventElapsedTime(kernel_ms, e0, e1)); cudaEventDestroy(e0); cudaEventDestroy(e1); }
}

int cuda_device_count() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

}  // namespace gw
