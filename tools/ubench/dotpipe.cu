// Throughput of the OP_DOT multiply-accumulate (64 IMAD.WIDE per term) at the occupancy of eval_batch_kernel (one CTA of
// 512 threads per SM, <= 128 registers), with the accumulators as separate 32-bit registers (ptxas re-pairs them around
// the loop: IMAD.MOV.U32 on the multiplier pipe) and as 64-bit pairs.  Answers: what do those moves cost?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I circom-witnesscalc_b200/csrc -o tools/ubench/dotpipe tools/ubench/dotpipe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "alu.cuh"
using namespace gw;

struct acc32 { uint32_t e[16], o[15], K[9]; };
__device__ __forceinline__ fe fe_from(const uint4 lo, const uint4 hi) { fe r; r.l[0] = lo.x; r.l[1] = lo.y; r.l[2] = lo.z; r.l[3] = lo.w; r.l[4] = hi.x; r.l[5] = hi.y; r.l[6] = hi.z; r.l[7] = hi.w; return r; }
__device__ __forceinline__ void mac32(acc32& A, const uint32_t* a, const uint32_t* b) {
#ifdef __CUDA_ARCH__
  uint32_t* e = A.e; uint32_t* o = A.o; uint32_t* K = A.K;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint32_t bi = b[i];
    const int p = i & 1;
    { const int c0 = i + p;
      e[c0] = ptx_mad_lo_cc(a[p], bi, e[c0]); e[c0 + 1] = ptx_madc_hi_cc(a[p], bi, e[c0 + 1]);
#pragma unroll
      for (int j = p + 2; j < 8; j += 2) { e[i + j] = ptx_madc_lo_cc(a[j], bi, e[i + j]); e[i + j + 1] = ptx_madc_hi_cc(a[j], bi, e[i + j + 1]); }
      K[c0] = ptx_addc(K[c0], 0); }
    { const int q = 1 - p; const int c0 = i + q - 1;
      o[c0] = ptx_mad_lo_cc(a[q], bi, o[c0]); o[c0 + 1] = ptx_madc_hi_cc(a[q], bi, o[c0 + 1]);
#pragma unroll
      for (int j = q + 2; j < 8; j += 2) { o[i + j - 1] = ptx_madc_lo_cc(a[j], bi, o[i + j - 1]); o[i + j] = ptx_madc_hi_cc(a[j], bi, o[i + j]); }
      K[c0 + 1] = ptx_addc(K[c0 + 1], 0); }
  }
#endif
}

// V = 0: 32-bit accumulators; 1: 64-bit pairs (field.cuh dot_acc).  The loop has the three paths of the kernel's term loop.
template <int V>
__global__ void __launch_bounds__(512, 1) bench(uint32_t* out, const uint32_t* kinds, int n_terms, int iters) {
#ifdef __CUDA_ARCH__
  extern __shared__ uint4 sm[];
  const int T = blockDim.x, tid = threadIdx.x;
  uint4* rf = sm + tid;
  for (int r = 0; r < 24; r++) rf[r * T] = make_uint4(tid * 2654435761u + r, r * 40503u + 7, tid ^ (r << 8), 0x01234567u + r);
  __syncthreads();
  uint32_t sink = 0;
  for (int it = 0; it < iters; it++) {
    acc32 A; dot_acc P;
    if (V == 0) { for (int k = 0; k < 16; k++) A.e[k] = 0; for (int k = 0; k < 15; k++) A.o[k] = 0; for (int k = 0; k < 9; k++) A.K[k] = 0; }
    else dot_init(P);
#pragma unroll 1
    for (int t = 0; t < n_terms; t++) {
      const uint32_t kind = __ldg(kinds + t), reg = (t * 5 + it) % 11;
      const fe x = fe_from(rf[(2 * reg) * T], rf[(2 * reg + 1) * T]);
      if (kind == 0) {
        const fe c = fe_from(sm[24 * T + 2 * (t & 7)], sm[24 * T + 2 * (t & 7) + 1]);
        if (V == 0) mac32(A, x.l, c.l); else dot_mac(P, x.l, c.l);
      } else if (kind == 3) {
        if (V == 0) { A.e[0] = ptx_add_cc(A.e[0], x.l[0]); for (int i = 1; i < 8; i++) A.e[i] = ptx_addc_cc(A.e[i], x.l[i]); A.K[0] = ptx_addc(A.K[0], 0); }
        else dot_add256(P, x.l, 0);
      } else {
        if (V == 0) { A.e[8] = ptx_add_cc(A.e[8], x.l[0]); for (int i = 1; i < 8; i++) A.e[8 + i] = ptx_addc_cc(A.e[8 + i], x.l[i]); A.K[8] = ptx_addc(A.K[8], 0); }
        else dot_add256(P, x.l, 8);
      }
    }
    if (V == 0) { for (int k = 0; k < 16; k++) sink ^= A.e[k]; for (int k = 0; k < 15; k++) sink ^= A.o[k]; for (int k = 0; k < 9; k++) sink += A.K[k]; }
    else { const fe r = fe_mont_reduce(P, 3); for (int k = 0; k < 8; k++) sink ^= r.l[k]; }
    rf[(it % 11) * 2 * T] = make_uint4(sink, sink + 1, sink + 2, sink + 3);
  }
  out[blockIdx.x * T + tid] = sink;
#endif
}

template <int V> void run(const char* name, int n_terms, int iters, const uint32_t* d_kinds, uint32_t* out, int sms) {
  const size_t smem = (24 * 512 + 16) * 16;
  cudaFuncSetAttribute(bench<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  bench<V><<<sms, 512, smem>>>(out, d_kinds, n_terms, 2);
  cudaEventRecord(e0);
  bench<V><<<sms, 512, smem>>>(out, d_kinds, n_terms, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double macs = (double)sms * 512 * iters * n_terms * 64;
  printf("%-46s %8.3f ms  %7.3f T wide-MAC/s (terms %d)  %s\n", name, ms, macs / (ms * 1e-3) / 1e12, n_terms, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  uint32_t h[64]; for (int i = 0; i < 64; i++) h[i] = 0;      // all T_MAC
  uint32_t* dk; cudaMalloc(&dk, sizeof h); cudaMemcpy(dk, h, sizeof h, cudaMemcpyHostToDevice);
  uint32_t* out; cudaMalloc(&out, (size_t)prop.multiProcessorCount * 512 * 4);
  for (int nt : {2, 4, 8}) {
    run<0>("dot term loop, 32-bit accumulators (no reduce)", nt, 4000 / nt, dk, out, prop.multiProcessorCount);
    run<1>("dot term loop, 64-bit pairs + mont_reduce", nt, 4000 / nt, dk, out, prop.multiProcessorCount);
  }
  return 0;
}
