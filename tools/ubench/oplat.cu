// Single-warp latency of the field operations the latency-mode kernel is made of (cycles per dependent op),
// with 1 and 32 active lanes, plus the cost of a CTA barrier.  Build + run on a B200:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -DGW_CHAINS -I circom-witnesscalc_b200/csrc -o tools/ubench/oplat tools/ubench/oplat.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "alu.cuh"
using namespace gw;

template <int K>
__global__ void bench(uint32_t* out, long long* cyc, int iters, int lanes) {
  fe x, y, z;
  for (int i = 0; i < 8; i++) { x.l[i] = 0x1234567u * (i + 1) + threadIdx.x; y.l[i] = 0x7654321u * (i + 3); z.l[i] = 0x3333333u * (i + 7); }
  x.l[7] &= 0x0FFFFFFF; y.l[7] &= 0x0FFFFFFF; z.l[7] &= 0x0FFFFFFF;
  __shared__ uint4 sm[64];
  long long t0 = 0, t1 = 0;
  if ((threadIdx.x & 31) < lanes) {
    t0 = clock64();
    for (int it = 0; it < iters; it++) {
      if (K == 0) x = fe_mul(x, y);
      else if (K == 1) x = fe_sqr(x);
      else if (K == 2) { dot_acc P; dot_init(P); dot_mac(P, x.l, y.l); dot_mac(P, y.l, z.l); dot_mac(P, z.l, x.l); x = fe_mont_reduce(P, 2); }
      else if (K == 3) x = fe_inv(x);
      else if (K == 4) x = fe_add(x, y);
      else if (K == 5) { dot_acc P; dot_init(P); dot_mac(P, x.l, y.l); dot_add256(P, z.l, 8); x = fe_mont_reduce(P, 2); }
      else if (K == 6) { sm[threadIdx.x & 31] = make_uint4(x.l[0], x.l[1], x.l[2], x.l[3]); __syncwarp(); uint4 v = sm[(threadIdx.x + 1) & 31]; x.l[0] ^= v.x; x.l[1] += v.y; }
      else if (K == 7) x = fe_mont_mul(x, y);
    }
    t1 = clock64();
  }
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  uint32_t a = 0; for (int i = 0; i < 8; i++) a ^= x.l[i];
  out[threadIdx.x] = a;
}
__global__ void bar_bench(long long* cyc, int iters) {
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int K> void run(const char* name, int iters) {
  uint32_t* out; long long* cyc; cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 8);
  for (int lanes : {1, 32}) {
    bench<K><<<1, 32>>>(out, cyc, 3, lanes);
    bench<K><<<1, 32>>>(out, cyc, iters, lanes);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s lanes=%2d  %9.1f cycles per op\n", name, lanes, (double)h / iters);
  }
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("fe_mul (schoolbook + Barrett)", 2000);
  run<1>("fe_sqr", 2000);
  run<7>("fe_mont_mul (CIOS)", 2000);
  run<2>("dot: 3 MAC + Montgomery reduce", 2000);
  run<5>("dot: 1 MAC + 1 ADDHI + reduce", 2000);
  run<4>("fe_add", 2000);
  run<6>("smem store + syncwarp + load", 2000);
  run<3>("fe_inv (safegcd 20x30)", 200);
  long long* cyc; cudaMalloc(&cyc, 8);
  for (int t : {128, 256, 512}) {
    bar_bench<<<1, t>>>(cyc, 10); bar_bench<<<1, t>>>(cyc, 4000);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("__syncthreads, %3d threads: %.1f cycles\n", t, (double)h / 4000);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
