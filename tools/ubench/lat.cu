// Single-warp latency / issue-rate microbenchmarks for the instructions the field multiplier is made of.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu && ./lat
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define REP4(x) x x x x
#define REP16(x) REP4(REP4(x))
#define REP64(x) REP4(REP16(x))

// K: which experiment
template <int K>
__global__ void bench(uint32_t* out, long long* cyc, uint32_t y, int iters) {
  uint32_t a0 = threadIdx.x + 1, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, a4 = a0 * 11, a5 = a0 * 13, a6 = a0 * 17, a7 = a0 * 19;
  uint32_t b0 = y, b1 = y + 1, b2 = y + 2, b3 = y + 3;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    if (K == 0) {        // one carry chain: mad.lo.cc / madc.hi.cc pairs, every link depends on the previous carry AND accumulator reuse 8 apart
      REP16(asm volatile("mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1; madc.lo.cc.u32 %2, %8, %10, %2; madc.hi.cc.u32 %3, %8, %10, %3;"
                         "madc.lo.cc.u32 %4, %8, %11, %4; madc.hi.cc.u32 %5, %8, %11, %5; madc.lo.cc.u32 %6, %8, %12, %6; madc.hi.u32 %7, %8, %12, %7;"
                         : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7) : "r"(b0), "r"(b1), "r"(b2), "r"(b3), "r"(y));)
    } else if (K == 1) { // dependent through the accumulator only: x = x*b + x (mad.wide)
      uint64_t x = ((uint64_t)a1 << 32) | a0;
      REP64(asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x) : "r"(b0), "r"(b1));)
      a0 = (uint32_t)x; a1 = (uint32_t)(x >> 32);
    } else if (K == 2) { // add.cc chain (IADD3.X), 8 long, repeated: dependent through carry and registers
      REP16(asm volatile("add.cc.u32 %0, %0, %8; addc.cc.u32 %1, %1, %9; addc.cc.u32 %2, %2, %10; addc.cc.u32 %3, %3, %11;"
                         "addc.cc.u32 %4, %4, %8; addc.cc.u32 %5, %5, %9; addc.cc.u32 %6, %6, %10; addc.u32 %7, %7, %11;"
                         : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7) : "r"(b0), "r"(b1), "r"(b2), "r"(b3));)
    } else if (K == 3) { // mad.lo dependent chain (IMAD latency)
      REP64(asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a0) : "r"(b0), "r"(b1));)
    } else if (K == 4) { // 64 independent mad.wide (issue rate of one warp)
      uint64_t x0 = a0, x1 = a1, x2 = a2, x3 = a3, x4 = a4, x5 = a5, x6 = a6, x7 = a7;
      REP4(asm volatile("mad.wide.u32 %0, %8, %9, %0; mad.wide.u32 %1, %8, %9, %1; mad.wide.u32 %2, %8, %9, %2; mad.wide.u32 %3, %8, %9, %3;"
                         "mad.wide.u32 %4, %8, %9, %4; mad.wide.u32 %5, %8, %9, %5; mad.wide.u32 %6, %8, %9, %6; mad.wide.u32 %7, %8, %9, %7;"
                         "mad.wide.u32 %0, %8, %9, %0; mad.wide.u32 %1, %8, %9, %1; mad.wide.u32 %2, %8, %9, %2; mad.wide.u32 %3, %8, %9, %3;"
                         "mad.wide.u32 %4, %8, %9, %4; mad.wide.u32 %5, %8, %9, %5; mad.wide.u32 %6, %8, %9, %6; mad.wide.u32 %7, %8, %9, %7;"
                         : "+l"(x0), "+l"(x1), "+l"(x2), "+l"(x3), "+l"(x4), "+l"(x5), "+l"(x6), "+l"(x7) : "r"(b0), "r"(b1));)
      a0 = (uint32_t)(x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7);
    } else if (K == 5) { // add dependent chain (IADD3 latency)
      REP64(asm volatile("add.u32 %0, %0, %1;" : "+r"(a0) : "r"(b0));)
    } else if (K == 6) { // two interleaved carry chains cannot be written in PTX (one CC); instead: chain of mad.lo.cc/madc.hi pairs of length 2
      REP16(asm volatile("mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.u32 %1, %8, %9, %1; mad.lo.cc.u32 %2, %8, %10, %2; madc.hi.u32 %3, %8, %10, %3;"
                         "mad.lo.cc.u32 %4, %8, %11, %4; madc.hi.u32 %5, %8, %11, %5; mad.lo.cc.u32 %6, %8, %12, %6; madc.hi.u32 %7, %8, %12, %7;"
                         : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7) : "r"(b0), "r"(b1), "r"(b2), "r"(b3), "r"(y));)
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}

template <int K> void run(const char* name, int ops_per_iter, int warps) {
  uint32_t* out; long long* cyc; cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  bench<K><<<1, 32 * warps>>>(out, cyc, 12345u, 10);
  bench<K><<<1, 32 * warps>>>(out, cyc, 12345u, iters);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-60s warps/SM=%2d  %.2f cycles per op (per warp)\n", name, warps, (double)h / ((double)iters * ops_per_iter));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {1, 2, 4, 8, 16}) {
    run<0>("carry chain of 8 IMAD.WIDE-pairs (4 wide MACs), back to back", 64, w);   // 16 x 4 wide MACs
    run<6>("independent lo.cc/hi pairs (1 wide MAC each, no long chain)", 64, w);
    run<1>("mad.wide dependent through the 64-bit accumulator", 64, w);
    run<4>("mad.wide, 8 independent accumulators", 64, w);
    run<2>("add.cc/addc chain (8 long)", 128, w);
    run<3>("mad.lo dependent chain", 64, w);
    run<5>("add dependent chain", 64, w);
  }
  return 0;
}
