// Microbenchmark: issue rate and dependent latency of FFMA vs FFMA2 (fma.rn.f32x2) on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t fma2(f2_t a, f2_t b, f2_t c) { f2_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

template <int CHAINS, bool PACKED>
__global__ void k(float* out, int iters, long long* cycles, f2_t w2, float w1) {
  f2_t a2[CHAINS]; float a1[CHAINS];
  for (int i = 0; i < CHAINS; i++) { a2[i] = (f2_t)threadIdx.x + i; a1[i] = threadIdx.x + i; }
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int i = 0; i < CHAINS; i++) { if (PACKED) a2[i] = fma2(a2[i], w2, a2[i]); else a1[i] = fma1(a1[i], w1, a1[i]); }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < CHAINS; i++) s += PACKED ? (float)(a2[i] & 0xffff) : a1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int CHAINS, bool PACKED>
void run(int warps) {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  k<CHAINS, PACKED><<<1, warps * 32>>>(out, iters, cyc, 0x3f8000003f800000ull, 1.0f);
  k<CHAINS, PACKED><<<1, warps * 32>>>(out, iters, cyc, 0x3f8000003f800000ull, 1.0f);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double n = (double)iters * 8 * CHAINS;   // instructions per warp
  printf("%s chains=%d warps/SM=%2d: %.2f cycles per warp-instruction, %.2f instr/clk/SM\n", PACKED ? "FFMA2" : "FFMA ", CHAINS, warps,
         h / n, n * warps / h);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<1, false>(1); run<1, true>(1);          // dependent latency
  run<2, true>(1); run<4, true>(1); run<8, true>(1);
  run<8, false>(4); run<8, true>(4);          // one warp per scheduler: issue rate
  run<8, false>(8); run<8, true>(8);
  run<8, false>(16); run<8, true>(16);
  return 0;
}
