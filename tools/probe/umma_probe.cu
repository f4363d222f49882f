// Bring-up probes for the tcgen05 attention kernels (run on a B200: tools/probe/run_probes.sh). Each probe is ONE CTA that
// checks one layout / instruction hypothesis against a CPU reference, so that a wrong guess costs a line of output and
// not a debugging session inside the pipelined kernels:
//   ss64   D[128,N]  = A[128,32] B[N,32]^T      A, B K-major SWIZZLE_64B tiles written by TMA (64-byte rows)     N in {64,96,192}
//   ts     O[128,32] = P[128,KK] V[KK,32]       P bf16 in TMEM (tcgen05.st, split column ranges), V MN-major SWIZZLE_64B by TMA
//   mnA    dQ[128,32]= dS[128q,128k] K[128k,32] dS MN-major SWIZZLE_128B written with st.shared (generic proxy), K MN-major SW64
//   red    throughput of red.global.add.v4.f32 / .v2.bf16x2 on an L2-resident table (dbias accumulation candidates)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../ct_clip_b200/csrc/ptx.cuh"

using namespace ctb;

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_));   \
      exit(2);                                                                                  \
    }                                                                                           \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_enc = nullptr;
static void tmap2d(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch, uint32_t bi, uint32_t bo,
                   CUtensorMapSwizzle swz) {
  if (!g_enc) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    g_enc = (EncodeTiledFn)p;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch};
  cuuint32_t box[2] = {bi, bo};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("cuTensorMapEncodeTiled failed %d\n", (int)r);
    exit(2);
  }
}

static float frand(uint32_t& s) {
  s = s * 1664525u + 1013904223u;
  return ((s >> 8) & 0xFFFF) / 65536.0f - 0.5f;
}
static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

// ------------------------------------------------------------------------------------------------------------------
// ss64: both operands K-major SWIZZLE_64B (rows of 32 bf16 = 64 B), the layout of the per-head slices of q_hat / k_hat
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) probe_ss64(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb,
                                                      float* out, int N) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                 // 128 x 64 B
  uint8_t* sB = smem + 8192;          // N x 64 B
  uint64_t* bar = (uint64_t*)(smem + 8192 + 256 * 64);
  uint32_t* holder = (uint32_t*)(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(holder, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *holder;
  if (tid == 0) {
    mbar_arrive_expect_tx(&bar[0], 128 * 64 + N * 64);
    tma_load_2d(sA, &ta, &bar[0], 0, 0);
    tma_load_2d(sB, &tb, &bar[0], 0, 0);
    mbar_wait(&bar[0], 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc(1, 0, 0, 128, N);
    for (int ks = 0; ks < 2; ks++) {
      const uint64_t ad = umma_smem_desc_sw(smem_u32(sA) + ks * 32, 16, 512, 4);
      const uint64_t bd = umma_smem_desc_sw(smem_u32(sB) + ks * 32, 16, 512, 4);
      umma_bf16(tm, ad, bd, idesc, ks > 0);
    }
    umma_commit(&bar[1]);
  }
  mbar_wait(&bar[1], 0);
  tc_fence_after();
  for (int c = 0; c < N; c += 32) {
    uint32_t v[32];
    tmem_ld_32x32(tm + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; j++)
      if (c + j < N) out[(size_t)tid * N + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 256);
}

static int run_ss64(int N) {
  std::vector<__nv_bfloat16> A(128 * 32), B((size_t)N * 32);
  uint32_t s = 1234u + N;
  for (auto& x : A) x = __float2bfloat16(frand(s));
  for (auto& x : B) x = __float2bfloat16(frand(s));
  __nv_bfloat16 *dA, *dB;
  float* dO;
  CK(cudaMalloc(&dA, A.size() * 2));
  CK(cudaMalloc(&dB, B.size() * 2));
  CK(cudaMalloc(&dO, (size_t)128 * N * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap ta, tb;
  tmap2d(&ta, dA, 32, 128, 64, 32, 128, CU_TENSOR_MAP_SWIZZLE_64B);
  tmap2d(&tb, dB, 32, N, 64, 32, N, CU_TENSOR_MAP_SWIZZLE_64B);
  const int smem = 1024 + 8192 + 256 * 64 + 64;
  CK(cudaFuncSetAttribute(probe_ss64, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_ss64<<<1, 128, smem>>>(ta, tb, dO, N);
  CK(cudaDeviceSynchronize());
  std::vector<float> O((size_t)128 * N);
  CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  for (int m = 0; m < 128; m++)
    for (int n = 0; n < N; n++) {
      float r = 0;
      for (int k = 0; k < 32; k++) r += __bfloat162float(A[m * 32 + k]) * __bfloat162float(B[(size_t)n * 32 + k]);
      maxerr = fmax(maxerr, fabs(r - O[(size_t)m * N + n]));
    }
  printf("probe ss64 N=%d: max abs err %.3e -> %s\n", N, maxerr, maxerr < 1e-3 ? "PASS" : "FAIL");
  return maxerr < 1e-3 ? 0 : 1;
}

// ------------------------------------------------------------------------------------------------------------------
// ts: A = P (bf16) in TMEM written with tcgen05.st; keys [0, KK/2) at column pcol0, keys [KK/2, KK) at column pcol1
//     (the two softmax warps that share a lane quarter each overwrite their own half of the S chunk);
//     B = V [KK keys][32] MN-major SWIZZLE_64B written by TMA.  O = P V.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) probe_ts(const __grid_constant__ CUtensorMap tv, const float* P, float* out, int KK, int pcol0,
                                                    int pcol1, int ocol) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sV = smem;   // KK x 64 B
  uint64_t* bar = (uint64_t*)(smem + 256 * 64);
  uint32_t* holder = (uint32_t*)(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(holder, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *holder;
  if (tid == 0) {
    mbar_arrive_expect_tx(&bar[0], KK * 64);
    tma_load_2d(sV, &tv, &bar[0], 0, 0);
  }
  // every thread: its row of P -> bf16 pairs -> TMEM (16 keys = 8 columns per store)
  const uint32_t lane_base = tm + ((uint32_t)(warp * 32) << 16);
  for (int k0 = 0; k0 < KK; k0 += 16) {
    uint32_t pk[8];
    for (int j = 0; j < 8; j++) pk[j] = pack_bf16x2(P[(size_t)tid * KK + k0 + 2 * j], P[(size_t)tid * KK + k0 + 2 * j + 1]);
    const int half = KK / 2;
    const uint32_t col = (k0 < half) ? (pcol0 + k0 / 2) : (pcol1 + (k0 - half) / 2);
    tmem_st_32x8(lane_base + col, pk);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    mbar_wait(&bar[0], 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc(1, 0, 1, 128, 32);   // A K-major (TMEM), B MN-major
    const int half = KK / 2;
    for (int ks = 0; ks < KK / 16; ks++) {
      const int k0 = ks * 16;
      const uint32_t acol = (k0 < half) ? (pcol0 + k0 / 2) : (pcol1 + (k0 - half) / 2);
      const uint64_t bd = umma_smem_desc_sw(smem_u32(sV) + ks * 1024, 512, 512, 4);
      umma_bf16_ts(tm + ocol, tm + acol, bd, idesc, ks > 0);
    }
    umma_commit(&bar[1]);
  }
  mbar_wait(&bar[1], 0);
  tc_fence_after();
  uint32_t v[32];
  tmem_ld_32x32(lane_base + ocol, v);
  tmem_ld_wait();
  for (int j = 0; j < 32; j++) out[tid * 32 + j] = __uint_as_float(v[j]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

static int run_ts(int KK, int pcol0, int pcol1, int ocol) {
  std::vector<float> P((size_t)128 * KK);
  std::vector<__nv_bfloat16> V((size_t)KK * 32);
  uint32_t s = 99u + KK;
  for (auto& x : P) x = frand(s) + 0.5f;
  for (auto& x : V) x = __float2bfloat16(frand(s));
  float *dP, *dO;
  __nv_bfloat16* dV;
  CK(cudaMalloc(&dP, P.size() * 4));
  CK(cudaMalloc(&dV, V.size() * 2));
  CK(cudaMalloc(&dO, 128 * 32 * 4));
  CK(cudaMemcpy(dP, P.data(), P.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dV, V.data(), V.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap tv;
  tmap2d(&tv, dV, 32, KK, 64, 32, KK, CU_TENSOR_MAP_SWIZZLE_64B);
  const int smem = 1024 + 256 * 64 + 64;
  CK(cudaFuncSetAttribute(probe_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_ts<<<1, 128, smem>>>(tv, dP, dO, KK, pcol0, pcol1, ocol);
  CK(cudaDeviceSynchronize());
  std::vector<float> O(128 * 32);
  CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  for (int m = 0; m < 128; m++)
    for (int d = 0; d < 32; d++) {
      float r = 0;
      for (int k = 0; k < KK; k++) r += bf(P[(size_t)m * KK + k]) * __bfloat162float(V[(size_t)k * 32 + d]);
      maxerr = fmax(maxerr, fabs(r - O[m * 32 + d]));
    }
  printf("probe ts KK=%d pcol=(%d,%d) ocol=%d: max abs err %.3e -> %s\n", KK, pcol0, pcol1, ocol, maxerr, maxerr < 2e-3 ? "PASS" : "FAIL");
  return maxerr < 2e-3 ? 0 : 1;
}

// ------------------------------------------------------------------------------------------------------------------
// mnA: A = dS as an MN-major SWIZZLE_128B tile written with plain st.shared by the thread that owns key row kk:
//      [2 groups of 64 queries][128 key rows][128 B], 16-byte chunk c of a row at chunk position c ^ (kk & 7);
//      B = K [128 keys][32] MN-major SWIZZLE_64B by TMA.   dQ[q, d] = sum_k dS[q, k] K[k, d]
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) probe_mna(const __grid_constant__ CUtensorMap tk, const float* dS /* [128 k][128 q] */, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;             // 2 x 16 KB
  uint8_t* sK = smem + 32768;     // 128 x 64 B
  uint64_t* bar = (uint64_t*)(smem + 32768 + 8192);
  uint32_t* holder = (uint32_t*)(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(holder, 32);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *holder;
  if (tid == 0) {
    mbar_arrive_expect_tx(&bar[0], 128 * 64);
    tma_load_2d(sK, &tk, &bar[0], 0, 0);
  }
  {   // thread = key row kk
    const int kk = tid;
    for (int g = 0; g < 2; g++)
      for (int c = 0; c < 8; c++) {
        uint32_t w[4];
        for (int j = 0; j < 4; j++) {
          const int q = g * 64 + c * 8 + 2 * j;
          w[j] = pack_bf16x2(dS[(size_t)kk * 128 + q], dS[(size_t)kk * 128 + q + 1]);
        }
        *reinterpret_cast<uint4*>(sA + g * 16384 + kk * 128 + ((c ^ (kk & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    mbar_wait(&bar[0], 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc(1, 1, 1, 128, 32);   // A MN-major, B MN-major
    for (int ks = 0; ks < 8; ks++) {
      const uint64_t ad = umma_smem_desc_sw(smem_u32(sA) + ks * 2048, 16384, 1024, 2);
      const uint64_t bd = umma_smem_desc_sw(smem_u32(sK) + ks * 1024, 512, 512, 4);
      umma_bf16(tm, ad, bd, idesc, ks > 0);
    }
    umma_commit(&bar[1]);
  }
  mbar_wait(&bar[1], 0);
  tc_fence_after();
  uint32_t v[32];
  tmem_ld_32x32(tm + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld_wait();
  for (int j = 0; j < 32; j++) out[tid * 32 + j] = __uint_as_float(v[j]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 32);
}

static int run_mna() {
  std::vector<float> dS(128 * 128);
  std::vector<__nv_bfloat16> K(128 * 32);
  uint32_t s = 4242u;
  for (auto& x : dS) x = frand(s);
  for (auto& x : K) x = __float2bfloat16(frand(s));
  float *dD, *dO;
  __nv_bfloat16* dK;
  CK(cudaMalloc(&dD, dS.size() * 4));
  CK(cudaMalloc(&dK, K.size() * 2));
  CK(cudaMalloc(&dO, 128 * 32 * 4));
  CK(cudaMemcpy(dD, dS.data(), dS.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dK, K.data(), K.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap tk;
  tmap2d(&tk, dK, 32, 128, 64, 32, 128, CU_TENSOR_MAP_SWIZZLE_64B);
  const int smem = 1024 + 32768 + 8192 + 64;
  CK(cudaFuncSetAttribute(probe_mna, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_mna<<<1, 128, smem>>>(tk, dD, dO);
  CK(cudaDeviceSynchronize());
  std::vector<float> O(128 * 32);
  CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  for (int q = 0; q < 128; q++)
    for (int d = 0; d < 32; d++) {
      float r = 0;
      for (int k = 0; k < 128; k++) r += bf(dS[(size_t)k * 128 + q]) * __bfloat162float(K[k * 32 + d]);
      maxerr = fmax(maxerr, fabs(r - O[q * 32 + d]));
    }
  printf("probe mnA: max abs err %.3e -> %s\n", maxerr, maxerr < 2e-3 ? "PASS" : "FAIL");
  return maxerr < 2e-3 ? 0 : 1;
}

// ------------------------------------------------------------------------------------------------------------------
// red: how fast can 148 x 8 CTAs push 1.3 MB-per-pass of fp32 / bf16x2 reductions into an L2-resident table?
// ------------------------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) probe_red(float* table, long long n_f32, int passes) {
  // MODE 0: red.global.add.v4.f32 (16 B = 4 values), MODE 1: red.global.add.noftz.v4.bf16x2 (16 B = 8 values), MODE 2: plain st.v4 (reference)
  const long long vecs = n_f32 / 4;
  for (int p = 0; p < passes; p++) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < vecs; i += (long long)gridDim.x * blockDim.x) {
      float* a = table + i * 4;
      if (MODE == 0) {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a), "f"(1.0f), "f"(2.0f), "f"(3.0f), "f"(4.0f) : "memory");
      } else if (MODE == 1) {
        asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(a), "r"(0x3f803f80u), "r"(0x3f803f80u), "r"(0x3f803f80u),
                     "r"(0x3f803f80u)
                     : "memory");
      } else {
        *reinterpret_cast<float4*>(a) = make_float4(1.f, 2.f, 3.f, 4.f);
      }
    }
  }
}
static int run_red() {
  const long long n = 8LL * 576 * 576;   // one layer's dbias table: 10.6 MB fp32
  float* t;
  CK(cudaMalloc(&t, n * 4));
  CK(cudaMemset(t, 0, n * 4));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int passes = 24;   // 24 x 10.6 MB = 255 MB of reduction payload
  for (int mode = 0; mode < 3; mode++) {
    for (int rep = 0; rep < 2; rep++) {
      CK(cudaEventRecord(e0));
      if (mode == 0) probe_red<0><<<148 * 8, 256>>>(t, n, passes);
      else if (mode == 1) probe_red<1><<<148 * 8, 256>>>(t, n, passes);
      else probe_red<2><<<148 * 8, 256>>>(t, n, passes);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep == 1)
        printf("probe red mode %d (%s): %.3f ms for %.1f MB payload -> %.2f TB/s payload, %.2f G 16B-ops/s\n", mode,
               mode == 0 ? "red.v4.f32" : (mode == 1 ? "red.v4.bf16x2" : "st.v4"), ms, passes * n * 4 / 1e6, passes * n * 4 / ms / 1e9,
               passes * (n / 4) / ms / 1e6);
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// diagnosis helpers: (bmn) SS with A K-major SW64 [128][32] and B = V [32 keys][32 d] MN-major SW64 (isolates the MN-major
// SWIZZLE_64B descriptor); (tsk) TS with P[128][32] in TMEM and B = W [32 n][32 k] K-major SW64 (isolates tcgen05.st + TS)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) probe_diag(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb,
                                                      const float* P, float* out, int mode) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + 8192;
  uint64_t* bar = (uint64_t*)(smem + 8192 + 2048);
  uint32_t* holder = (uint32_t*)(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(holder, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *holder;
  const uint32_t lane_base = tm + ((uint32_t)(warp * 32) << 16);
  if (tid == 0) {
    mbar_arrive_expect_tx(&bar[0], (mode == 0 ? 128 * 64 : 0) + 32 * 64);
    if (mode == 0) tma_load_2d(sA, &ta, &bar[0], 0, 0);
    tma_load_2d(sB, &tb, &bar[0], 0, 0);
  }
  if (mode == 1) {
    for (int k0 = 0; k0 < 32; k0 += 16) {
      uint32_t pk[8];
      for (int j = 0; j < 8; j++) pk[j] = pack_bf16x2(P[tid * 32 + k0 + 2 * j], P[tid * 32 + k0 + 2 * j + 1]);
      tmem_st_32x8(lane_base + 32 + k0 / 2, pk);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    mbar_wait(&bar[0], 0);
    tc_fence_after();
    if (mode == 0) {
      const uint32_t idesc = umma_idesc(1, 0, 1, 128, 32);
      for (int ks = 0; ks < 2; ks++)
        umma_bf16(tm, umma_smem_desc_sw(smem_u32(sA) + ks * 32, 16, 512, 4), umma_smem_desc_sw(smem_u32(sB) + ks * 1024, 512, 512, 4), idesc, ks > 0);
    } else {
      const uint32_t idesc = umma_idesc(1, 0, 0, 128, 32);
      for (int ks = 0; ks < 2; ks++)
        umma_bf16_ts(tm, tm + 32 + ks * 8, umma_smem_desc_sw(smem_u32(sB) + ks * 32, 16, 512, 4), idesc, ks > 0);
    }
    umma_commit(&bar[1]);
  }
  mbar_wait(&bar[1], 0);
  tc_fence_after();
  uint32_t v[32];
  tmem_ld_32x32(lane_base, v);
  tmem_ld_wait();
  for (int j = 0; j < 32; j++) out[tid * 32 + j] = __uint_as_float(v[j]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 64);
}
static int run_diag(int mode) {
  std::vector<float> P(128 * 32);
  std::vector<__nv_bfloat16> A(128 * 32), B(32 * 32);
  uint32_t s = 777u + mode;
  for (auto& x : P) x = frand(s);
  for (size_t i = 0; i < A.size(); i++) A[i] = __float2bfloat16(P[i]);
  for (auto& x : B) x = __float2bfloat16(frand(s));
  float *dP, *dO;
  __nv_bfloat16 *dA, *dB;
  CK(cudaMalloc(&dP, P.size() * 4));
  CK(cudaMalloc(&dA, A.size() * 2));
  CK(cudaMalloc(&dB, B.size() * 2));
  CK(cudaMalloc(&dO, 128 * 32 * 4));
  CK(cudaMemcpy(dP, P.data(), P.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap ta, tb;
  tmap2d(&ta, dA, 32, 128, 64, 32, 128, CU_TENSOR_MAP_SWIZZLE_64B);
  tmap2d(&tb, dB, 32, 32, 64, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
  const int smem = 1024 + 8192 + 2048 + 64;
  CK(cudaFuncSetAttribute(probe_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_diag<<<1, 128, smem>>>(ta, tb, dP, dO, mode);
  CK(cudaDeviceSynchronize());
  std::vector<float> O(128 * 32);
  CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  for (int m = 0; m < 128; m++)
    for (int n = 0; n < 32; n++) {
      float r = 0;
      for (int k = 0; k < 32; k++)
        r += bf(P[m * 32 + k]) * (mode == 0 ? __bfloat162float(B[k * 32 + n]) /* V[k][n] */ : __bfloat162float(B[n * 32 + k]) /* W[n][k] */);
      maxerr = fmax(maxerr, fabs(r - O[m * 32 + n]));
    }
  printf("probe diag mode %d (%s): max abs err %.3e -> %s\n", mode, mode == 0 ? "SS, B MN-major SW64" : "TS, B K-major SW64", maxerr,
         maxerr < 2e-3 ? "PASS" : "FAIL");
  return maxerr < 2e-3 ? 0 : 1;
}

// ------------------------------------------------------------------------------------------------------------------
// ns16: K = 16 operands in the NO-SWIZZLE K-major canonical layout (8-row x 16-byte core matrices; LBO = 128 B between the two
// K halves, SBO = 256 B between 8-row groups), written with plain st.shared: the "augmented k-step" tiles of the backward
// kernel (A = ones, B = (-lse/sc2) split into three bf16 terms).  D[128, 64] (+)= A[128,16] B[64,16]^T, accumulating onto a
// first SW64 MMA so that the accumulate path is the one the kernel uses.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) probe_ns16(const float* A, const float* B, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;            // 16 groups x 256 B
  uint8_t* sB = smem + 4096;     // 8 groups x 256 B
  uint64_t* bar = (uint64_t*)(smem + 8192);
  uint32_t* holder = (uint32_t*)(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(holder, 64);
    tmem_relinquish();
  }
  auto put = [](uint8_t* tile, int row, const float* src) {   // 16 values of one row -> two 16-byte core-matrix rows
    for (int kh = 0; kh < 2; kh++) {
      uint32_t w[4];
      for (int j = 0; j < 4; j++) w[j] = pack_bf16x2(src[kh * 8 + 2 * j], src[kh * 8 + 2 * j + 1]);
      *reinterpret_cast<uint4*>(tile + (row >> 3) * 256 + kh * 128 + (row & 7) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  };
  put(sA, tid, A + tid * 16);
  if (tid < 64) put(sB, tid, B + tid * 16);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *holder;
  if (tid == 0) {
    const uint32_t idesc = umma_idesc(1, 0, 0, 128, 64);
    const uint64_t ad = umma_smem_desc_sw(smem_u32(sA), 128, 256, 0);
    const uint64_t bd = umma_smem_desc_sw(smem_u32(sB), 128, 256, 0);
    umma_bf16(tm, ad, bd, idesc, 0);
    umma_commit(&bar[0]);
  }
  mbar_wait(&bar[0], 0);
  tc_fence_after();
  for (int c = 0; c < 64; c += 32) {
    uint32_t v[32];
    tmem_ld_32x32(tm + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; j++) out[tid * 64 + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 64);
}
static int run_ns16() {
  std::vector<float> A(128 * 16), B(64 * 16);
  uint32_t s = 31337u;
  for (auto& x : A) x = frand(s);
  for (auto& x : B) x = frand(s);
  float *dA, *dB, *dO;
  CK(cudaMalloc(&dA, A.size() * 4));
  CK(cudaMalloc(&dB, B.size() * 4));
  CK(cudaMalloc(&dO, 128 * 64 * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  const int smem = 1024 + 8192 + 64;
  CK(cudaFuncSetAttribute(probe_ns16, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_ns16<<<1, 128, smem>>>(dA, dB, dO);
  CK(cudaDeviceSynchronize());
  std::vector<float> O(128 * 64);
  CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  for (int m = 0; m < 128; m++)
    for (int n = 0; n < 64; n++) {
      float r = 0;
      for (int k = 0; k < 16; k++) r += bf(A[m * 16 + k]) * bf(B[n * 16 + k]);
      maxerr = fmax(maxerr, fabs(r - O[m * 64 + n]));
    }
  printf("probe ns16 (no-swizzle K-major, K = 16): max abs err %.3e -> %s\n", maxerr, maxerr < 1e-3 ? "PASS" : "FAIL");
  return maxerr < 1e-3 ? 0 : 1;
}

int main(int argc, char** argv) {
  const char* which = argc > 1 ? argv[1] : "all";
  int rc = 0;
  if (!strcmp(which, "ss64")) rc = run_ss64(argc > 2 ? atoi(argv[2]) : 96);
  else if (!strcmp(which, "ts")) rc = run_ts(argc > 2 ? atoi(argv[2]) : 96, argc > 3 ? atoi(argv[3]) : 0, argc > 4 ? atoi(argv[4]) : 48, argc > 5 ? atoi(argv[5]) : 96);
  else if (!strcmp(which, "mna")) rc = run_mna();
  else if (!strcmp(which, "red")) rc = run_red();
  else if (!strcmp(which, "diag")) rc = run_diag(argc > 2 ? atoi(argv[2]) : 0);
  else if (!strcmp(which, "ns16")) rc = run_ns16();
  else printf("usage: umma_probe ss64 N | ts KK pcol0 pcol1 ocol | mna | red\n");
  return rc;
}
