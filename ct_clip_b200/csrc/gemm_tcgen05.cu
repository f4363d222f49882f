// Persistent warp-specialised GEMM for sm_100a: TMA (SWIZZLE_128B) -> shared memory ring ->
// tcgen05.mma (cta_group::1, M=128, N=BN, K=16, bf16 in / fp32 accumulate in TMEM, two
// accumulator stages) -> tcgen05.ld epilogue fused with bias / residual / GEGLU / split-K
// reduction / argmax.
//
// Roles (320 threads): warp 0 = TMA producer (1 lane), warp 1 = MMA issuer (1 lane),
// warps 2..9 = epilogue (warp 2 also owns the TMEM allocation). Epilogue warp w reads TMEM lanes
// [32*(w%4), 32*(w%4)+32), i.e. one output row per thread; the two warps sharing a lane quarter take
// alternate 32-column chunks (the fused epilogues, not the MMA, bound the K=512 GEMMs otherwise).
//
// Reference ops replaced: see include/ctclip_b200.h (ctclip_gemm_bf16).
#include <stdlib.h>
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/ctclip_b200.h"

namespace ctb {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int EPI_WARPS = 8;    // generic kernel; the EPW = 16 instantiation serves the register-light fused epilogues
constexpr int GEMM_THREADS = (2 + EPI_WARPS) * 32;
constexpr int EPW_LIGHT = 12;   // epilogue warps of the register-light variant: 14 warps -> 128 registers / thread

enum { EPI_BF16 = 0, EPI_F32 = 1, EPI_RESID_F32 = 2, EPI_GEGLU = 3, EPI_ATOMIC_F32 = 4, EPI_ARGMAX = 5, EPI_L2NORM = 6, EPI_BIAS_GELU = 7,
       EPI_GEGLU_BWD = 8 };

struct GemmKParams {
  int M, N, K;
  int m_blks, n_blks, k_blks;
  int splits, k_blks_per_split;
  int n_per_unit;  // 1, or n_blks (argmax: one CTA sweeps all N tiles of its M tile)
  int n_groups;    // n_blks / n_per_unit
  int num_units;
  int epi;
  void* C;
  long long ldc;
  const float* bias;
  const float* resid;
  long long ldr;
  void* C2;
  long long ldc2;
  int* arg_out;
  float* argval_out;
  int* arg2_out;   // ARGMAX: optional runner-up index per row (fp32 re-ranking of the bf16 top-2, ctclip_vq_rerank)
  int norm_cols;
  const float* norm_scale;
  int fast_store;  // all output / residual rows are 16-byte aligned: staged, fully coalesced epilogue stores
  float* colsum;   // GEGLU_BWD: global column sums of the result (2N floats), accumulated through shared memory
  int fast_epi;    // use the specialised epilogue loops (fast_store && N % 32 == 0 && not ARGMAX / ATOMIC)
};

template <int BN, int EPW = EPI_WARPS>
struct GemmCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;  // 128 / 256 / 512
  static constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + 256 /*barriers*/ + 1024 /*argmax merge*/ +
                                    EPW * 2048 /*epilogue store staging*/ + 1024 /*align*/;
};

__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* dst, const float (&v)[32]) {
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint4 u;
    u.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
    u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
    u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
    u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
    d4[i] = u;
  }
}

// ---- epilogue store staging -------------------------------------------------------------------------------------
// After tcgen05.ld every thread owns ONE row (32 consecutive columns). Storing that directly makes each warp-wide
// 16-byte store touch 32 different rows = 32 half-filled 32-byte sectors, and the L2 write path (not the MMA) bounds
// the K=512 GEMMs. Each epilogue warp therefore bounces its 32 x 64 B block through a private 2 KB shared-memory
// buffer (16-byte pieces XOR-swizzled so that both directions are bank-conflict free) and writes it back with 4
// consecutive lanes covering 64 contiguous bytes of one row: full sectors, 8 rows per instruction.
__device__ __forceinline__ void stage_put(uint8_t* st, int lane, int piece, const uint4& u) {
  *reinterpret_cast<uint4*>(st + lane * 64 + ((piece ^ ((lane >> 1) & 3)) << 4)) = u;
}
__device__ __forceinline__ uint4 stage_get(const uint8_t* st, int r, int piece) {
  return *reinterpret_cast<const uint4*>(st + r * 64 + ((piece ^ ((r >> 1) & 3)) << 4));
}
// 32 rows x 32 bf16: base -> element (first row of the warp, col0); ld in elements
__device__ __forceinline__ void staged_store_bf16(uint8_t* st, int lane, const float (&v)[32], __nv_bfloat16* base,
                                                  long long ld, int rows_valid) {
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint4 u;
    u.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
    u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
    u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
    u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
    stage_put(st, lane, i, u);
  }
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 4; it++) {
    const int r = it * 8 + (lane >> 2), pc = lane & 3;
    const uint4 u = stage_get(st, r, pc);
    if (r < rows_valid) *reinterpret_cast<uint4*>(base + (long long)r * ld + pc * 8) = u;
  }
  __syncwarp();
}
// 32 rows x 16 bf16 (32 B per row): two lanes per row
__device__ __forceinline__ void staged_store_bf16_half(uint8_t* st, int lane, const float (&g)[16], __nv_bfloat16* base,
                                                       long long ld, int rows_valid) {
#pragma unroll
  for (int i = 0; i < 2; i++) {
    uint4 u;
    u.x = pack_bf16x2(g[8 * i + 0], g[8 * i + 1]);
    u.y = pack_bf16x2(g[8 * i + 2], g[8 * i + 3]);
    u.z = pack_bf16x2(g[8 * i + 4], g[8 * i + 5]);
    u.w = pack_bf16x2(g[8 * i + 6], g[8 * i + 7]);
    stage_put(st, lane, i, u);
  }
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 2; it++) {
    const int r = it * 16 + (lane >> 1), pc = lane & 1;
    const uint4 u = stage_get(st, r, pc);
    if (r < rows_valid) *reinterpret_cast<uint4*>(base + (long long)r * ld + pc * 8) = u;
  }
  __syncwarp();
}
// 32 rows x 32 fp32 (+ optional residual), in two 16-column halves. rbuf: residual values already loaded in the
// transposed ownership (rbuf[h*4 + it] belongs to row it*8 + lane/4, columns h*16 + (lane%4)*4 ..), or nullptr to load here.
__device__ __forceinline__ void staged_store_f32(uint8_t* st, int lane, const float (&v)[32], float* base, long long ld,
                                                 const float* rbase, long long ldr, const float4* rbuf, int rows_valid) {
#pragma unroll
  for (int h = 0; h < 2; h++) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      uint4 u;
      u.x = __float_as_uint(v[16 * h + 4 * i + 0]);
      u.y = __float_as_uint(v[16 * h + 4 * i + 1]);
      u.z = __float_as_uint(v[16 * h + 4 * i + 2]);
      u.w = __float_as_uint(v[16 * h + 4 * i + 3]);
      stage_put(st, lane, i, u);
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 4; it++) {
      const int r = it * 8 + (lane >> 2), pc = lane & 3;
      const uint4 u = stage_get(st, r, pc);
      float4 o = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
      if (r < rows_valid) {
        if (rbase != nullptr) {
          const float4 rr = (rbuf != nullptr) ? rbuf[h * 4 + it]
                                              : *reinterpret_cast<const float4*>(rbase + (long long)r * ldr + h * 16 + pc * 4);
          o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
        }
        *reinterpret_cast<float4*>(base + (long long)r * ld + h * 16 + pc * 4) = o;
      }
    }
    __syncwarp();
  }
}
// residual prefetch in the transposed ownership used by staged_store_f32
__device__ __forceinline__ void resid_prefetch(float4 (&buf)[8], const float* rbase, long long ldr, int lane, int rows_valid) {
#pragma unroll
  for (int h = 0; h < 2; h++)
#pragma unroll
    for (int it = 0; it < 4; it++) {
      const int r = it * 8 + (lane >> 2), pc = lane & 3;
      buf[h * 4 + it] = (r < rows_valid) ? *reinterpret_cast<const float4*>(rbase + (long long)r * ldr + h * 16 + pc * 4)
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// ---- fast epilogue (v2) ------------------------------------------------------------------------------------------
// One specialised loop per epilogue kind (compile-time EPI), entered once per kernel: no per-chunk dispatch, explicit
// shared-space staging (st/ld.shared with 32-bit addresses; the generic-pointer version compiled to generic LD/ST), the
// TMEM load of chunk i+1 in flight while chunk i is processed, and the bias row of a chunk fetched ONE chunk ahead by a
// single coalesced load per warp and broadcast through 128 B of shared memory (v1 issued eight dependent
// ld.global.nc.v4 per thread right before their first use; that long-scoreboard stall was the top stall of the K=512
// GEMMs).
constexpr int EPI_NONE_DEBUG = 9;   // probe: drain the accumulator without storing (mainloop ceiling)

__device__ __forceinline__ void sts128(uint32_t addr, const uint4& u) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 u;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(addr) : "memory");
  return u;
}
__device__ __forceinline__ void sts32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ uint32_t stg_put_addr(uint32_t st, int lane, int piece) {
  return st + lane * 64 + ((piece ^ ((lane >> 1) & 3)) << 4);
}
__device__ __forceinline__ uint32_t stg_get_addr(uint32_t st, int r, int piece) {
  return st + r * 64 + ((piece ^ ((r >> 1) & 3)) << 4);
}
// make the compiler treat the asynchronously written tcgen05.ld destination registers as redefined here
__device__ __forceinline__ void reg_fence32(uint32_t (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 32; i++) asm volatile("" : "+r"(v[i]));
}
// 32 rows x 32 bf16 (64 B per row) from one row per thread to 4 lanes per row
__device__ __forceinline__ void fstore_bf16(uint32_t st, int lane, const float (&v)[32], __nv_bfloat16* base, long long ld,
                                            int rows_valid) {
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint4 u;
    u.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
    u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
    u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
    u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
    sts128(stg_put_addr(st, lane, i), u);
  }
  __syncwarp();
  __nv_bfloat16* dst = base + (long long)(lane >> 2) * ld + (lane & 3) * 8;
#pragma unroll
  for (int it = 0; it < 4; it++) {
    const int r = it * 8 + (lane >> 2);
    const uint4 u = lds128(stg_get_addr(st, r, lane & 3));
    if (r < rows_valid) *reinterpret_cast<uint4*>(dst) = u;
    dst += 8 * ld;
  }
  __syncwarp();
}
// 32 rows x 16 bf16 (32 B per row): two lanes per row
__device__ __forceinline__ void fstore_bf16_half(uint32_t st, int lane, const float (&g)[16], __nv_bfloat16* base,
                                                 long long ld, int rows_valid) {
#pragma unroll
  for (int i = 0; i < 2; i++) {
    uint4 u;
    u.x = pack_bf16x2(g[8 * i + 0], g[8 * i + 1]);
    u.y = pack_bf16x2(g[8 * i + 2], g[8 * i + 3]);
    u.z = pack_bf16x2(g[8 * i + 4], g[8 * i + 5]);
    u.w = pack_bf16x2(g[8 * i + 6], g[8 * i + 7]);
    sts128(stg_put_addr(st, lane, i), u);
  }
  __syncwarp();
  __nv_bfloat16* dst = base + (long long)(lane >> 1) * ld + (lane & 1) * 8;
#pragma unroll
  for (int it = 0; it < 2; it++) {
    const int r = it * 16 + (lane >> 1);
    const uint4 u = lds128(stg_get_addr(st, r, lane & 1));
    if (r < rows_valid) *reinterpret_cast<uint4*>(dst) = u;
    dst += 16 * ld;
  }
  __syncwarp();
}
// 32 rows x 32 fp32 in two 16-column halves, optional residual already prefetched in the transposed ownership
template <bool RESID>
__device__ __forceinline__ void fstore_f32(uint32_t st, int lane, const float (&v)[32], float* base, long long ld,
                                           const float4 (&rbuf)[8], int rows_valid) {
#pragma unroll
  for (int h = 0; h < 2; h++) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      uint4 u;
      u.x = __float_as_uint(v[16 * h + 4 * i + 0]);
      u.y = __float_as_uint(v[16 * h + 4 * i + 1]);
      u.z = __float_as_uint(v[16 * h + 4 * i + 2]);
      u.w = __float_as_uint(v[16 * h + 4 * i + 3]);
      sts128(stg_put_addr(st, lane, i), u);
    }
    __syncwarp();
    float* dst = base + (long long)(lane >> 2) * ld + h * 16 + (lane & 3) * 4;
#pragma unroll
    for (int it = 0; it < 4; it++) {
      const int r = it * 8 + (lane >> 2);
      const uint4 u = lds128(stg_get_addr(st, r, lane & 3));
      float4 o = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
      if (RESID) {
        const float4 rr = rbuf[h * 4 + it];
        o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
      }
      if (r < rows_valid) *reinterpret_cast<float4*>(dst) = o;
      dst += 8 * ld;
    }
    __syncwarp();
  }
}

template <int BN, int EPI, int EPW>
__device__ __forceinline__ void epilogue_fast(const GemmKParams& p, uint32_t tmem_base, uint64_t* tfull_bar,
                                              uint64_t* tempty_bar, uint32_t st, uint32_t sb, int warp, int lane) {
  constexpr int CSTR = EPW / 4;                           // warps per TMEM lane quarter = chunk stride of one warp
  constexpr int NCH = (BN / 32 + CSTR - 1) / CSTR;        // 32-column chunks per warp and tile (chunks half, half+CSTR, ...)
  constexpr bool kResid = (EPI == EPI_RESID_F32);
  // RESID keeps its registers for the residual prefetch; the 16-warp variant (112 registers / thread) has none to spare
  constexpr bool kTmemPrefetch = !kResid && EPW == 8;
  const int q = warp & 3;
  const int half = (warp - 2) >> 2;                       // 0 .. CSTR-1
  const bool has_bias = p.bias != nullptr;
  uint32_t it = 0;
  for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, it++) {
    const int n_blk = unit % p.n_groups;
    const int m_blk = (unit / p.n_groups) % p.m_blks;
    const long long row0 = (long long)m_blk * BM + q * 32;
    const int rows_valid = (int)max(0LL, min(32LL, (long long)p.M - row0));
    const uint32_t acc = it & 1;
    const uint32_t acc_phase = (it >> 1) & 1;
    const int colbase = n_blk * BN + half * 32;
    float bnext = 0.f;
    if (has_bias && colbase < p.N) bnext = __ldg(p.bias + colbase + lane);
    float4 rnext[8], rcur[8];
    if (kResid && colbase < p.N) resid_prefetch(rnext, p.resid + row0 * p.ldr + colbase, p.ldr, lane, rows_valid);
    if (kResid) {
      // Pull the residual tile of the NEXT unit of this CTA into L2 now (one 128-byte line per thread and chunk): the
      // register prefetch above keeps only 4 KB per warp in flight, which capped these HBM-bound epilogues at ~3.8 TB/s.
      const int nunit = unit + gridDim.x;
      if (nunit < p.num_units) {
        const int nn = nunit % p.n_groups;
        const long long nrow = (long long)((nunit / p.n_groups) % p.m_blks) * BM + q * 32 + lane;
        if (nrow < p.M) {
          const float* pr = p.resid + nrow * p.ldr + nn * BN + half * 32;
#pragma unroll
          for (int i = 0; i < NCH; i++)
            if ((half + i * CSTR) * 32 < BN && nn * BN + half * 32 + i * (CSTR * 32) < p.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(pr + i * (CSTR * 32)));
        }
      }
    }
    mbar_wait(&tfull_bar[acc], acc_phase);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + half * 32;
    uint32_t raw[2][32];
    if (colbase < p.N) tmem_ld_32x32(taddr, raw[0]);
#pragma unroll
    for (int i = 0; i < NCH; i++) {
      const int col0 = colbase + i * (CSTR * 32);
      if ((half + i * CSTR) * 32 >= BN || col0 >= p.N) break;   // warp-uniform
      constexpr int kCurMask = kTmemPrefetch ? 1 : 0;
      uint32_t(&cur)[32] = raw[i & kCurMask];
      tmem_ld_wait();
      reg_fence32(cur);
      const bool has_next = (i + 1 < NCH) && ((half + (i + 1) * CSTR) * 32 < BN) && (col0 + CSTR * 32 < p.N);
      if (kTmemPrefetch && has_next) tmem_ld_32x32(taddr + (i + 1) * (CSTR * 32), raw[(i + 1) & 1]);
      float v[32];
      if (has_bias) {
        sts32(sb + lane * 4, bnext);
        __syncwarp();
        if (has_next) bnext = __ldg(p.bias + col0 + CSTR * 32 + lane);
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const uint4 b = lds128(sb + k * 16);
          v[4 * k + 0] = __uint_as_float(cur[4 * k + 0]) + __uint_as_float(b.x);
          v[4 * k + 1] = __uint_as_float(cur[4 * k + 1]) + __uint_as_float(b.y);
          v[4 * k + 2] = __uint_as_float(cur[4 * k + 2]) + __uint_as_float(b.z);
          v[4 * k + 3] = __uint_as_float(cur[4 * k + 3]) + __uint_as_float(b.w);
        }
        __syncwarp();
      } else {
#pragma unroll
        for (int k = 0; k < 32; k++) v[k] = __uint_as_float(cur[k]);
      }
      if (kResid) {
#pragma unroll
        for (int k = 0; k < 8; k++) rcur[k] = rnext[k];
        if (has_next) resid_prefetch(rnext, p.resid + row0 * p.ldr + col0 + CSTR * 32, p.ldr, lane, rows_valid);
      }
      if (!kTmemPrefetch && has_next) tmem_ld_32x32(taddr + (i + 1) * (CSTR * 32), raw[0]);   // v[] holds this chunk already
      if (EPI == EPI_BF16) {
        fstore_bf16(st, lane, v, reinterpret_cast<__nv_bfloat16*>(p.C) + row0 * p.ldc + col0, p.ldc, rows_valid);
      } else if (EPI == EPI_F32 || EPI == EPI_RESID_F32) {
        fstore_f32<kResid>(st, lane, v, reinterpret_cast<float*>(p.C) + row0 * p.ldc + col0, p.ldc, rcur, rows_valid);
      } else if (EPI == EPI_GEGLU) {
        if (p.C != nullptr)
          fstore_bf16(st, lane, v, reinterpret_cast<__nv_bfloat16*>(p.C) + row0 * p.ldc + col0, p.ldc, rows_valid);
        float g[16];
#pragma unroll
        for (int k = 0; k < 16; k++) g[k] = gelu_erf_fast(v[2 * k + 1]) * v[2 * k];
        fstore_bf16_half(st, lane, g, reinterpret_cast<__nv_bfloat16*>(p.C2) + row0 * p.ldc2 + (col0 >> 1), p.ldc2, rows_valid);
      } else if (EPI == EPI_L2NORM) {
        if (p.C != nullptr)
          fstore_bf16(st, lane, v, reinterpret_cast<__nv_bfloat16*>(p.C) + row0 * p.ldc + col0, p.ldc, rows_valid);
        if (col0 < p.norm_cols) {
          float ss = 0.f;
#pragma unroll
          for (int k = 0; k < 32; k++) ss = fmaf(v[k], v[k], ss);
          const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
          for (int k = 0; k < 8; k++) {
            const float4 sc = __ldg(reinterpret_cast<const float4*>(p.norm_scale) + k);
            v[4 * k + 0] = v[4 * k + 0] * inv * sc.x;
            v[4 * k + 1] = v[4 * k + 1] * inv * sc.y;
            v[4 * k + 2] = v[4 * k + 2] * inv * sc.z;
            v[4 * k + 3] = v[4 * k + 3] * inv * sc.w;
          }
          fstore_bf16(st, lane, v, reinterpret_cast<__nv_bfloat16*>(p.C2) + row0 * p.ldc2 + col0, p.ldc2, rows_valid);
        }
      } else if (EPI == EPI_BIAS_GELU) {
        if (p.C2 != nullptr)
          fstore_bf16(st, lane, v, reinterpret_cast<__nv_bfloat16*>(p.C2) + row0 * p.ldc2 + col0, p.ldc2, rows_valid);
#pragma unroll
        for (int k = 0; k < 32; k++) v[k] = gelu_erf_fast(v[k]);
        fstore_bf16(st, lane, v, reinterpret_cast<__nv_bfloat16*>(p.C) + row0 * p.ldc + col0, p.ldc, rows_valid);
      }
    }
    tmem_ld_wait();
    // release this accumulator stage back to the MMA warp
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&tempty_bar[acc]);
  }
}

// ---- GEGLU backward fused into the d(g) GEMM -------------------------------------------------------------------------
// acc = dg[m, j] = dL/d(gelu(gate_j) value_j); C = h (bf16, interleaved pre-activation (value_j, gate_j) at columns 2j, 2j+1),
// updated IN PLACE to (dvalue_j, dgate_j) = (dg gelu(gate), dg value gelu'(gate)); column sums of the result (the bias
// gradient of the folded LayerNorm) are accumulated in shared memory per CTA and flushed once at the end.
// Replaces a 311 MB bf16 round trip of dg plus a separate elementwise kernel (attention.py:39-42 backward).
// gelu / gelu' share one exp and one erf (Abramowitz-Stegun 7.1.26, |err| <= 1.5e-7).
__device__ __forceinline__ void geglu_bwd_pair(float dg, float value, float gate, float& dval, float& dgate) {
  const float ax = fabsf(gate);
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f * 0.70710678118654752440f, ax, 1.0f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"((gate * gate) * (-0.5f * 1.4426950408889634f)));
  const float half_erfc = 0.5f * poly * e;                         // 0.5 erfc(|gate| / sqrt 2)
  const float cdf = (gate >= 0.f) ? 1.0f - half_erfc : half_erfc;  // Phi(gate)
  const float pdf = 0.39894228040143267794f * e;                   // phi(gate)
  dval = dg * (gate * cdf);
  dgate = dg * value * fmaf(gate, pdf, cdf);
}

template <int BN>
__device__ __forceinline__ void epilogue_geglu_bwd(const GemmKParams& p, uint32_t tmem_base, uint64_t* tfull_bar,
                                                   uint64_t* tempty_bar, uint32_t st, float* s_cs, int warp, int lane) {
  constexpr int NCH = (BN / 32 + 1) / 2;
  const int q = warp & 3;
  const int half = (warp - 2) >> 2;
  const int tid_e = threadIdx.x - 64;    // 0 .. 255 among the epilogue warps
  for (int i = tid_e; i < 2 * p.N; i += EPI_WARPS * 32) s_cs[i] = 0.f;
  asm volatile("bar.sync 1, %0;" ::"r"(EPI_WARPS * 32) : "memory");
  __nv_bfloat16* hbase = reinterpret_cast<__nv_bfloat16*>(p.C);
  uint32_t it_ = 0;
  for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, it_++) {
    const int n_blk = unit % p.n_groups;
    const int m_blk = (unit / p.n_groups) % p.m_blks;
    const long long row0 = (long long)m_blk * BM + q * 32;
    const int rows_valid = (int)max(0LL, min(32LL, (long long)p.M - row0));
    const uint32_t acc = it_ & 1;
    const uint32_t acc_phase = (it_ >> 1) & 1;
    const int colbase = n_blk * BN + half * 32;
    mbar_wait(&tfull_bar[acc], acc_phase);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + half * 32;
#pragma unroll 1
    for (int i = 0; i < NCH; i++) {
      const int col0 = colbase + i * 64;
      if (col0 >= p.N) break;   // warp-uniform
      uint32_t raw[32];
      tmem_ld_32x32(taddr + i * 64, raw);
      // this lane's h elements in the transposed ownership: rows it*8 + lane/4, dg columns hh*16 + (lane%4)*4 .. +3
      uint4 hv[2][4];
      __nv_bfloat16* hp = hbase + (row0 + (lane >> 2)) * p.ldc + 2 * (col0 + (lane & 3) * 4);
#pragma unroll
      for (int hh = 0; hh < 2; hh++)
#pragma unroll
        for (int it = 0; it < 4; it++) {
          const int r = it * 8 + (lane >> 2);
          hv[hh][it] = (r < rows_valid) ? *reinterpret_cast<const uint4*>(hp + (long long)it * 8 * p.ldc + hh * 32)
                                        : make_uint4(0, 0, 0, 0);
        }
      tmem_ld_wait();
#pragma unroll
      for (int hh = 0; hh < 2; hh++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
          uint4 u;
          u.x = raw[16 * hh + 4 * k + 0];
          u.y = raw[16 * hh + 4 * k + 1];
          u.z = raw[16 * hh + 4 * k + 2];
          u.w = raw[16 * hh + 4 * k + 3];
          sts128(stg_put_addr(st, lane, k), u);
        }
        __syncwarp();
        float cs[8];
#pragma unroll
        for (int k = 0; k < 8; k++) cs[k] = 0.f;
#pragma unroll
        for (int it = 0; it < 4; it++) {
          const int r = it * 8 + (lane >> 2);
          const uint4 d4 = lds128(stg_get_addr(st, r, lane & 3));
          const float dgv[4] = {__uint_as_float(d4.x), __uint_as_float(d4.y), __uint_as_float(d4.z), __uint_as_float(d4.w)};
          uint4 hw = hv[hh][it];
          uint32_t* ph = reinterpret_cast<uint32_t*>(&hw);
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const float2 xv = unpack_bf16x2(ph[k]);   // (value, gate)
            float dval, dgate;
            geglu_bwd_pair(dgv[k], xv.x, xv.y, dval, dgate);
            ph[k] = pack_bf16x2(dval, dgate);
            if (r < rows_valid) {
              cs[2 * k] += dval;
              cs[2 * k + 1] += dgate;
            }
          }
          if (r < rows_valid) *reinterpret_cast<uint4*>(hp + (long long)it * 8 * p.ldc + hh * 32) = hw;
        }
        __syncwarp();
        // column sums: reduce over the 8 lanes that share lane % 4, then one shared-memory add per column
#pragma unroll
        for (int k = 0; k < 8; k++) {
          cs[k] += __shfl_xor_sync(0xffffffffu, cs[k], 4);
          cs[k] += __shfl_xor_sync(0xffffffffu, cs[k], 8);
          cs[k] += __shfl_xor_sync(0xffffffffu, cs[k], 16);
        }
        if ((lane >> 2) == 0) {
          float* dst = s_cs + 2 * (col0 + hh * 16 + (lane & 3) * 4);
#pragma unroll
          for (int k = 0; k < 8; k++) atomicAdd(dst + k, cs[k]);
        }
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&tempty_bar[acc]);
  }
  asm volatile("bar.sync 1, %0;" ::"r"(EPI_WARPS * 32) : "memory");
  if (p.colsum != nullptr)
    for (int i = tid_e; i < 2 * p.N; i += EPI_WARPS * 32) atomicAdd(p.colsum + i, s_cs[i]);
}

template <int BN, int AMAJ, int BMAJ, int EPW>
__global__ void __launch_bounds__((2 + EPW) * 32, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const GemmKParams p) {
  using Cfg = GemmCfg<BN, EPW>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + STAGES * Cfg::B_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  float* arg_merge = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [128][2] (value, index)
  uint8_t* stage_all = reinterpret_cast<uint8_t*>(bars) + 256 + 1024;                     // [EPI_WARPS][2048]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], EPW);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_holder, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
        const int ng = unit % p.n_groups;
        const int rest = unit / p.n_groups;
        const int m_blk = rest % p.m_blks;
        const int split = rest / p.m_blks;
        const int kb0 = split * p.k_blks_per_split;
        const int kb1 = min(p.k_blks, kb0 + p.k_blks_per_split);
        for (int j = 0; j < p.n_per_unit; j++) {
          const int n_blk = ng * p.n_per_unit + j;
          for (int kb = kb0; kb < kb1; kb++) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::A_BYTES + Cfg::B_BYTES);
            uint8_t* a_dst = sA + stage * Cfg::A_BYTES;
            uint8_t* b_dst = sB + stage * Cfg::B_BYTES;
            if (AMAJ == 0) {
              tma_load_2d(a_dst, &tma_a, &full_bar[stage], kb * BK, m_blk * BM);
            } else {
#pragma unroll
              for (int i = 0; i < BM / 64; i++)
                tma_load_2d(a_dst + i * (BK * 128), &tma_a, &full_bar[stage], m_blk * BM + i * 64,
                            kb * BK);
            }
            if (BMAJ == 0) {
              tma_load_2d(b_dst, &tma_b, &full_bar[stage], kb * BK, n_blk * BN);
            } else {
#pragma unroll
              for (int i = 0; i < BN / 64; i++)
                tma_load_2d(b_dst + i * (BK * 128), &tma_b, &full_bar[stage], n_blk * BN + i * 64,
                            kb * BK);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(/*bf16*/ 1, AMAJ, BMAJ, BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;
      for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
        const int rest = unit / p.n_groups;
        const int split = rest / p.m_blks;
        const int kb0 = split * p.k_blks_per_split;
        const int kb1 = min(p.k_blks, kb0 + p.k_blks_per_split);
        for (int j = 0; j < p.n_per_unit; j++, it++) {
          const uint32_t acc = it & 1;
          const uint32_t acc_phase = (it >> 1) & 1;
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc * BN;
          for (int kb = kb0; kb < kb1; kb++) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(sA + stage * Cfg::A_BYTES);
            const uint32_t b_addr = smem_u32(sB + stage * Cfg::B_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; k++) {
              const uint64_t adesc = (AMAJ == 0) ? umma_smem_desc(a_addr + k * 32, 16, 1024)
                                                 : umma_smem_desc(a_addr + k * 2048, BK * 128, 1024);
              const uint64_t bdesc = (BMAJ == 0) ? umma_smem_desc(b_addr + k * 32, 16, 1024)
                                                 : umma_smem_desc(b_addr + k * 2048, BK * 128, 1024);
              umma_bf16(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs retire
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
        }
      }
    }
  } else {
    // ===================== epilogue warps (2..9) =====================
    if (p.fast_epi) {   // kernel-uniform: specialised loops (fast_store shapes, one N tile per unit)
      const uint32_t st32 = smem_u32(stage_all + (warp - 2) * 2048);
      const uint32_t sb32 = st32;   // bias broadcast buffer = first 128 B of the staging buffer (disjoint in time)
      if (EPW == EPW_LIGHT) {              // register-light epilogues only (128 registers / thread)
        switch (p.epi) {
          case EPI_BF16: epilogue_fast<BN, EPI_BF16, EPW>(p, tmem_base, tfull_bar, tempty_bar, st32, sb32, warp, lane); break;
          case EPI_F32: epilogue_fast<BN, EPI_F32, EPW>(p, tmem_base, tfull_bar, tempty_bar, st32, sb32, warp, lane); break;
          case EPI_GEGLU: epilogue_fast<BN, EPI_GEGLU, EPW>(p, tmem_base, tfull_bar, tempty_bar, st32, sb32, warp, lane); break;
          case EPI_L2NORM: epilogue_fast<BN, EPI_L2NORM, EPW>(p, tmem_base, tfull_bar, tempty_bar, st32, sb32, warp, lane); break;
          case EPI_BIAS_GELU: epilogue_fast<BN, EPI_BIAS_GELU, EPW>(p, tmem_base, tfull_bar, tempty_bar, st32, sb32, warp, lane); break;
          default: epilogue_fast<BN, EPI_NONE_DEBUG, EPW>(p, tmem_base, tfull_bar, tempty_bar, st32, sb32, warp, lane); break;
        }
      } else {
        switch (p.epi) {
          case EPI_BF16: epilogue_fast<BN, EPI_BF16, EPW>(p, tmem_base, tfull_bar, tempty_bar, st32, sb32, warp, lane); break;
          case EPI_F32: epilogue_fast<BN, EPI_F32, EPW>(p, tmem_base, tfull_bar, tempty_bar, st32, sb32, warp, lane); break;
          case EPI_RESID_F32: epilogue_fast<BN, EPI_RESID_F32, EPW>(p, tmem_base, tfull_bar, tempty_bar, st32, sb32, warp, lane); break;
          case EPI_GEGLU: epilogue_fast<BN, EPI_GEGLU, EPW>(p, tmem_base, tfull_bar, tempty_bar, st32, sb32, warp, lane); break;
          case EPI_L2NORM: epilogue_fast<BN, EPI_L2NORM, EPW>(p, tmem_base, tfull_bar, tempty_bar, st32, sb32, warp, lane); break;
          case EPI_BIAS_GELU: epilogue_fast<BN, EPI_BIAS_GELU, EPW>(p, tmem_base, tfull_bar, tempty_bar, st32, sb32, warp, lane); break;
          case EPI_GEGLU_BWD:
            if (EPW == EPI_WARPS)
              epilogue_geglu_bwd<BN>(p, tmem_base, tfull_bar, tempty_bar, st32, reinterpret_cast<float*>(stage_all + EPW * 2048), warp, lane);
            break;
          default: epilogue_fast<BN, EPI_NONE_DEBUG, EPW>(p, tmem_base, tfull_bar, tempty_bar, st32, sb32, warp, lane); break;
        }
      }
    } else if (EPW == EPI_WARPS) {   // v1 generic epilogue (odd shapes, ARGMAX, ATOMIC): 8 epilogue warps only
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;   // which alternate 32-column chunks this warp handles
    uint32_t it = 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
      const int ng = unit % p.n_groups;
      const int rest = unit / p.n_groups;
      const int m_blk = rest % p.m_blks;
      const long long row = (long long)m_blk * BM + q * 32 + lane;
      const bool row_ok = row < p.M;
      const long long row0 = (long long)m_blk * BM + q * 32;   // first row of this warp
      const int rows_valid = (int)max(0LL, min(32LL, (long long)p.M - row0));
      uint8_t* st = stage_all + (warp - 2) * 2048;
      float best_v = -INFINITY, sec_v = -INFINITY;
      int best_i = 0, sec_i = 0;
      const bool top2 = p.arg2_out != nullptr;   // kernel-uniform
      for (int j = 0; j < p.n_per_unit; j++, it++) {
        const int n_blk = ng * p.n_per_unit + j;
        const uint32_t acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        // RESID_F32: the residual tile is prefetched one chunk ahead (the first chunk while the MMAs of this tile are
        // still in flight) -- these epilogues are HBM-latency-bound otherwise
        // staged (coalesced) stores need 16-byte aligned rows and whole 32-column chunks
        const bool fast = p.fast_store && (p.N % 32 == 0);
        const bool pf = (p.epi == EPI_RESID_F32) && fast;
        float4 rnext[8];
        if (pf && n_blk * BN + half * 32 < p.N)
          resid_prefetch(rnext, p.resid + row0 * p.ldr + n_blk * BN + half * 32, p.ldr, lane, rows_valid);
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
#pragma unroll 1
        for (int c = half; c < BN / 32; c += EPI_WARPS / 4) {
          const int col0 = n_blk * BN + c * 32;
          if (col0 >= p.N) break;  // warp-uniform
          uint32_t raw[32];
          tmem_ld_32x32(taddr + c * 32, raw);
          float4 rcur[8];
          if (pf) {
#pragma unroll
            for (int i = 0; i < 8; i++) rcur[i] = rnext[i];
            const int cn = c + EPI_WARPS / 4;
            if (cn < BN / 32 && n_blk * BN + cn * 32 < p.N)
              resid_prefetch(rnext, p.resid + row0 * p.ldr + n_blk * BN + cn * 32, p.ldr, lane, rows_valid);
          }
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; i++) v[i] = __uint_as_float(raw[i]);
          const int ncols = min(32, p.N - col0);
          if (p.bias != nullptr) {
            const float* bp = p.bias + col0;
            if (ncols == 32 && ((reinterpret_cast<uintptr_t>(bp) & 15) == 0)) {
#pragma unroll
              for (int i = 0; i < 8; i++) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp) + i);
                v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; i++)
                if (i < ncols) v[i] += __ldg(bp + i);
            }
          }
          if (p.epi == EPI_ARGMAX) {
            if (!top2) {
#pragma unroll
              for (int i = 0; i < 32; i++)
                if (i < ncols && v[i] > best_v) { best_v = v[i]; best_i = col0 + i; }
            } else {
              // running top-2 on PACKED keys: the low 13 mantissa bits of the score carry (8191 - column), so the pair is
              // tracked with three FMNMX per element and no index registers (the (value, index) select chains made this
              // epilogue the bottleneck of the code-book GEMM: 1.30 ms vs 0.65 ms). The scores lose 2^-10 of relative
              // resolution, which only matters for candidates the fp32 re-ranking (ctclip_vq_rerank) orders anyway; equal
              // truncated scores prefer the lower column, like torch.argmax. N <= 8192 (checked by the launcher).
#pragma unroll
              for (int i = 0; i < 32; i++) {
                if (i < ncols) {
                  const float x = __uint_as_float((__float_as_uint(v[i]) & 0xFFFFE000u) | (uint32_t)(8191 - (col0 + i)));
                  const float lo = fminf(x, best_v);
                  best_v = fmaxf(x, best_v);
                  sec_v = fmaxf(sec_v, lo);
                }
              }
            }
            continue;
          }
          if (fast) {   // warp-uniform: every lane takes part in the shared-memory bounce
            if (p.epi == EPI_BF16) {
              staged_store_bf16(st, lane, v, reinterpret_cast<__nv_bfloat16*>(p.C) + row0 * p.ldc + col0, p.ldc, rows_valid);
            } else if (p.epi == EPI_F32 || p.epi == EPI_RESID_F32) {
              staged_store_f32(st, lane, v, reinterpret_cast<float*>(p.C) + row0 * p.ldc + col0, p.ldc,
                               (p.epi == EPI_RESID_F32) ? p.resid + row0 * p.ldr + col0 : nullptr, p.ldr, pf ? rcur : nullptr,
                               rows_valid);
            } else if (p.epi == EPI_GEGLU) {
              if (p.C != nullptr)
                staged_store_bf16(st, lane, v, reinterpret_cast<__nv_bfloat16*>(p.C) + row0 * p.ldc + col0, p.ldc, rows_valid);
              float g[16];
#pragma unroll
              for (int i = 0; i < 16; i++) g[i] = gelu_erf_fast(v[2 * i + 1]) * v[2 * i];
              staged_store_bf16_half(st, lane, g, reinterpret_cast<__nv_bfloat16*>(p.C2) + row0 * p.ldc2 + (col0 >> 1), p.ldc2,
                                     rows_valid);
            } else if (p.epi == EPI_L2NORM) {
              if (p.C != nullptr)
                staged_store_bf16(st, lane, v, reinterpret_cast<__nv_bfloat16*>(p.C) + row0 * p.ldc + col0, p.ldc, rows_valid);
              if (col0 < p.norm_cols) {
                float ss = 0.f;
#pragma unroll
                for (int i = 0; i < 32; i++) ss += v[i] * v[i];
                const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
                for (int i = 0; i < 32; i++) v[i] = v[i] * inv * __ldg(p.norm_scale + i);
                staged_store_bf16(st, lane, v, reinterpret_cast<__nv_bfloat16*>(p.C2) + row0 * p.ldc2 + col0, p.ldc2, rows_valid);
              }
            } else if (p.epi == EPI_BIAS_GELU) {
              if (p.C2 != nullptr)
                staged_store_bf16(st, lane, v, reinterpret_cast<__nv_bfloat16*>(p.C2) + row0 * p.ldc2 + col0, p.ldc2, rows_valid);
#pragma unroll
              for (int i = 0; i < 32; i++) v[i] = gelu_erf_fast(v[i]);
              staged_store_bf16(st, lane, v, reinterpret_cast<__nv_bfloat16*>(p.C) + row0 * p.ldc + col0, p.ldc, rows_valid);
            }
            if (p.epi != EPI_ATOMIC_F32) continue;
          }
          if (!row_ok) continue;
          if (p.epi == EPI_BF16) {
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.C) + row * p.ldc + col0;
            if (ncols == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
              store_bf16x32(dst, v);
            } else {
#pragma unroll
              for (int i = 0; i < 32; i++)
                if (i < ncols) dst[i] = __float2bfloat16(v[i]);
            }
          } else if (p.epi == EPI_F32 || p.epi == EPI_RESID_F32) {
            float* dst = reinterpret_cast<float*>(p.C) + row * p.ldc + col0;
            const float* rs = (p.epi == EPI_RESID_F32) ? (p.resid + row * p.ldr + col0) : nullptr;
            const bool vec = ncols == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) &&
                             (rs == nullptr || (reinterpret_cast<uintptr_t>(rs) & 15) == 0);
            if (vec) {
#pragma unroll
              for (int i = 0; i < 8; i++) {
                float4 o = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                if (rs != nullptr) {
                  const float4 r = *reinterpret_cast<const float4*>(rs + 4 * i);
                  o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
                }
                *reinterpret_cast<float4*>(dst + 4 * i) = o;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; i++)
                if (i < ncols) dst[i] = v[i] + (rs != nullptr ? rs[i] : 0.f);
            }
          } else if (p.epi == EPI_GEGLU) {
            if (p.C != nullptr) {
              __nv_bfloat16* hdst = reinterpret_cast<__nv_bfloat16*>(p.C) + row * p.ldc + col0;
              if (ncols == 32 && ((reinterpret_cast<uintptr_t>(hdst) & 15) == 0)) {
                store_bf16x32(hdst, v);
              } else {
#pragma unroll
                for (int i = 0; i < 32; i++)
                  if (i < ncols) hdst[i] = __float2bfloat16(v[i]);
              }
            }
            __nv_bfloat16* gdst = reinterpret_cast<__nv_bfloat16*>(p.C2) + row * p.ldc2 + (col0 >> 1);
            float g[16];
#pragma unroll
            for (int i = 0; i < 16; i++) g[i] = gelu_erf_fast(v[2 * i + 1]) * v[2 * i];
            if (ncols == 32 && ((reinterpret_cast<uintptr_t>(gdst) & 15) == 0)) {
              uint4* d4 = reinterpret_cast<uint4*>(gdst);
#pragma unroll
              for (int i = 0; i < 2; i++) {
                uint4 u;
                u.x = pack_bf16x2(g[8 * i + 0], g[8 * i + 1]);
                u.y = pack_bf16x2(g[8 * i + 2], g[8 * i + 3]);
                u.z = pack_bf16x2(g[8 * i + 4], g[8 * i + 5]);
                u.w = pack_bf16x2(g[8 * i + 6], g[8 * i + 7]);
                d4[i] = u;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; i++)
                if (2 * i < ncols) gdst[i] = __float2bfloat16(g[i]);
            }
          } else if (p.epi == EPI_L2NORM) {
            // one 32-column chunk == one attention head (dim_head 32): raw -> C, l2norm*scale -> C2
            if (p.C != nullptr) store_bf16x32(reinterpret_cast<__nv_bfloat16*>(p.C) + row * p.ldc + col0, v);
            if (col0 < p.norm_cols) {
              float ss = 0.f;
#pragma unroll
              for (int i = 0; i < 32; i++) ss += v[i] * v[i];
              const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
              for (int i = 0; i < 32; i++) v[i] = v[i] * inv * __ldg(p.norm_scale + i);
              store_bf16x32(reinterpret_cast<__nv_bfloat16*>(p.C2) + row * p.ldc2 + col0, v);
            }
          } else if (p.epi == EPI_BIAS_GELU) {
            // C(bf16) = gelu_erf(acc + bias); C2(bf16, optional) = acc + bias (pre-activation)
            if (p.C2 != nullptr) {
              __nv_bfloat16* pdst = reinterpret_cast<__nv_bfloat16*>(p.C2) + row * p.ldc2 + col0;
              if (ncols == 32) store_bf16x32(pdst, v);
              else {
#pragma unroll
                for (int i = 0; i < 32; i++)
                  if (i < ncols) pdst[i] = __float2bfloat16(v[i]);
              }
            }
#pragma unroll
            for (int i = 0; i < 32; i++) v[i] = gelu_erf_fast(v[i]);
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.C) + row * p.ldc + col0;
            if (ncols == 32) store_bf16x32(dst, v);
            else {
#pragma unroll
              for (int i = 0; i < 32; i++)
                if (i < ncols) dst[i] = __float2bfloat16(v[i]);
            }
          } else if (p.epi == EPI_ATOMIC_F32) {
            float* dst = reinterpret_cast<float*>(p.C) + row * p.ldc + col0;
            if (ncols == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
              for (int i = 0; i < 8; i++) {
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * i),
                             "f"(v[4 * i]), "f"(v[4 * i + 1]), "f"(v[4 * i + 2]), "f"(v[4 * i + 3])
                             : "memory");
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; i++)
                if (i < ncols) atomicAdd(dst + i, v[i]);
            }
          }
        }
        // release this accumulator stage back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      }
      if (p.epi == EPI_ARGMAX) {
        // merge the two column-halves of every row (first maximum wins, like torch.argmax)
        const int rit = q * 32 + lane;
        if (top2) {   // unpack the column indices of the packed keys
          best_i = 8191 - (int)(__float_as_uint(best_v) & 0x1FFFu);
          sec_i = 8191 - (int)(__float_as_uint(sec_v) & 0x1FFFu);
        }
        float* merge2 = reinterpret_cast<float*>(stage_all);   // [128][2] runner-up (value, index): the store staging buffers are idle here
        if (half == 1) {
          arg_merge[2 * rit] = best_v;
          arg_merge[2 * rit + 1] = __int_as_float(best_i);
          if (top2) {
            merge2[2 * rit] = sec_v;
            merge2[2 * rit + 1] = __int_as_float(sec_i);
          }
        }
        asm volatile("bar.sync 1, %0;" ::"r"(EPI_WARPS * 32) : "memory");
        if (half == 0 && row_ok) {
          const float ov = arg_merge[2 * rit];
          const int oi = __float_as_int(arg_merge[2 * rit + 1]);
          if (top2) {
            // merge two sorted pairs (a1 >= a2), (b1 >= b2); earlier index wins ties
            const float o2v = merge2[2 * rit];
            const int o2i = __float_as_int(merge2[2 * rit + 1]);
            auto before = [](float xv, int xi, float yv, int yi) { return xv > yv || (xv == yv && xi < yi); };
            if (before(ov, oi, best_v, best_i)) {   // other half holds the winner: runner-up = better of (its second, our best)
              if (before(o2v, o2i, best_v, best_i)) { sec_v = o2v; sec_i = o2i; }
              else { sec_v = best_v; sec_i = best_i; }
              best_v = ov; best_i = oi;
            } else if (before(ov, oi, sec_v, sec_i)) { sec_v = ov; sec_i = oi; }
            if (sec_v == -INFINITY) sec_i = best_i;   // fewer than two columns seen: no runner-up
            p.arg2_out[row] = sec_i;
          } else if (ov > best_v || (ov == best_v && oi < best_i)) { best_v = ov; best_i = oi; }
          p.arg_out[row] = best_i;
          if (p.argval_out != nullptr) p.argval_out[row] = best_v;
        }
        asm volatile("bar.sync 1, %0;" ::"r"(EPI_WARPS * 32) : "memory");
      }
    }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN, int AMAJ, int BMAJ, int EPW>
static int launch_gemm_epw(const ctclip_gemm_args* a, GemmKParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, EPW>;
  CUtensorMap ta, tb;
  int rc;
  if (AMAJ == 0)
    rc = encode_tmap_2d(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->A, (uint64_t)a->K, (uint64_t)a->M,
                        (uint64_t)a->lda * 2, BK, BM, CU_TENSOR_MAP_SWIZZLE_128B);
  else
    rc = encode_tmap_2d(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->A, (uint64_t)a->M, (uint64_t)a->K,
                        (uint64_t)a->lda * 2, 64, BK, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  if (BMAJ == 0)
    rc = encode_tmap_2d(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->B, (uint64_t)a->K, (uint64_t)a->N,
                        (uint64_t)a->ldb * 2, BK, BN, CU_TENSOR_MAP_SWIZZLE_128B);
  else
    rc = encode_tmap_2d(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->B, (uint64_t)a->N, (uint64_t)a->K,
                        (uint64_t)a->ldb * 2, 64, BK, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;

  p.n_blks = ceil_div(a->N, BN);
  p.n_per_unit = (p.epi == EPI_ARGMAX) ? p.n_blks : 1;
  p.n_groups = p.n_blks / p.n_per_unit;
  p.num_units = p.m_blks * p.n_groups * p.splits;

  auto kern = gemm_tc_kernel<BN, AMAJ, BMAJ, EPW>;
  constexpr int kColsumBytes = 12 * 1024;   // GEGLU_BWD: per-CTA column-sum accumulator (2N floats, N <= 1536)
  constexpr int kMaxSmem = (Cfg::SMEM_BYTES + kColsumBytes <= 227 * 1024) ? Cfg::SMEM_BYTES + kColsumBytes : Cfg::SMEM_BYTES;
  static bool attr_set = false;
  if (!attr_set) {
    CTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    attr_set = true;
  }
  int smem = Cfg::SMEM_BYTES;
  if (p.epi == EPI_GEGLU_BWD) {
    CTB_CHECK_ARG(EPW == EPI_WARPS && kMaxSmem > Cfg::SMEM_BYTES && 2 * p.N * 4 <= kColsumBytes && p.fast_epi,
                  "gemm: GEGLU_BWD needs N <= 1536, 16-byte aligned rows of C and the 8-warp kernel");
    smem = kMaxSmem;
  }
  const int grid = p.num_units < num_sms() ? p.num_units : num_sms();
  kern<<<grid, (2 + EPW) * 32, smem, stream>>>(ta, tb, p);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}

// 12 epilogue warps (3 per TMEM lane quarter, 128 registers each) for the fused epilogues that are latency-bound with
// two warps per scheduler (GEGLU: 390 instructions per 32-column chunk incl. 16 erf-GELUs); the residual / atomic /
// argmax / odd-shape epilogues keep the 8-warp kernel with its larger register budget.
template <int BN, int AMAJ, int BMAJ>
static int launch_gemm(const ctclip_gemm_args* a, GemmKParams& p, cudaStream_t stream) {
  static const int no16 = getenv("CTCLIP_GEMM_EPW8") ? atoi(getenv("CTCLIP_GEMM_EPW8")) : 0;   // debug knob
  const bool light = p.epi == EPI_BF16 || p.epi == EPI_F32 || p.epi == EPI_GEGLU || p.epi == EPI_L2NORM ||
                     p.epi == EPI_BIAS_GELU || p.epi == EPI_NONE_DEBUG;
  if (BN >= 128 && p.fast_epi && light && !no16) return launch_gemm_epw<(BN >= 128 ? BN : 128), AMAJ, BMAJ, EPW_LIGHT>(a, p, stream);
  return launch_gemm_epw<BN, AMAJ, BMAJ, EPI_WARPS>(a, p, stream);
}

}  // namespace ctb

using namespace ctb;

extern "C" int ctclip_gemm_bf16(const ctclip_gemm_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(a != nullptr, "gemm: null args");
  CTB_CHECK_ARG(a->M > 0 && a->N > 0 && a->K > 0, "gemm: bad shape M=%d N=%d K=%d", a->M, a->N, a->K);
  CTB_CHECK_ARG(a->A && a->B, "gemm: null operand");
  CTB_CHECK_ARG((a->lda * 2) % 16 == 0 && (a->ldb * 2) % 16 == 0,
                "gemm: operand pitch must be a multiple of 16 bytes (lda=%lld ldb=%lld)",
                (long long)a->lda, (long long)a->ldb);
  CTB_CHECK_ARG(((uintptr_t)a->A % 16) == 0 && ((uintptr_t)a->B % 16) == 0, "gemm: operands must be 16B aligned");
  CTB_CHECK_ARG(a->epilogue >= 0 && a->epilogue <= 8, "gemm: bad epilogue %d", a->epilogue);
  if (a->epilogue == EPI_GEGLU_BWD)
    CTB_CHECK_ARG(a->C != nullptr && a->N % 32 == 0 && a->ldc % 8 == 0 && a->bias == nullptr,
                  "gemm: GEGLU_BWD needs C = h (bf16 [M, 2N], 16-byte aligned rows), N a multiple of 32 and no bias");
  if (a->epilogue == EPI_L2NORM)
    CTB_CHECK_ARG(a->C2 != nullptr && a->norm_scale != nullptr && a->N % 32 == 0 && a->norm_cols % 32 == 0 &&
                      a->ldc % 8 == 0 && a->ldc2 % 8 == 0,
                  "gemm: L2NORM needs C2, norm_scale, N/norm_cols multiples of 32 and 16B-aligned rows");
  if (a->epilogue == EPI_BIAS_GELU) CTB_CHECK_ARG(a->ldc % 8 == 0 && (a->C2 == nullptr || a->ldc2 % 8 == 0), "gemm: BIAS_GELU needs 16B-aligned rows");
  CTB_CHECK_ARG(a->splits >= 0, "gemm: splits must be >= 0 (0 = choose automatically, ATOMIC_F32 only)");
  CTB_CHECK_ARG(a->splits == 1 || a->epilogue == EPI_ATOMIC_F32, "gemm: split-K needs the ATOMIC_F32 epilogue");
  if (a->epilogue == EPI_ARGMAX) {
    CTB_CHECK_ARG(a->arg_out != nullptr, "gemm: ARGMAX needs arg_out");
    if (a->arg2_out != nullptr)
      CTB_CHECK_ARG(a->N >= 2 && a->N <= 8192, "gemm: ARGMAX with arg2_out packs the column into 13 mantissa bits: N must be in [2, 8192] (got %d)", a->N);
  } else CTB_CHECK_ARG(a->C != nullptr || a->epilogue == EPI_GEGLU || a->epilogue == EPI_L2NORM, "gemm: null C");
  if (a->epilogue == EPI_GEGLU) CTB_CHECK_ARG(a->C2 != nullptr && (a->N % 2) == 0, "gemm: GEGLU needs C2 and even N");
  if (a->epilogue == EPI_RESID_F32) CTB_CHECK_ARG(a->resid != nullptr, "gemm: RESID_F32 needs resid");

  // Tile-N selection: 256 whenever N > 128 (a partial last tile costs less than the doubled shared-memory traffic per
  // flop of 128-wide tiles: gemm_probe, profiles/), 128 for 64 < N <= 128, 64 for tiny N.
  int bn;
  if (a->epilogue == EPI_ARGMAX) bn = (a->N % 256 == 0) ? 256 : (a->N > 64 ? 128 : 64);
  else if (a->N > 128) bn = 256;
  else if (a->N > 64) bn = 128;
  else bn = 64;
  // keep at least ~1 wave of work for mid-sized problems without split-K
  if (bn == 256 && a->epilogue != EPI_ARGMAX && a->epilogue != EPI_ATOMIC_F32 &&
      (long long)ceil_div(a->M, BM) * ceil_div(a->N, 256) < num_sms() && a->N % 128 == 0) bn = 128;
  static const int force_bn = getenv("CTCLIP_GEMM_BN") ? atoi(getenv("CTCLIP_GEMM_BN")) : 0;
  if ((force_bn == 64 || force_bn == 128 || force_bn == 256) && a->epilogue != EPI_ARGMAX) bn = force_bn;

  GemmKParams p;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.m_blks = ceil_div(a->M, BM);
  p.k_blks = ceil_div(a->K, BK);
  int want_splits = a->splits;
  if (want_splits == 0) {
    // automatic split-K: minimise  waves x (k-blocks per unit + epilogue cost of one unit in k-block equivalents)
    const long long tiles = (long long)p.m_blks * ceil_div(a->N, bn);
    const int sms = num_sms();
    long long best_cost = -1;
    want_splits = 1;
    for (int sp = 1; sp <= p.k_blks && sp <= 1024; sp++) {
      const int kbps = ceil_div(p.k_blks, sp);
      const int real = ceil_div(p.k_blks, kbps);
      if (real != sp) continue;
      const long long waves = (tiles * real + sms - 1) / sms;
      const long long cost = waves * (kbps + 6);
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; want_splits = sp; }
    }
  }
  p.splits = want_splits > p.k_blks ? p.k_blks : want_splits;
  p.k_blks_per_split = ceil_div(p.k_blks, p.splits);
  p.splits = ceil_div(p.k_blks, p.k_blks_per_split);  // no empty splits
  p.epi = a->epilogue;
  p.C = a->C; p.ldc = a->ldc;
  p.bias = a->bias;
  p.resid = a->resid; p.ldr = a->ldr;
  p.C2 = a->C2; p.ldc2 = a->ldc2;
  p.arg_out = a->arg_out; p.argval_out = a->argval_out; p.arg2_out = a->arg2_out;
  p.norm_cols = a->norm_cols; p.norm_scale = a->norm_scale;
  p.colsum = a->colsum;
  {
    const bool out_f32 = (a->epilogue == EPI_F32 || a->epilogue == EPI_RESID_F32);
    const int esz = out_f32 ? 4 : 2;
    bool ok = a->epilogue != EPI_ARGMAX && a->epilogue != EPI_ATOMIC_F32;
    if (ok && a->C != nullptr) ok = ((uintptr_t)a->C % 16 == 0) && ((a->ldc * esz) % 16 == 0);
    if (ok && a->C2 != nullptr) ok = ((uintptr_t)a->C2 % 16 == 0) && ((a->ldc2 * 2) % 16 == 0);
    if (ok && a->epilogue == EPI_RESID_F32) ok = ((uintptr_t)a->resid % 16 == 0) && ((a->ldr * 4) % 16 == 0);
    if (ok && a->epilogue == EPI_GEGLU) ok = (a->N % 64 == 0);   // 16-column halves of C2 stay 16-byte aligned
    p.fast_store = ok ? 1 : 0;
    // debug / probing knobs (read once): CTCLIP_GEMM_OLD_EPI=1 keeps the v1 generic epilogue, CTCLIP_GEMM_EPI_NONE=1 drains
    // accumulators without storing (mainloop ceiling; results are garbage), CTCLIP_GEMM_BN=64|128|256 forces the N tile
    static const int old_epi = getenv("CTCLIP_GEMM_OLD_EPI") ? atoi(getenv("CTCLIP_GEMM_OLD_EPI")) : 0;
    static const int epi_none = getenv("CTCLIP_GEMM_EPI_NONE") ? atoi(getenv("CTCLIP_GEMM_EPI_NONE")) : 0;
    p.fast_epi = (ok && (a->N % 32 == 0) && !old_epi) ? 1 : 0;
    if (epi_none && p.fast_epi) p.epi = EPI_NONE_DEBUG;
  }

#define CTB_DISPATCH(BN_)                                                                    \
  do {                                                                                       \
    if (a->a_major == 0 && a->b_major == 0) return launch_gemm<BN_, 0, 0>(a, p, stream);     \
    if (a->a_major == 1 && a->b_major == 1) return launch_gemm<BN_, 1, 1>(a, p, stream);     \
    if (a->a_major == 0 && a->b_major == 1) return launch_gemm<BN_, 0, 1>(a, p, stream);     \
    return launch_gemm<BN_, 1, 0>(a, p, stream);                                             \
  } while (0)
  if (bn == 256) CTB_DISPATCH(256);
  if (bn == 128) CTB_DISPATCH(128);
  CTB_DISPATCH(64);
#undef CTB_DISPATCH
}
