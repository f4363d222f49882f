// EXPERIMENTAL -- tcgen05 / TMEM attention forward (dim_head 32, no key mask, n % 64 == 0).
//
// STATUS: written at the end of round 1 after the GPU budget was spent: it COMPILES for sm_100a but has NOT run on
// hardware yet. It is not called by any default path (the product uses the mma.sync kernels of attention.cu); the only
// caller is tests/test_attention_tc_gpu.py, which is skipped unless CTCLIP_EXPERIMENTAL=1. Round 2 starts by bringing it
// up. Every tcgen05 / mbarrier / descriptor idiom below is the one gemm_tcgen05.cu uses and
// has validated on B200; what is new here is listed under "unverified" at the end of this comment.
//
// Replaces (once validated) attention.py:156-178 for the spatial stack:  out = softmax(q_hat k_hat^T * scale + bias) v.
//
// Structure (correctness-first, fully serialised; the overlap comes later):
//   CTA = one (sequence, head); 160 threads: warps 0-3 = softmax (thread r owns query row r of the current 128-row tile =
//   TMEM lane r), warp 4 = TMEM owner + single-thread MMA issuer.
//   Shared memory (SWIZZLE_128B K-major tiles, exactly the layout TMA writes for the GEMM kernel, produced here with
//   st.shared: 16-byte chunk c of row r lands at chunk c ^ (r & 7)):
//     sQ  [128 rows][128 B]   query tile, d in bytes 0..63 of a row (bytes 64..127 never read: only 2 of 4 UMMA_K steps issued)
//     sK  [n rows][128 B]     keys, same row format                                      (B operand of S = Q K^T)
//     sVt [n/64][32 rows = d][128 B = 64 keys]   V transposed per 64-key block              (B operand of O = P V)
//     sP  [NCH/64][128 rows][128 B = 64 keys]    probabilities of the current key chunk     (A operand of O = P V)
//   TMEM: columns [0, NCH) = S chunk (128 x NCH fp32), [NCH, NCH+32) = O chunk (128 x 32 fp32); 256 columns allocated.
//   Per query tile and key chunk (NCH = 192 keys when 192 | n, else 128 or 64):
//     MMA:      S = Q K_chunk^T                      (tcgen05.mma M=128, N=NCH, 2 x K=16)   -> commit
//     softmax:  tcgen05.ld S, x = S*scale*log2e + bias*log2e, online max / sum, P = exp2(x - m) -> bf16 -> sP
//     MMA:      Oc = P V_chunk                       (M=128, N=32, NCH/16 x K=16, fresh accumulator) -> commit
//     softmax:  tcgen05.ld Oc, o = o*alpha + Oc      (the running output lives in 32 registers per row, FA2-style)
//   then o / l -> bf16, lse = m + log2(l) (log2 domain, same convention as attention.cu).
//
// Unverified on hardware (bring-up checklist): (1) the manual SWIZZLE_128B placement of sQ / sK / sVt / sP against the
// UMMA descriptors; (2) N = 192 and N = 32 instruction descriptors; (3) generic-proxy writes -> fence.proxy.async ->
// tcgen05.mma reads; (4) the S / O TMEM column split and the lane mapping of tcgen05.ld for warps 0-3.
#include <stdlib.h>
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/ctclip_b200.h"

namespace ctb {

constexpr int TCA_THREADS = 160;
constexpr int TCA_DH = 32;
constexpr float kTcaLog2e = 1.4426950408889634f;

struct TcaGeom {
  int n, heads, seq_inner;
  long long seq_outer_stride, tok_stride;
  __device__ __forceinline__ long long row(int seq, int i) const {
    return (long long)(seq / seq_inner) * seq_outer_stride + (seq % seq_inner) + (long long)i * tok_stride;
  }
};

__device__ __forceinline__ float tca_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// byte offset of 16-byte chunk `c` of row `r` inside a [rows][128 B] SWIZZLE_128B tile
__device__ __forceinline__ uint32_t tca_swz(int r, int c) { return (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4); }

__global__ void __launch_bounds__(TCA_THREADS, 1) attn_tc_fwd_kernel(ctclip_attn_args a, int nch) {
  extern __shared__ uint8_t tca_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tca_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const TcaGeom g{a.n, a.heads, a.seq_inner, a.seq_outer_stride, a.tok_stride};
  const int n = a.n;
  uint8_t* sQ = smem;                                   // 16 KB
  uint8_t* sK = sQ + 128 * 128;                         // n * 128 B
  uint8_t* sVt = sK + (size_t)n * 128;                  // (n / 64) * 4 KB
  uint8_t* sP = sVt + (size_t)(n / 64) * 4096;          // (nch / 64) * 16 KB
  uint64_t* bar_mma = reinterpret_cast<uint64_t*>(sP + (size_t)(nch / 64) * 16384);
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bar_mma + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int item = blockIdx.x;
  const int head = item % a.heads, seq = item / a.heads;
  const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(a.q);
  const __nv_bfloat16* k = reinterpret_cast<const __nv_bfloat16*>(a.k);
  const __nv_bfloat16* v = reinterpret_cast<const __nv_bfloat16*>(a.v);
  const __nv_bfloat16* bias = reinterpret_cast<const __nv_bfloat16*>(a.bias);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.o);

  if (tid == 128) {
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 4) {
    tmem_alloc(tmem_holder, 256);
    tmem_relinquish();
  }
  // ---- K and V^T tiles of this (sequence, head): all 160 threads
  for (int idx = tid; idx < n * 4; idx += TCA_THREADS) {
    const int r = idx >> 2, c = idx & 3;
    const uint4 kv = *reinterpret_cast<const uint4*>(k + g.row(seq, r) * a.ldk + head * TCA_DH + c * 8);
    *reinterpret_cast<uint4*>(sK + tca_swz(r, c)) = kv;
    const uint4 vv = *reinterpret_cast<const uint4*>(v + g.row(seq, r) * a.ldv + head * TCA_DH + c * 8);
    const __nv_bfloat16* ve = reinterpret_cast<const __nv_bfloat16*>(&vv);
    uint8_t* blk = sVt + (size_t)(r >> 6) * 4096;          // 64-key block
    const int kc = (r & 63) >> 3, kb = (r & 7) * 2;        // 16-byte chunk and byte inside it of key r within a d-row
#pragma unroll
    for (int e = 0; e < 8; e++) {
      const int d = c * 8 + e;
      *reinterpret_cast<__nv_bfloat16*>(blk + tca_swz(d, kc) + kb) = ve[e];
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const uint32_t sQ_u = smem_u32(sQ), sK_u = smem_u32(sK), sVt_u = smem_u32(sVt), sP_u = smem_u32(sP);
  const uint32_t idesc_qk = umma_idesc(1, 0, 0, 128, nch);
  const uint32_t idesc_pv = umma_idesc(1, 0, 0, 128, 32);
  const float sc2 = a.scale * kTcaLog2e;
  const int n_chunks = n / nch;
  const int kblk = nch / 64;      // 64-key blocks per chunk
  uint32_t ph = 0;                // parity of bar_mma

  for (int q0 = 0; q0 < n; q0 += 128) {
    const int row_i = q0 + tid;    // query row of this softmax thread
    const bool row_ok = tid < 128 && row_i < n;
    // ---- Q tile (softmax threads: one 64-byte row each, zero beyond n)
    if (tid < 128) {
#pragma unroll
      for (int c = 0; c < 4; c++) {
        uint4 qv = make_uint4(0, 0, 0, 0);
        if (row_ok) qv = *reinterpret_cast<const uint4*>(q + g.row(seq, row_i) * a.ldq + head * TCA_DH + c * 8);
        *reinterpret_cast<uint4*>(sQ + tca_swz(tid, c)) = qv;
      }
      fence_proxy_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    float m_run = -INFINITY, l_run = 0.f;
    float oacc[32];
#pragma unroll
    for (int i = 0; i < 32; i++) oacc[i] = 0.f;
    for (int ck = 0; ck < n_chunks; ck++) {
      // ---- S = Q K_chunk^T
      if (tid == 128) {
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 2; ks++) {
          const uint64_t ad = umma_smem_desc(sQ_u + ks * 32, 16, 1024);
          const uint64_t bd = umma_smem_desc(sK_u + (uint32_t)ck * nch * 128 + ks * 32, 16, 1024);
          umma_bf16(tmem_base, ad, bd, idesc_qk, ks > 0 ? 1u : 0u);
        }
        umma_commit(bar_mma);
      }
      float alpha = 1.f;
      if (tid < 128) {
        mbar_wait(bar_mma, ph);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
        const __nv_bfloat16* brow = (bias != nullptr && row_ok) ? bias + ((long long)head * n + row_i) * n + (long long)ck * nch : nullptr;
        // pass 1: row maximum of the chunk
        float mloc = -INFINITY;
        for (int s32 = 0; s32 < nch / 32; s32++) {
          uint32_t raw[32];
          tmem_ld_32x32(taddr + s32 * 32, raw);
          uint4 bq[4];
#pragma unroll
          for (int j = 0; j < 4; j++) bq[j] = brow ? __ldg(reinterpret_cast<const uint4*>(brow + s32 * 32) + j) : make_uint4(0, 0, 0, 0);
          tmem_ld_wait();
          const uint32_t* bw = reinterpret_cast<const uint32_t*>(bq);
#pragma unroll
          for (int j = 0; j < 16; j++) {
            const float2 bb = unpack_bf16x2(bw[j]);
            mloc = fmaxf(mloc, fmaf(__uint_as_float(raw[2 * j]), sc2, bb.x * kTcaLog2e));
            mloc = fmaxf(mloc, fmaf(__uint_as_float(raw[2 * j + 1]), sc2, bb.y * kTcaLog2e));
          }
        }
        const float m_new = fmaxf(m_run, mloc);
        alpha = tca_exp2(m_run - m_new);     // first chunk: exp2(-inf) = 0
        // pass 2: P = exp2(x - m_new) -> bf16 -> sP (row = tid, 64-key blocks)
        float lsum = 0.f;
        for (int s32 = 0; s32 < nch / 32; s32++) {
          uint32_t raw[32];
          tmem_ld_32x32(taddr + s32 * 32, raw);
          uint4 bq[4];
#pragma unroll
          for (int j = 0; j < 4; j++) bq[j] = brow ? __ldg(reinterpret_cast<const uint4*>(brow + s32 * 32) + j) : make_uint4(0, 0, 0, 0);
          tmem_ld_wait();
          const uint32_t* bw = reinterpret_cast<const uint32_t*>(bq);
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; j++) {
            const float2 bb = unpack_bf16x2(bw[j]);
            const float p0 = tca_exp2(fmaf(__uint_as_float(raw[2 * j]), sc2, bb.x * kTcaLog2e) - m_new);
            const float p1 = tca_exp2(fmaf(__uint_as_float(raw[2 * j + 1]), sc2, bb.y * kTcaLog2e) - m_new);
            lsum += p0 + p1;
            pk[j] = pack_bf16x2(p0, p1);
          }
          // 32 keys = 4 chunks of 16 B inside 64-key block (s32 / 2), chunk index (s32 & 1) * 4 + j
          uint8_t* pblk = sP + (size_t)(s32 >> 1) * 16384;
#pragma unroll
          for (int j = 0; j < 4; j++)
            *reinterpret_cast<uint4*>(pblk + tca_swz(tid, (s32 & 1) * 4 + j)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        }
        l_run = l_run * alpha + lsum;
        m_run = m_new;
        fence_proxy_async_smem();
      }
      ph ^= 1;
      tc_fence_before();
      __syncthreads();
      // ---- Oc = P V_chunk (fresh accumulator)
      if (tid == 128) {
        tc_fence_after();
        for (int kb_ = 0; kb_ < kblk; kb_++) {
#pragma unroll
          for (int ks = 0; ks < 4; ks++) {
            const uint64_t ad = umma_smem_desc(sP_u + (uint32_t)kb_ * 16384 + ks * 32, 16, 1024);
            const uint64_t bd = umma_smem_desc(sVt_u + (uint32_t)(ck * kblk + kb_) * 4096 + ks * 32, 16, 1024);
            umma_bf16(tmem_base + nch, ad, bd, idesc_pv, (kb_ > 0 || ks > 0) ? 1u : 0u);
          }
        }
        umma_commit(bar_mma);
      }
      if (tid < 128) {
        mbar_wait(bar_mma, ph);
        tc_fence_after();
        uint32_t raw[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + nch, raw);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i++) oacc[i] = fmaf(oacc[i], alpha, __uint_as_float(raw[i]));
      }
      ph ^= 1;
      tc_fence_before();
      __syncthreads();
    }
    // ---- normalise and store this query tile
    if (row_ok) {
      const float inv = 1.f / l_run;
      const long long grow = g.row(seq, row_i);
      __nv_bfloat16* orow = o + grow * a.ldo + head * TCA_DH;
#pragma unroll
      for (int c = 0; c < 4; c++) {
        uint4 u;
        u.x = pack_bf16x2(oacc[8 * c + 0] * inv, oacc[8 * c + 1] * inv);
        u.y = pack_bf16x2(oacc[8 * c + 2] * inv, oacc[8 * c + 3] * inv);
        u.z = pack_bf16x2(oacc[8 * c + 4] * inv, oacc[8 * c + 5] * inv);
        u.w = pack_bf16x2(oacc[8 * c + 6] * inv, oacc[8 * c + 7] * inv);
        *reinterpret_cast<uint4*>(orow + c * 8) = u;
      }
      if (a.lse != nullptr) a.lse[grow * a.heads + head] = m_run + log2f(l_run);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace ctb

using namespace ctb;

extern "C" int ctclip_attn_fwd_tc(const ctclip_attn_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(a != nullptr && a->q && a->k && a->v && a->o, "attn_fwd_tc: null pointer");
  CTB_CHECK_ARG(a->dim_head == 32 && a->key_mask == nullptr, "attn_fwd_tc: dim_head 32 without key mask only");
  CTB_CHECK_ARG(a->n >= 64 && a->n % 64 == 0 && a->n <= 768, "attn_fwd_tc: n must be a multiple of 64 in [64, 768] (got %d)", a->n);
  CTB_CHECK_ARG(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0 && a->ldo % 8 == 0, "attn_fwd_tc: rows must be 16B aligned");
  const int nch = (a->n % 192 == 0) ? 192 : ((a->n % 128 == 0) ? 128 : 64);
  const size_t smem = 1024 + 128 * 128 + (size_t)a->n * 128 + (size_t)(a->n / 64) * 4096 + (size_t)(nch / 64) * 16384 + 64;
  CTB_CHECK_ARG(smem <= 227 * 1024, "attn_fwd_tc: %zu B of shared memory needed", smem);
  CTB_CUDA(cudaFuncSetAttribute(attn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attn_tc_fwd_kernel<<<a->num_seqs * a->heads, TCA_THREADS, smem, stream>>>(*a, nch);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}
