// Spatial attention core on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), forward and backward.
// Replaces attention.py:152-178 of the reference for the spatial stack of CTViT (ctvit.py:291-297):
//     sim = q_hat k_hat^T * scale + cpb_bias[h, i, j];  attn = softmax(sim);  out = attn v       (+ its autograd backward)
// where q_hat / k_hat are the l2-normalised, per-channel-scaled projections, dim_head = 32, the tokens of one sequence are
// the H x W grid of one frame (n = H*W = 576 at configs[1]) and cpb_bias[h,i,j] = table[rel(i,j), h] is the continuous
// position bias (attention.py:245-282) -- (2H-1)(2W-1) distinct values per head, kept as a TABLE in shared memory
// instead of the 5.3 MB [heads,n,n] tensor the mma.sync kernels (attention.cu) stream from L2.
//
// Why the cosine attention permits a FIXED softmax reference (no online max, no rescaling of O):
//   |q_hat . k_hat| <= max_d |q_scale_d k_scale_d| =: qk_bound  (both are unit vectors times a per-channel scale), so every
//   logit of head h is <= M_h := scale*qk_bound + max_r table[r,h]. exp(logit - M_h) can neither overflow nor -- the
//   logits span at most 2*scale*qk_bound + range(table) ~ 25 nats -- underflow in fp32/bf16, so P = exp2(x - M_h) is
//   formed in ONE pass over the scores, O accumulates in TMEM over all key chunks, and lse = M_h + log2(sum P).
//
// Forward  (attn_tc_fwd_kernel): CTA = one (sequence, head); 2 CTAs / SM (103 KB smem, 256 TMEM columns each).
//   warp 8   TMA producer: K, V of the item (SWIZZLE_64B tiles, 64-byte rows = the head's slice of a token row), Q tiles ring
//   warp 9   MMA issuer (one thread): S = Q_tile K_chunk^T (SS, M=128, N=NCH, K=32) into a 2-stage TMEM ring; after the
//            softmax warps have overwritten S with P (bf16, tcgen05.st): O += P V_chunk (TS: A from TMEM, V MN-major)
//   warps 0-7 softmax: warp w owns TMEM lanes 32*(w&3).. (= query rows) and the column half (w>>2) of the chunk:
//            tcgen05.ld -> x = s*scale*log2e + table'[rel] -> ex2 -> row sum, bf16 pack -> tcgen05.st over its own S columns.
//   The bias look-up is ONE conflict-free LDS per score: table rows are padded to STR = W + 32 floats, so that the 32
//   consecutive queries of a warp (which may wrap to the next grid row) hit 32 different banks, and the key offset of a
//   column is a compile-time immediate because every warp's column range covers whole grid rows (NCH/2 = 2W or W).
// Backward (attn_tc_bwd_kernel): see the comment above that kernel.
#include <stdlib.h>
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/ctclip_b200.h"

namespace ctb {

constexpr float kTcLog2e = 1.4426950408889634f;
constexpr uint32_t SW64 = 4;   // cute::UMMA::LayoutType::SWIZZLE_64B

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

struct TcFwdParams {
  int n, heads, H, num_seqs;
  const float* table;      // [(2H-1)(2W-1), heads] fp32 (may be null: no bias)
  const float* qk_bound;   // device scalar: max_d |q_scale_d * k_scale_d| (null: 1)
  float scale;
  __nv_bfloat16* o;
  long long ldo;
  float* lse;
};

template <int W>
struct TcGeom {
  static constexpr int STR = W + 32;   // padded table row: >= 2W-1 and == W (mod 32)
  static_assert(W == 24 || W == 32, "supported grid widths");
  __host__ __device__ static constexpr int off(int cc) { return (cc / W) * STR + (cc % W); }   // key cc of a row-aligned range
};

// table'[dy][dx] = table[...] * log2e - M_h for one head, rows padded to STR; returns M_h (log2 domain) to every thread.
// Called by ALL threads of the CTA (contains __syncthreads).
template <int W>
__device__ __forceinline__ float tc_load_table(float* sTab, float* sRed, const float* table, const float* qk_bound, float scale, int H,
                                               int heads, int head, int tid, int nthreads) {
  constexpr int STR = TcGeom<W>::STR;
  const int elems = (2 * H - 1) * STR;
  float m = -INFINITY;
  for (int e = tid; e < elems; e += nthreads) {
    const int dyi = e / STR, dxi = e % STR;
    float v = 0.f;
    if (dxi < 2 * W - 1) {
      v = (table != nullptr) ? __ldg(table + ((long long)dyi * (2 * W - 1) + dxi) * heads + head) * kTcLog2e : 0.f;
      m = fmaxf(m, v);
    }
    sTab[e] = v;
  }
  m = warp_max(m);
  if ((tid & 31) == 0) sRed[tid >> 5] = m;
  __syncthreads();
  float mh = -INFINITY;
  for (int w = 0; w < nthreads / 32; w++) mh = fmaxf(mh, sRed[w]);
  const float qkb = (qk_bound != nullptr) ? __ldg(qk_bound) : 1.0f;
  mh += scale * kTcLog2e * qkb;
  return mh;
}

// ------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------
template <int NCH, int W>
__global__ void __launch_bounds__(320, (NCH == 96) ? 2 : 1)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                   const __grid_constant__ CUtensorMap tv, const TcFwdParams p) {
  constexpr int HALF = NCH / 2;
  constexpr int STR = TcGeom<W>::STR;
  constexpr int OCOL = 2 * NCH;
  static_assert(HALF % W == 0 && HALF % 16 == 0, "a softmax warp's column range must cover whole grid rows");
  extern __shared__ uint8_t tc_smem_raw[];
  // (offset arithmetic on the __shared__ array keeps the address space known to the compiler: LDS / STS, not generic LD / ST)
  uint8_t* smem = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);
  const int n = p.n;
  uint8_t* sK = smem;
  uint8_t* sV = sK + (size_t)n * 64;
  uint8_t* sQ = sV + (size_t)n * 64;                       // 2 x [128][64 B]
  float* sTab = reinterpret_cast<float*>(sQ + 2 * 8192);
  const int tab_elems = ((2 * p.H - 1) * STR + 31) & ~31;
  float* sL = sTab + tab_elems;                            // [2 tile parities][2 halves][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sL + 512);
  uint64_t* k_full = bars + 0;
  uint64_t* v_full = bars + 1;
  uint64_t* q_full = bars + 2;    // [2]
  uint64_t* q_empty = bars + 4;   // [2]
  uint64_t* s_full = bars + 6;    // [2]
  uint64_t* p_full = bars + 8;    // [2]
  uint64_t* o_full = bars + 10;
  uint32_t* holder = reinterpret_cast<uint32_t*>(bars + 12);
  float* sRed = reinterpret_cast<float*>(holder + 2);      // [10]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int item = blockIdx.x;
  const int head = item % p.heads, seq = item / p.heads;
  const int NT = (n + 127) / 128, NC = n / NCH, NB = NT * NC;
  const long long row0 = (long long)seq * n;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tq);
    tma_prefetch_desc(&tk);
    tma_prefetch_desc(&tv);
    mbar_init(k_full, 1);
    mbar_init(v_full, 1);
    for (int i = 0; i < 2; i++) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 8);
    }
    mbar_init(o_full, 1);
    fence_barrier_init();
    // the loads of this item do not depend on anything else: start them before the table set-up
    mbar_arrive_expect_tx(k_full, (uint32_t)n * 64);
    for (int r = 0; r < n; r += 64) tma_load_2d(sK + (size_t)r * 64, &tk, k_full, head * 32, (int)(row0 + r));
    mbar_arrive_expect_tx(&q_full[0], 8192);
    tma_load_2d(sQ, &tq, &q_full[0], head * 32, (int)row0);
    mbar_arrive_expect_tx(v_full, (uint32_t)n * 64);
    for (int r = 0; r < n; r += 64) tma_load_2d(sV + (size_t)r * 64, &tv, v_full, head * 32, (int)(row0 + r));
  }
  if (warp == 9) {
    tmem_alloc(holder, 256);
    tmem_relinquish();
  }
  const float sc2 = p.scale * kTcLog2e;
  tc_fence_before();
  const float Mh = tc_load_table<W>(sTab, sRed, p.table, p.qk_bound, p.scale, p.H, p.heads, head, tid, 320);   // contains __syncthreads
  tc_fence_after();
  const uint32_t tmem_base = *holder;

  if (warp < 8) {
    // ===================== softmax warps =====================
    const int q = warp & 3, hf = warp >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int e = tid; e < (2 * p.H - 1) * STR; e += 256) sTab[e] -= Mh;   // fold the softmax reference into the table
    named_bar_sync(1, 256);
    for (int t = 0; t < NT; t++) {
      const int i = t * 128 + r;
      const bool warp_valid = (t * 128 + q * 32) < n;      // n % 32 == 0: a warp is valid or idle as a whole
      const int ic = i < n ? i : n - 1;
      const int qi = (ic / W + p.H - 1) * STR + (ic % W) + (W - 1);
      float l0 = 0.f, l1 = 0.f;
      for (int c = 0; c < NC; c++) {
        const int b = t * NC + c, st = b & 1;
        mbar_wait_tag(&s_full[st], (b >> 1) & 1, 10 + st);
        if (warp_valid) {
          tc_fence_after();
          const uint32_t scol = lane_base + st * NCH + hf * HALF;
          // groups of 16 keys: the TMEM load of group g+1 is in flight while group g goes through the exponentials
          uint32_t sv[HALF / 16][16];
          tmem_ld_32x16(scol, sv[0]);
          const int j0 = c * NCH + hf * HALF;              // multiple of W
          const float* tp = sTab + (qi - (j0 / W) * STR);
#pragma unroll
          for (int g = 0; g < HALF / 16; g++) {
            tmem_ld_wait_dep1(sv[g]);
            if (g + 1 < HALF / 16) tmem_ld_32x16(scol + (g + 1) * 16, sv[g + 1]);
            uint32_t pk[8];
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
              const int cc = g * 16 + e;
              const float x0 = fmaf(__uint_as_float(sv[g][e]), sc2, tp[-TcGeom<W>::off(cc)]);
              const float x1 = fmaf(__uint_as_float(sv[g][e + 1]), sc2, tp[-TcGeom<W>::off(cc + 1)]);
              const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
              l0 += p0;
              l1 += p1;
              pk[e / 2] = pack_bf16x2(p0, p1);
            }
            tmem_st_32x8(scol + g * 8, pk);
          }
          tmem_st_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[st]);
      }
      // ---- tile epilogue: o = O / l, lse = M_h + log2 l
      const int lp = t & 1;
      if (warp_valid) sL[(lp * 2 + hf) * 128 + r] = l0 + l1;
      named_bar_sync(1, 256);
      mbar_wait_tag(o_full, t & 1, 12);
      if (warp_valid) {
        tc_fence_after();
        const float lt = sL[(lp * 2) * 128 + r] + sL[(lp * 2 + 1) * 128 + r];
        const float inv = 1.f / lt;
        uint32_t ov[16];
        tmem_ld_32x16(lane_base + OCOL + hf * 16, ov);
        tmem_ld_wait();
        if (i < n) {
          const long long grow = row0 + i;
          uint4 u0, u1;
          u0.x = pack_bf16x2(__uint_as_float(ov[0]) * inv, __uint_as_float(ov[1]) * inv);
          u0.y = pack_bf16x2(__uint_as_float(ov[2]) * inv, __uint_as_float(ov[3]) * inv);
          u0.z = pack_bf16x2(__uint_as_float(ov[4]) * inv, __uint_as_float(ov[5]) * inv);
          u0.w = pack_bf16x2(__uint_as_float(ov[6]) * inv, __uint_as_float(ov[7]) * inv);
          u1.x = pack_bf16x2(__uint_as_float(ov[8]) * inv, __uint_as_float(ov[9]) * inv);
          u1.y = pack_bf16x2(__uint_as_float(ov[10]) * inv, __uint_as_float(ov[11]) * inv);
          u1.z = pack_bf16x2(__uint_as_float(ov[12]) * inv, __uint_as_float(ov[13]) * inv);
          u1.w = pack_bf16x2(__uint_as_float(ov[14]) * inv, __uint_as_float(ov[15]) * inv);
          uint4* dst = reinterpret_cast<uint4*>(p.o + grow * p.ldo + head * 32 + hf * 16);
          dst[0] = u0;
          dst[1] = u1;
          if (hf == 0 && p.lse != nullptr) p.lse[grow * p.heads + head] = Mh + log2f(lt);
        }
      }
    }
  } else if (warp == 8) {
    // ===================== TMA producer: remaining Q tiles =====================
    if (lane == 0) {
      for (int t = 1; t < NT; t++) {
        const int slot = t & 1;
        mbar_wait_tag(&q_empty[slot], ((t >> 1) & 1) ^ 1, 20 + slot);
        mbar_arrive_expect_tx(&q_full[slot], 8192);
        tma_load_2d(sQ + slot * 8192, &tq, &q_full[slot], head * 32, (int)(row0 + t * 128));
      }
    }
  } else {
    // ===================== MMA issuer (all 32 lanes run the loop; tcgen05 instructions on the elected lane) =====================
    {
      const bool leader = elect_one_sync();
      constexpr uint32_t idesc_s = umma_idesc(1, 0, 0, 128, NCH);
      constexpr uint32_t idesc_pv = umma_idesc(1, 0, 1, 128, 32);     // A = P from TMEM (K-major), B = V MN-major
      constexpr uint32_t hi64 = umma_desc_hi(512, SW64);              // every operand tile here: 64-byte rows, 8-row groups 512 B apart
      const uint32_t q_lo = umma_desc_lo(smem_u32(sQ), 16), k_lo = umma_desc_lo(smem_u32(sK), 16), v_lo = umma_desc_lo(smem_u32(sV), 512);
      // (t, c) of the NEXT S block to issue / of the block whose P is consumed: advanced incrementally, no divisions in this thread
      int ts = 0, cs = 0;
      auto issue_s = [&]() {
        const int st = (ts * NC + cs) & 1;
        if (cs == 0) mbar_wait_tag(&q_full[ts & 1], (ts >> 1) & 1, 30 + (ts & 1));
        tc_fence_after();
        const uint32_t ql = q_lo + (ts & 1) * (8192 >> 4), kl = k_lo + (uint32_t)cs * (NCH * 64 >> 4);
        if (leader) {
          umma_bf16(tmem_base + st * NCH, umma_desc_join(ql, hi64), umma_desc_join(kl, hi64), idesc_s, 0u);
          umma_bf16(tmem_base + st * NCH, umma_desc_join(ql + 2, hi64), umma_desc_join(kl + 2, hi64), idesc_s, 1u);
          umma_commit(&s_full[st]);
          if (cs == NC - 1) umma_commit(&q_empty[ts & 1]);   // every S MMA of this Q tile has been issued
        }
        __syncwarp();
        if (++cs == NC) { cs = 0; ts++; }
      };
      mbar_wait_tag(k_full, 0, 32);
      issue_s();
      int c = 0;
      for (int b = 0; b < NB; b++) {
        if (b + 1 < NB) issue_s();
        const int st = b & 1;
        if (b == 0) mbar_wait_tag(v_full, 0, 33);
        mbar_wait_tag(&p_full[st], (b >> 1) & 1, 34 + st);
        tc_fence_after();
        const uint32_t vl = v_lo + (uint32_t)c * (NCH * 64 >> 4);
        if (leader) {
#pragma unroll
          for (int ks = 0; ks < NCH / 16; ks++) {
            const int k0 = ks * 16;
            const uint32_t acol = st * NCH + (k0 / HALF) * HALF + (k0 % HALF) / 2;
            umma_bf16_ts(tmem_base + OCOL, tmem_base + acol, umma_desc_join(vl + ks * (1024 >> 4), hi64), idesc_pv, (c > 0 || ks > 0) ? 1u : 0u);
          }
          if (c == NC - 1) umma_commit(o_full);
        }
        __syncwarp();
        if (++c == NC) c = 0;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}


// ------------------------------------------------------------------------------------------------------------------
// backward: ONE pass, 5 contractions + 1 exponential per score (the mma.sync path of attention.cu runs two recomputing
// passes: 7 contractions + 2 exponentials). Persistent, one CTA per SM, TMEM 480 / 512 columns.
//
// Orientation: KEYS on the TMEM lanes. Per (sequence, head) item, per block of 128 keys (kb) and half-block of 64 queries (qc):
//   MMA      S^T  = K_blk  Q_c^T      (SS, M=128 keys, N=64 queries, K=32)        -> stage columns [0, 64)
//   MMA      dP^T = V_blk dO_c^T      (SS)                                         -> stage columns [64, 128)
//   warps0-7 P^T = exp2(S^T*scale*log2e + table'[rel] - lse_i),  dS^T = P^T o (dP^T - delta_i)   (thread = key row, warp = 32 keys x 32 queries)
//            P^T  -> bf16 -> TMEM over its own S^T columns (tcgen05.st)
//            dS^T -> bf16 -> (a) shared tile [key][query] (SWIZZLE_128B), (b) the global spill for the table gradient
//   MMA      dV_blk += P^T dO_c       (TS: A from TMEM; B = the dO chunk read MN-major)
//   MMA      dK_blk += dS^T Q_c       (SS: A = the shared dS tile read K-major; B = the Q chunk read MN-major)
//   MMA      dQ_pair += dS K_blk      (SS: A = the SAME shared dS tile read MN-major, 128 queries = two half-blocks; B = K_blk MN-major)
// dV/dK accumulate over the query chunks of one key block, dQ (5 tiles of 128 queries = 160 columns) over the key blocks of
// the item; both are drained by the softmax warps one block later, while the tensor pipe already works on the next block.
// Operands stream through TMA rings that run across item boundaries (K/V block ring of 2, Q/dO chunk ring of 4 -- each
// chunk is re-read once per key block from L2), so there is no per-item prologue bubble.
// The gradient of the position-bias table needs sum over sequences of dS: the bf16 dS^T tiles are spilled (read back from the
// shared tile so that every global store covers full 32-byte sectors) and attn_dtab_reduce_kernel sums them over the
// sequences straight into the (2H-1)(2W-1) table bins -- the [heads, n, n] dbias tensor of the mma.sync path never exists.
// The kernel is bound by shared-memory wavefronts (ncu: LSU 39 % + tensor-core operand fetch 26 %), not by the tensor pipe
// (17 %) or the SFU (22 %): profiles/r2_ncu_attn_bwd.txt.
// ------------------------------------------------------------------------------------------------------------------
struct TcBwdParams {
  int n, heads, H, num_seqs;
  const float* table;
  float scale;
  const float* lse;
  const float* delta;
  __nv_bfloat16* dq; long long ld_dq;
  __nv_bfloat16* dk; long long ld_dk;
  __nv_bfloat16* dv; long long ld_dv;
  __nv_bfloat16* ds_spill;   // [num_seqs*heads][n keys][n queries] or null
  float* dbias_t;            // alternative to the spill: fp32 [heads][n keys][n queries], accumulated with red.global.add.v4.f32 (L2-resident)
};

constexpr int TCB_QD_SLOTS = 4;
// Q chunk, dO chunk, then EITHER the per-query records {lse, delta, 4*ci} (1 KB) OR the two "augmented k-step" tiles (2 KB each)
constexpr int TCB_QD_BYTES = 4096 + 4096 + 4096;
constexpr int TCB_COL_DV = 256, TCB_COL_DK = 288, TCB_COL_DQ = 320;
constexpr uint32_t SW128 = 2;

// EW = number of softmax-backward warps (8 or 16): warp w owns TMEM lanes 32*(w & 3).. and the (w >> 2)-th slice of CW = 256 / EW
// query columns of every half-block. 16 warps (4 per scheduler) hide the LDS -> LDS -> MUFU latency chain of the element loop
// that 8 warps (2 per scheduler, 38 % issue utilisation in profiles/r2c) leave exposed.
// AUG: the per-QUERY terms of the element loop are folded into the tensor-core products by one extra k-step each:
//   S'^T  = K_blk Q_c^T  + ONES [-lse_i / (scale log2e)]^T      dP'^T = V_blk dO_c^T + ONES [-delta_i]^T
// (ONES: [128 keys][16] with three leading 1.0; the query vectors are split into three bf16 terms so that the fp32 accumulator
// receives them to ~2^-24; both tiles in the no-swizzle K-major layout, 8-row x 16-byte core matrices). The element loop then is
//   p = exp2(S' * scale log2e + table'),   dS = p * dP'
// without the per-column {lse, delta, table offset} record: its broadcast LDS.128 cost FOUR shared-memory wavefronts per element
// and made the kernel shared-memory bound (profiles/r2d: LSU wavefronts 53 % + tensor-core operand reads 18 % of the pipe); the
// table offset of a column is now arithmetic (16 consecutive queries wrap at most once around the grid width).
template <int W, int EW, bool AUG>
__global__ void __launch_bounds__((EW + 2) * 32, 1)
attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tdo,
                   const __grid_constant__ CUtensorMap tk, const __grid_constant__ CUtensorMap tv, const TcBwdParams p) {
  constexpr int STR = TcGeom<W>::STR;
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* smem = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);
  const int n = p.n;
  uint8_t* sKV = smem;                                   // 2 slots x (K block 8 KB + V block 8 KB)
  uint8_t* sDS = sKV + 2 * 16384;                        // 2 buffers x [2 query groups][128 keys][128 B]
  uint8_t* sQD = sDS + 2 * 32768;                        // 4 slots x TCB_QD_BYTES
  uint8_t* sOnes = sQD + TCB_QD_SLOTS * TCB_QD_BYTES;    // [128][16] bf16, no-swizzle K-major (AUG)
  float* sTab = reinterpret_cast<float*>(sOnes + 4096);
  const int tab_elems = ((2 * p.H - 1) * STR + 31) & ~31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sTab + tab_elems);
  uint64_t* kv_full = bars + 0;     // [2]
  uint64_t* kv_empty = bars + 2;    // [2]
  uint64_t* qd_full = bars + 4;     // [4]
  uint64_t* qd_empty = bars + 8;    // [4]
  uint64_t* s_full = bars + 12;     // [2]
  uint64_t* p_full = bars + 14;     // [2]
  uint64_t* ds_free = bars + 16;    // [2]
  uint64_t* acc_full = bars + 18;
  uint64_t* dq_full = bars + 19;
  uint32_t* holder = reinterpret_cast<uint32_t*>(bars + 20);
  float* sRed = reinterpret_cast<float*>(holder + 2);    // [EW + 2]
  constexpr int NTHREADS = (EW + 2) * 32;
  constexpr int CW = 256 / EW;                           // query columns per warp and half-block (32 or 16)
  constexpr int NG = CW / 16;                            // groups of 16 columns
  constexpr int CPR = CW / 8;                            // 16-byte chunks of a dS row per warp
  constexpr int DC = 128 / EW;                           // accumulator columns a warp drains (16 or 8)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NKB = (n + 127) / 128, NQC = n / 64;
  const int BPI = NKB * NQC;                             // half-blocks per item
  const int items = p.num_seqs * p.heads;
  const int my_items = (items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const long long total_blocks = (long long)my_items * BPI;

  if (warp == EW && lane == 0) {
    tma_prefetch_desc(&tq);
    tma_prefetch_desc(&tdo);
    tma_prefetch_desc(&tk);
    tma_prefetch_desc(&tv);
    for (int i = 0; i < 2; i++) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], EW);
      mbar_init(&ds_free[i], 1);
    }
    for (int i = 0; i < TCB_QD_SLOTS; i++) {
      mbar_init(&qd_full[i], 33);      // 1 expect_tx arrival (TMA bytes) + 32 lanes that wrote the per-query records
      mbar_init(&qd_empty[i], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(dq_full, 1);
    fence_barrier_init();
  }
  if (warp == EW + 1) {
    tmem_alloc(holder, 512);
    tmem_relinquish();
  }
  if (AUG) {
    // zero the augmented tiles of every slot (their unused k columns must stay zero) and build ONES
    for (int i = tid; i < TCB_QD_SLOTS * 256; i += NTHREADS)
      *reinterpret_cast<uint4*>(sQD + (i >> 8) * TCB_QD_BYTES + 8192 + (i & 255) * 16) = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < 256; i += NTHREADS) {   // 16 groups x (k 0-7 core matrix, k 8-15 core matrix) x 8 rows
      const bool first_half = ((i >> 3) & 1) == 0;
      *reinterpret_cast<uint4*>(sOnes + i * 16) = first_half ? make_uint4(0x3F803F80u, 0x00003F80u, 0, 0) : make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
  }
  // NOTE: the table of head h is only valid for items of head h: a CTA walks items blockIdx.x + k*gridDim.x, so with
  // gridDim.x % heads == 0 (enforced by the launcher) every item of this CTA has the same head.
  const int head = (int)blockIdx.x % p.heads;
  tc_fence_before();
  (void)tc_load_table<W>(sTab, sRed, p.table, nullptr, 0.f, p.H, p.heads, head, tid, NTHREADS);   // table * log2e (no reference shift: lse is subtracted)
  tc_fence_after();
  const uint32_t tmem_base = *holder;
  const float sc2 = p.scale * kTcLog2e;

  if (warp < EW) {
    // ===================== softmax-backward warps =====================
    const int q = warp & 3, cq = warp >> 2;
    const int r = q * 32 + lane;                          // key row inside the block = TMEM lane
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    int u = 0, v = 0, m = 0;                              // key blocks, query pairs, items completed so far (ring parities)
    int prev_item = 0, prev_kb = 0;
    long long g = 0;
    auto ld_acc = [&](uint32_t col, uint32_t (&a)[DC]) {
      if constexpr (DC == 16) tmem_ld_32x16(col, a);
      else tmem_ld_32x8(col, a);
    };
    auto st_acc = [&](__nv_bfloat16* dst, const uint32_t (&a)[DC], float mul) {   // DC fp32 -> bf16, DC * 2 bytes
      uint32_t x[DC / 2];
#pragma unroll
      for (int e = 0; e < DC / 2; e++) x[e] = pack_bf16x2(__uint_as_float(a[2 * e]) * mul, __uint_as_float(a[2 * e + 1]) * mul);
      uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
      for (int e = 0; e < DC / 8; e++) d4[e] = make_uint4(x[4 * e], x[4 * e + 1], x[4 * e + 2], x[4 * e + 3]);
    };
    auto drain_dvdk = [&](int item_, int kb_, uint32_t parity) {   // dV / dK of key block (item_, kb_): TMEM -> bf16 -> global
      mbar_wait_tag(acc_full, parity, 40);
      tc_fence_after();
      const int key = kb_ * 128 + r;
      uint32_t a[DC], b[DC];
      ld_acc(lane_base + TCB_COL_DV + cq * DC, a);
      ld_acc(lane_base + TCB_COL_DK + cq * DC, b);
      tmem_ld_wait();
      if (key < n) {
        const int hd = item_ % p.heads;
        const long long row = (long long)(item_ / p.heads) * n + key;
        st_acc(p.dv + row * p.ld_dv + hd * 32 + cq * DC, a, 1.0f);
        st_acc(p.dk + row * p.ld_dk + hd * 32 + cq * DC, b, p.scale);
      }
    };
    auto drain_dq = [&](int item_, uint32_t parity) {     // dQ of a finished item: NKB tiles of 128 queries
      mbar_wait_tag(dq_full, parity, 41);
      tc_fence_after();
      const int hd = item_ % p.heads;
      for (int t = 0; t < NKB; t++) {
        const int qi = t * 128 + r;
        uint32_t a[DC];
        ld_acc(lane_base + TCB_COL_DQ + t * 32 + cq * DC, a);
        tmem_ld_wait();
        if (qi < n) {
          const long long row = (long long)(item_ / p.heads) * n + qi;
          st_acc(p.dq + row * p.ld_dq + hd * 32 + cq * DC, a, p.scale);
        }
      }
    };
    for (int il = 0; il < my_items; il++) {
      const int item = (int)blockIdx.x + il * (int)gridDim.x;
      for (int kb = 0; kb < NKB; kb++) {
        const int key = kb * 128 + r;
        const bool warp_valid = (kb * 128 + q * 32) < n;
        const int kc = key < n ? key : n - 1;
        const int tj = (p.H - 1 - kc / W) * STR + (W - 1 - kc % W);     // table index = tj + ci(query)
        const char* tpb = reinterpret_cast<const char*>(sTab + tj);      // + ci * 4 (the records hold the byte offset)
        for (int qc = 0; qc < NQC; qc++, g++) {
          const int st = (int)(g & 1), slot = (int)(g % TCB_QD_SLOTS);
          const int buf = v & 1, grp = qc & 1;
          uint8_t* tile_row = sDS + buf * 32768 + grp * 16384 + r * 128;
          mbar_wait_tag(&qd_full[slot], (uint32_t)((g / TCB_QD_SLOTS) & 1), 42);
          mbar_wait_tag(&s_full[st], (uint32_t)((g >> 1) & 1), 43);
          if (grp == 0) mbar_wait_tag(&ds_free[buf], (uint32_t)(((v >> 1) & 1) ^ 1), 44);
          if (warp_valid) {
            tc_fence_after();
            const uint32_t scol = lane_base + st * 128 + cq * CW;       // this warp's S^T columns; dP^T sits 64 columns further
            const float4* rec = reinterpret_cast<const float4*>(sQD + slot * TCB_QD_BYTES + 8192) + cq * CW;
            (void)rec;
            // groups of 16 queries: the TMEM loads of group g+1 are in flight while group g is computed
            uint32_t sv[NG][16], dp[NG][16];
            tmem_ld_32x16(scol, sv[0]);
            tmem_ld_32x16(scol + 64, dp[0]);
#pragma unroll
            for (int gi = 0; gi < NG; gi++) {
              tmem_ld_wait_dep(sv[gi], dp[gi]);
              if (gi + 1 < NG) {
                tmem_ld_32x16(scol + (gi + 1) * 16, sv[gi + 1]);
                tmem_ld_32x16(scol + 64 + (gi + 1) * 16, dp[gi + 1]);
              }
              uint32_t pk[8], dk_[8];
              float dsf[16];
              // AUG: table offsets of the 16 queries i0 .. i0+15 = c0 + e (+ STR - W once the grid row wraps at e = w0)
              const int i0 = qc * 64 + cq * CW + gi * 16;
              const int qrow = i0 / W, qcol = i0 - qrow * W;
              const float* tA = reinterpret_cast<const float*>(tpb) + qrow * STR + qcol;
              const float* tB = tA + (STR - W);
              const int w0 = W - qcol;
#pragma unroll
              for (int e = 0; e < 16; e += 2) {
                float x0, x1, d0, d1, p0, p1;
                if constexpr (AUG) {
                  const float t0 = (e < w0 ? tA : tB)[e], t1 = (e + 1 < w0 ? tA : tB)[e + 1];
                  x0 = fmaf(__uint_as_float(sv[gi][e]), sc2, t0);
                  x1 = fmaf(__uint_as_float(sv[gi][e + 1]), sc2, t1);
                  p0 = ex2_approx(x0);
                  p1 = ex2_approx(x1);
                  d0 = p0 * __uint_as_float(dp[gi][e]);
                  d1 = p1 * __uint_as_float(dp[gi][e + 1]);
                } else {
                  const float4 r0 = rec[gi * 16 + e], r1 = rec[gi * 16 + e + 1];   // {lse_i, delta_i, 4*ci}: broadcast LDS.128
                  const float t0 = *reinterpret_cast<const float*>(tpb + __float_as_int(r0.z));
                  const float t1 = *reinterpret_cast<const float*>(tpb + __float_as_int(r1.z));
                  x0 = fmaf(__uint_as_float(sv[gi][e]), sc2, t0) - r0.x;
                  x1 = fmaf(__uint_as_float(sv[gi][e + 1]), sc2, t1) - r1.x;
                  p0 = ex2_approx(x0);
                  p1 = ex2_approx(x1);
                  d0 = p0 * (__uint_as_float(dp[gi][e]) - r0.y);
                  d1 = p1 * (__uint_as_float(dp[gi][e + 1]) - r1.y);
                }
                pk[e / 2] = pack_bf16x2(p0, p1);
                dk_[e / 2] = pack_bf16x2(d0, d1);
                dsf[e] = d0;
                dsf[e + 1] = d1;
              }
              if (p.dbias_t != nullptr) {   // table gradient: fp32 reductions straight into the L2-resident [h][key][query] table
                float* dst = p.dbias_t + ((long long)(item % p.heads) * n + key) * n + qc * 64 + cq * CW + gi * 16;
#pragma unroll
                for (int k4 = 0; k4 < 4; k4++)
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * k4), "f"(dsf[4 * k4]), "f"(dsf[4 * k4 + 1]),
                               "f"(dsf[4 * k4 + 2]), "f"(dsf[4 * k4 + 3])
                               : "memory");
              }
              tmem_st_32x8(scol + gi * 8, pk);                          // P^T over this warp's own (already consumed) S^T columns
#pragma unroll
              for (int c2 = 0; c2 < 2; c2++)
                *reinterpret_cast<uint4*>(tile_row + (((cq * CPR + gi * 2 + c2) ^ (r & 7)) << 4)) =
                    make_uint4(dk_[4 * c2], dk_[4 * c2 + 1], dk_[4 * c2 + 2], dk_[4 * c2 + 3]);
            }
            if (p.ds_spill != nullptr) {
              // spill this warp's 32 keys x CW queries through the tile it has just written: (32 / CPR) rows x CW*2 contiguous bytes
              // per store instruction (full 32-byte sectors) instead of 32 rows x 16 bytes straight from the registers
              __syncwarp();
              const uint8_t* tile_w = sDS + buf * 32768 + grp * 16384 + (q * 32) * 128;
              __nv_bfloat16* sp = p.ds_spill + ((long long)item * n + kb * 128 + q * 32) * n + qc * 64;
              // a 16-byte shared load is served per quarter-warp: the 8 lanes of a quarter take rows CPR apart, so that their
              // swizzled chunks (ch ^ (row & 7)) cover all eight 16-byte bank groups (no conflict; ncu r2e showed 2-way before)
              const int sp_row = ((lane & 7) / CPR) * CPR + ((lane >> 3) % CPR) + 8 * ((lane >> 3) / CPR);
#pragma unroll
              for (int j = 0; j < CPR; j++) {
                const int rr = j * (32 / CPR) + sp_row, ch = cq * CPR + lane % CPR;
                const uint4 val = *reinterpret_cast<const uint4*>(tile_w + rr * 128 + ((ch ^ (rr & 7)) << 4));
                __stcs(reinterpret_cast<uint4*>(sp + (long long)rr * n + ch * 8), val);
              }
            }
            tmem_st_wait();
          } else {
            // key rows beyond the sequence: their dS must be ZERO in the shared tile (dQ contracts over all 128 keys)
#pragma unroll
            for (int c4 = 0; c4 < CPR; c4++)
              *reinterpret_cast<uint4*>(tile_row + (((cq * CPR + c4) ^ (r & 7)) << 4)) = make_uint4(0, 0, 0, 0);
          }
          // drain the accumulators of the previous key block / item before this block's MMAs may overwrite them
          if (qc == 0 && u > 0) drain_dvdk(prev_item, prev_kb, (uint32_t)((u - 1) & 1));
          if (qc == 0 && kb == 0 && m > 0) drain_dq(prev_item, (uint32_t)((m - 1) & 1));
          fence_proxy_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[st]);
          if (grp == 1 || qc == NQC - 1) v++;
        }
        u++;
        prev_item = item;
        prev_kb = kb;
      }
      m++;
    }
    if (my_items > 0) {
      drain_dvdk(prev_item, prev_kb, (uint32_t)((u - 1) & 1));
      drain_dq(prev_item, (uint32_t)((m - 1) & 1));
    }
  } else if (warp == EW) {
    // ===================== producer (whole warp: lane 0 drives TMA, all lanes write the per-query records) =====================
    long long g = 0;
    int u = 0;
    const float inv_sc2 = 1.0f / sc2;
    (void)inv_sc2;
    for (int il = 0; il < my_items; il++) {
      const int item = (int)blockIdx.x + il * (int)gridDim.x;
      const int hd = item % p.heads;
      const long long row0 = (long long)(item / p.heads) * n;
      for (int kb = 0; kb < NKB; kb++, u++) {
        if (lane == 0) {
          const int ks = u & 1;
          mbar_wait_tag(&kv_empty[ks], (uint32_t)(((u >> 1) & 1) ^ 1), 50);
          mbar_arrive_expect_tx(&kv_full[ks], 16384);
          tma_load_2d(sKV + ks * 16384, &tk, &kv_full[ks], hd * 32, (int)(row0 + kb * 128));
          tma_load_2d(sKV + ks * 16384 + 8192, &tv, &kv_full[ks], hd * 32, (int)(row0 + kb * 128));
        }
        for (int qc = 0; qc < NQC; qc++, g++) {
          const int slot = (int)(g % TCB_QD_SLOTS);
          uint8_t* sl = sQD + slot * TCB_QD_BYTES;
          if (lane == 0) mbar_wait_tag(&qd_empty[slot], (uint32_t)(((g / TCB_QD_SLOTS) & 1) ^ 1), 51);
          __syncwarp();
          if (lane == 0) {
            mbar_arrive_expect_tx(&qd_full[slot], 8192);
            tma_load_2d(sl, &tq, &qd_full[slot], hd * 32, (int)(row0 + qc * 64));
            tma_load_2d(sl + 4096, &tdo, &qd_full[slot], hd * 32, (int)(row0 + qc * 64));
          }
          float4* rec = reinterpret_cast<float4*>(sl + 8192);
#pragma unroll
          for (int h2 = 0; h2 < 2; h2++) {
            const int i = qc * 64 + h2 * 32 + lane;          // < n (n % 64 == 0)
            const long long idx = (row0 + i) * p.heads + hd;
            if constexpr (AUG) {
              auto split3 = [](float v) {                    // v ~ hi + mid + lo, three bf16 terms (k columns 0..2 of the tile)
                const __nv_bfloat16 hi = __float2bfloat16(v);
                const float r1 = v - __bfloat162float(hi);
                const __nv_bfloat16 mid = __float2bfloat16(r1);
                const __nv_bfloat16 lo = __float2bfloat16(r1 - __bfloat162float(mid));
                return make_uint4((uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(mid) << 16),
                                  (uint32_t)__bfloat16_as_ushort(lo), 0u, 0u);
              };
              const int row = h2 * 32 + lane;
              const uint32_t off = (row >> 3) * 256 + (row & 7) * 16;          // k 0-7 core matrix of the row's 8-row group
              *reinterpret_cast<uint4*>(sl + 8192 + off) = split3(-__ldg(p.lse + idx) * inv_sc2);
              *reinterpret_cast<uint4*>(sl + 8192 + 2048 + off) = split3(-__ldg(p.delta + idx));
            } else {
              rec[h2 * 32 + lane] = make_float4(__ldg(p.lse + idx), __ldg(p.delta + idx), __int_as_float(4 * ((i / W) * STR + (i % W))), 0.f);
            }
          }
          if constexpr (AUG) fence_proxy_async_smem();       // generic-proxy tile writes -> tcgen05.mma operand reads
          mbar_arrive(&qd_full[slot]);
        }
      }
    }
  } else {
    // ===================== MMA issuer (all 32 lanes run the loop; tcgen05 instructions on the elected lane) =====================
    if (total_blocks > 0) {
      const bool leader = elect_one_sync();
      constexpr uint32_t idesc_s = umma_idesc(1, 0, 0, 128, 64);     // S^T, dP^T : A, B K-major
      constexpr uint32_t idesc_dv = umma_idesc(1, 0, 1, 128, 32);    // dV (A TMEM), dK (A smem K-major): B MN-major
      constexpr uint32_t idesc_dq = umma_idesc(1, 1, 1, 128, 32);    // dQ: A, B MN-major
      constexpr uint32_t hi64 = umma_desc_hi(512, SW64), hi128 = umma_desc_hi(1024, SW128);
      const uint32_t kv_lo_k = umma_desc_lo(smem_u32(sKV), 16), kv_lo_mn = umma_desc_lo(smem_u32(sKV), 512);
      const uint32_t qd_lo_k = umma_desc_lo(smem_u32(sQD), 16), qd_lo_mn = umma_desc_lo(smem_u32(sQD), 512);
      const uint32_t ds_lo_k = umma_desc_lo(smem_u32(sDS), 16), ds_lo_mn = umma_desc_lo(smem_u32(sDS), 16384);
      // augmented k-step tiles: no swizzle, K-major, LBO = 128 B (k 8-15 core matrix), SBO = 256 B (next 8 rows)
      constexpr uint32_t hi_ns = umma_desc_hi(256, 0);
      const uint32_t ones_lo = umma_desc_lo(smem_u32(sOnes), 128), qd_lo_ns = umma_desc_lo(smem_u32(sQD), 128);
      (void)ones_lo;
      (void)qd_lo_ns;
      // incrementally advanced state of the look-ahead S issue (s*) and of the consumer side (no divisions in this thread)
      int s_qc = 0, s_slot = 0, s_st = 0;
      uint32_t s_ub = 0, s_qpar = 0, s_round = 0;     // key-block counter, qd ring parity, (f >> 1) of the look-ahead block
      auto issue_s = [&]() {
        const int ks = (int)(s_ub & 1);
        if (s_qc == 0) mbar_wait_tag(&kv_full[ks], (s_ub >> 1) & 1, 60);
        mbar_wait_tag(&qd_full[s_slot], s_qpar, 61);
        tc_fence_after();
        const uint32_t kA = kv_lo_k + ks * (16384 >> 4), vA = kA + (8192 >> 4);
        const uint32_t qB = qd_lo_k + s_slot * (TCB_QD_BYTES >> 4), dB = qB + (4096 >> 4);
        const uint32_t eB = qd_lo_ns + s_slot * (TCB_QD_BYTES >> 4) + (8192 >> 4);     // (-lse/sc2) tile; the (-delta) tile 2 KB further
        (void)eB;
        const uint32_t d0 = tmem_base + s_st * 128;
        if (leader) {
          umma_bf16(d0, umma_desc_join(kA, hi64), umma_desc_join(qB, hi64), idesc_s, 0u);
          umma_bf16(d0, umma_desc_join(kA + 2, hi64), umma_desc_join(qB + 2, hi64), idesc_s, 1u);
          if constexpr (AUG) umma_bf16(d0, umma_desc_join(ones_lo, hi_ns), umma_desc_join(eB, hi_ns), idesc_s, 1u);
          umma_bf16(d0 + 64, umma_desc_join(vA, hi64), umma_desc_join(dB, hi64), idesc_s, 0u);
          umma_bf16(d0 + 64, umma_desc_join(vA + 2, hi64), umma_desc_join(dB + 2, hi64), idesc_s, 1u);
          if constexpr (AUG) umma_bf16(d0 + 64, umma_desc_join(ones_lo, hi_ns), umma_desc_join(eB + (2048 >> 4), hi_ns), idesc_s, 1u);
          umma_commit(&s_full[s_st]);
        }
        __syncwarp();
        s_st ^= 1;
        if (++s_slot == TCB_QD_SLOTS) { s_slot = 0; s_qpar ^= 1; }
        if (++s_qc == NQC) { s_qc = 0; s_ub++; }
        (void)s_round;
      };
      issue_s();
      int v = 0, qc = 0, kb = 0, slot = 0, st = 0;
      uint32_t ub = 0, ppar = 0;       // p_full parity of this stage use = (f >> 1) & 1
      for (long long f = 0; f < total_blocks; f++) {
        if (f + 1 < total_blocks) issue_s();
        const int ks = (int)(ub & 1);
        const int buf = v & 1, grp = qc & 1;
        mbar_wait_tag(&p_full[st], ppar, 62);
        tc_fence_after();
        const uint32_t qmn = qd_lo_mn + slot * (TCB_QD_BYTES >> 4), dmn = qmn + (4096 >> 4);
        const uint32_t tile_k = ds_lo_k + buf * (32768 >> 4) + grp * (16384 >> 4);
        const uint32_t acol = tmem_base + st * 128;
        const uint32_t acc0 = qc > 0 ? 1u : 0u;
        if (leader) {
        // dV += P^T dO_c : K = 64 queries, A = P^T in TMEM (16 queries = 8 columns; the two column halves sit 32 columns apart)
        // (queries 16*k4 .. +15 were written by warp slice (16*k4)/CW as its group ((16*k4) % CW)/16: 8 columns each)
#pragma unroll
        for (int k4 = 0; k4 < 4; k4++)
          umma_bf16_ts(tmem_base + TCB_COL_DV, acol + ((16 * k4) / CW) * CW + (((16 * k4) % CW) / 16) * 8,
                       umma_desc_join(dmn + k4 * (1024 >> 4), hi64), idesc_dv, k4 > 0 ? 1u : acc0);
        // dK += dS^T Q_c : A = dS tile group grp read K-major (128-byte rows), B = Q chunk MN-major
#pragma unroll
        for (int k4 = 0; k4 < 4; k4++)
          umma_bf16(tmem_base + TCB_COL_DK, umma_desc_join(tile_k + k4 * 2, hi128), umma_desc_join(qmn + k4 * (1024 >> 4), hi64), idesc_dv,
                    (k4 > 0) ? 1u : acc0);
        umma_commit(&qd_empty[slot]);
        if (grp == 1 || qc == NQC - 1) {     // dQ tile (qc/2) += dS K_blk : A = both groups of the tile, MN-major, K = 128 keys
          const uint32_t tile_mn = ds_lo_mn + buf * (32768 >> 4), kmn = kv_lo_mn + ks * (16384 >> 4);
          const uint32_t dqc = tmem_base + TCB_COL_DQ + (qc >> 1) * 32;
#pragma unroll
          for (int k8 = 0; k8 < 8; k8++)
            umma_bf16(dqc, umma_desc_join(tile_mn + k8 * (2048 >> 4), hi128), umma_desc_join(kmn + k8 * (1024 >> 4), hi64), idesc_dq,
                      (kb > 0 || k8 > 0) ? 1u : 0u);
          umma_commit(&ds_free[buf]);
        }
        if (qc == NQC - 1) {
          umma_commit(&kv_empty[ks]);
          umma_commit(acc_full);
          if (kb == NKB - 1) umma_commit(dq_full);
        }
        }   // leader
        __syncwarp();
        if (grp == 1 || qc == NQC - 1) v++;
        // advance
        if (st == 1) ppar ^= 1;
        st ^= 1;
        if (++slot == TCB_QD_SLOTS) slot = 0;
        if (++qc == NQC) {
          qc = 0;
          ub++;
          if (++kb == NKB) kb = 0;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == EW + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// dtable[rel(i,j), h] += sum over sequences of dS^T[seq, h][j][i]   (the spill of attn_tc_bwd_kernel).
// thread = 8 consecutive queries of one (head, key); block = 256 threads; bins of the block in shared memory, then one
// global atomic per touched bin.
__global__ void __launch_bounds__(256) attn_dtab_reduce_kernel(const __nv_bfloat16* __restrict__ ds, float* __restrict__ dtable,
                                                               int num_seqs, int heads, int H, int W) {
  extern __shared__ float s_bins[];
  const int n = H * W, R = (2 * H - 1) * (2 * W - 1);
  for (int i = threadIdx.x; i < R; i += 256) s_bins[i] = 0.f;
  __syncthreads();
  const int groups_per_head = n * (n / 8);
  const int blocks_per_head = (groups_per_head + 255) / 256;
  const int h = blockIdx.x / blocks_per_head;
  const int gidx = (blockIdx.x % blocks_per_head) * 256 + threadIdx.x;
  if (gidx < groups_per_head) {
    const int j = gidx / (n / 8), i0 = (gidx % (n / 8)) * 8;
    const long long nn = (long long)n * n;
    const __nv_bfloat16* src = ds + (long long)h * nn + (long long)j * n + i0;
    const long long stride = (long long)heads * nn;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; e++) acc[e] = 0.f;
#pragma unroll 8
    for (int sq = 0; sq < num_seqs; sq++) {
      const uint4 u = __ldcs(reinterpret_cast<const uint4*>(src + (long long)sq * stride));
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const float2 f = unpack_bf16x2(w[e]);
        acc[2 * e] += f.x;
        acc[2 * e + 1] += f.y;
      }
    }
    const int yj = j / W, xj = j % W;
#pragma unroll
    for (int e = 0; e < 8; e++) {
      const int i = i0 + e;
      const int rr = (i / W - yj + H - 1) * (2 * W - 1) + (i % W - xj + W - 1);
      atomicAdd(&s_bins[rr], acc[e]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < R; i += 256) {
    const float vv = s_bins[i];
    if (vv != 0.f) atomicAdd(dtable + (long long)i * heads + h, vv);
  }
}

__global__ void qk_bound_kernel(const float* __restrict__ qs, const float* __restrict__ ks, int dh, float* __restrict__ out) {
  float m = 0.f;
  for (int d = threadIdx.x; d < dh; d += 32) m = fmaxf(m, fabsf(qs[d] * ks[d]));
  m = warp_max(m);
  if (threadIdx.x == 0) out[0] = m;
}

}  // namespace ctb

using namespace ctb;

extern "C" int ctclip_qk_bound(const float* q_scale, const float* k_scale, int32_t dim_head, float* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(q_scale && k_scale && out && dim_head > 0, "qk_bound: bad args");
  qk_bound_kernel<<<1, 32, 0, stream>>>(q_scale, k_scale, dim_head, out);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}

static int tc_grid_w(const ctclip_attn_args* a) { return a->grid_w; }

extern "C" int ctclip_attn_tc_supported(int32_t n, int32_t grid_h, int32_t grid_w, int32_t dim_head) {
  if (dim_head != 32 || grid_h <= 0 || grid_w <= 0 || n != grid_h * grid_w) return 0;
  // bit 0: forward kernel, bit 1: backward kernel (its dQ accumulators need ceil(n/128)*32 <= 192 TMEM columns)
  // (n % 64: K / V / Q-chunk TMA boxes are 64 rows; n % 96 resp. 64: whole key chunks)
  if (grid_w == 24 && n % 192 == 0 && n <= 1152) return 1 | (((n + 127) / 128 * 32 <= 192) ? 2 : 0);
  if (grid_w == 32 && n % 64 == 0 && n <= 1024) return 1;
  return 0;
}

template <int NCH, int W>
static int launch_tc_fwd(const ctclip_attn_args* a, cudaStream_t stream) {
  const long long rows = (long long)a->num_seqs * a->n;
  CUtensorMap tq, tk, tv;
  const uint64_t inner = (uint64_t)a->heads * 32;
  if (int rc = encode_tmap_2d(&tq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->q, inner, (uint64_t)rows, (uint64_t)a->ldq * 2, 32, 128,
                              CU_TENSOR_MAP_SWIZZLE_64B))
    return rc;
  if (int rc = encode_tmap_2d(&tk, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->k, inner, (uint64_t)rows, (uint64_t)a->ldk * 2, 32, 64,
                              CU_TENSOR_MAP_SWIZZLE_64B))
    return rc;
  if (int rc = encode_tmap_2d(&tv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->v, inner, (uint64_t)rows, (uint64_t)a->ldv * 2, 32, 64,
                              CU_TENSOR_MAP_SWIZZLE_64B))
    return rc;
  TcFwdParams p;
  p.n = a->n; p.heads = a->heads; p.H = a->grid_h; p.num_seqs = a->num_seqs;
  p.table = a->cpb_table; p.qk_bound = a->qk_bound; p.scale = a->scale;
  p.o = reinterpret_cast<__nv_bfloat16*>(a->o); p.ldo = a->ldo; p.lse = a->lse;
  const int tab_elems = ((2 * a->grid_h - 1) * TcGeom<W>::STR + 31) & ~31;
  const size_t smem = 1024 + (size_t)a->n * 128 + 2 * 8192 + (size_t)tab_elems * 4 + 512 * 4 + 12 * 8 + 8 + 16 * 4 + 64;
  CTB_CHECK_ARG(smem <= 227 * 1024, "attn_fwd(tc): %zu B of shared memory needed", smem);
  auto kern = attn_tc_fwd_kernel<NCH, W>;
  CTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<a->num_seqs * a->heads, 320, smem, stream>>>(tq, tk, tv, p);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}

// Forward on the tcgen05 path. Called by ctclip_attn_fwd when args->cpb_table (or args->tc_path) selects it.
int ctb_attn_fwd_tc(const ctclip_attn_args* a, cudaStream_t stream) {
  CTB_CHECK_ARG(a->q && a->k && a->v && a->o, "attn_fwd(tc): null pointer");
  CTB_CHECK_ARG(ctclip_attn_tc_supported(a->n, a->grid_h, a->grid_w, a->dim_head),
                "attn_fwd(tc): unsupported geometry n=%d grid=%dx%d dim_head=%d", a->n, a->grid_h, a->grid_w, a->dim_head);
  CTB_CHECK_ARG(a->key_mask == nullptr, "attn_fwd(tc): no key mask on this path");
  CTB_CHECK_ARG(a->seq_inner == 1 && a->tok_stride == 1 && a->seq_outer_stride == a->n, "attn_fwd(tc): sequences must be contiguous");
  CTB_CHECK_ARG(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0 && a->ldo % 8 == 0, "attn_fwd(tc): rows must be 16B aligned");
  CTB_CHECK_ARG((reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) | reinterpret_cast<uintptr_t>(a->v) |
                 reinterpret_cast<uintptr_t>(a->o)) % 16 == 0, "attn_fwd(tc): q/k/v/o must be 16B aligned");
  if (tc_grid_w(a) == 24) return launch_tc_fwd<96, 24>(a, stream);
  return launch_tc_fwd<64, 32>(a, stream);
}

static int g_tc_bwd_warps = 8, g_tc_bwd_aug = 1;

template <int EW, bool AUG>
static int launch_tc_bwd(const ctclip_attn_args* a, cudaStream_t stream) {
  const long long rows = (long long)a->num_seqs * a->n;
  CUtensorMap tq, tdo, tk, tv;
  const uint64_t inner = (uint64_t)a->heads * 32;
  if (int rc = encode_tmap_2d(&tq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->q, inner, (uint64_t)rows, (uint64_t)a->ldq * 2, 32, 64,
                              CU_TENSOR_MAP_SWIZZLE_64B))
    return rc;
  if (int rc = encode_tmap_2d(&tdo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->d_o, inner, (uint64_t)rows, (uint64_t)a->ldo * 2, 32, 64,
                              CU_TENSOR_MAP_SWIZZLE_64B))
    return rc;
  if (int rc = encode_tmap_2d(&tk, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->k, inner, (uint64_t)rows, (uint64_t)a->ldk * 2, 32, 128,
                              CU_TENSOR_MAP_SWIZZLE_64B))
    return rc;
  if (int rc = encode_tmap_2d(&tv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->v, inner, (uint64_t)rows, (uint64_t)a->ldv * 2, 32, 128,
                              CU_TENSOR_MAP_SWIZZLE_64B))
    return rc;
  TcBwdParams p;
  p.n = a->n; p.heads = a->heads; p.H = a->grid_h; p.num_seqs = a->num_seqs;
  p.table = a->cpb_table; p.scale = a->scale; p.lse = a->lse; p.delta = a->delta;
  p.dq = reinterpret_cast<__nv_bfloat16*>(a->dq); p.ld_dq = a->ld_dq;
  p.dk = reinterpret_cast<__nv_bfloat16*>(a->dk); p.ld_dk = a->ld_dk;
  p.dv = reinterpret_cast<__nv_bfloat16*>(a->dv); p.ld_dv = a->ld_dv;
  p.ds_spill = (a->dcpb_table != nullptr) ? reinterpret_cast<__nv_bfloat16*>(a->ds_scratch) : nullptr;
  p.dbias_t = a->dbias;      // (with cpb_table set) fp32 [heads, n keys, n queries], accumulated by reductions; see ctclip_cpb_reduce_t
  const int tab_elems = ((2 * a->grid_h - 1) * TcGeom<24>::STR + 31) & ~31;
  const size_t smem = 1024 + 2 * 16384 + 2 * 32768 + (size_t)TCB_QD_SLOTS * TCB_QD_BYTES + 4096 + (size_t)tab_elems * 4 + 20 * 8 + 8 + 24 * 4 + 64;
  CTB_CHECK_ARG(smem <= 227 * 1024, "attn_bwd(tc): %zu B of shared memory needed", smem);
  auto kern = attn_tc_bwd_kernel<24, EW, AUG>;
  CTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // persistent grid: a multiple of `heads` CTAs so that every item a CTA walks (blockIdx.x + k*gridDim.x) has the same head
  const int items = a->num_seqs * a->heads;
  int grid = (num_sms() / a->heads) * a->heads;
  if (grid < a->heads) grid = a->heads;
  if (grid > items) grid = items;
  kern<<<grid, (EW + 2) * 32, smem, stream>>>(tq, tdo, tk, tv, p);
  CTB_LAUNCH_CHECK();
  if (a->dcpb_table != nullptr) {
    const int n = a->n, R = (2 * a->grid_h - 1) * (2 * a->grid_w - 1);
    const int blocks_per_head = (n * (n / 8) + 255) / 256;
    attn_dtab_reduce_kernel<<<blocks_per_head * a->heads, 256, R * sizeof(float), stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(a->ds_scratch), a->dcpb_table, a->num_seqs, a->heads, a->grid_h, a->grid_w);
    CTB_LAUNCH_CHECK();
  }
  return CTCLIP_OK;
}

// Backward on the tcgen05 path (delta already computed by the caller, ctclip_attn_bwd).
int ctb_attn_bwd_tc(const ctclip_attn_args* a, cudaStream_t stream) {
  CTB_CHECK_ARG(a->q && a->k && a->v && a->d_o && a->lse && a->delta && a->dq && a->dk && a->dv, "attn_bwd(tc): null pointer");
  CTB_CHECK_ARG((ctclip_attn_tc_supported(a->n, a->grid_h, a->grid_w, a->dim_head) & 2) != 0,
                "attn_bwd(tc): unsupported geometry n=%d grid=%dx%d dim_head=%d", a->n, a->grid_h, a->grid_w, a->dim_head);
  CTB_CHECK_ARG(a->key_mask == nullptr, "attn_bwd(tc): no key mask on this path");
  CTB_CHECK_ARG(a->seq_inner == 1 && a->tok_stride == 1 && a->seq_outer_stride == a->n, "attn_bwd(tc): sequences must be contiguous");
  CTB_CHECK_ARG(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0 && a->ldo % 8 == 0 && a->ld_dq % 8 == 0 && a->ld_dk % 8 == 0 &&
                    a->ld_dv % 8 == 0, "attn_bwd(tc): rows must be 16B aligned");
  CTB_CHECK_ARG(a->dcpb_table == nullptr || a->ds_scratch != nullptr, "attn_bwd(tc): dcpb_table needs ds_scratch");
  CTB_CHECK_ARG(a->dbias == nullptr || a->dcpb_table == nullptr, "attn_bwd(tc): give dcpb_table (+ ds_scratch) OR dbias (transposed fp32 table)");
  if (g_tc_bwd_aug) return (g_tc_bwd_warps == 8) ? launch_tc_bwd<8, true>(a, stream) : launch_tc_bwd<16, true>(a, stream);
  return (g_tc_bwd_warps == 8) ? launch_tc_bwd<8, false>(a, stream) : launch_tc_bwd<16, false>(a, stream);
}

// Measurement knob (tools/attn_tc_probe.py): number of softmax-backward warps of the tcgen05 backward kernel (8 or 16); +100 selects
// the variant with the per-query terms folded into the tensor-core products (AUG).
extern "C" int ctclip_debug_set_attn_bwd_warps(int32_t warps) {
  const int w = warps % 100;
  CTB_CHECK_ARG((w == 8 || w == 16) && (warps == w || warps == 100 + w), "attn bwd variant must be 8, 16, 108 or 116");
  g_tc_bwd_warps = w;
  g_tc_bwd_aug = warps >= 100;
  return CTCLIP_OK;
}
