// Cosine-similarity multi-head attention core, forward and backward, flash-style (scores never
// leave the SM). Replaces attention.py:156-178:
//     sim = q_hat k_hat^T * scale (+ attn_bias[h,i,j]);  attn = softmax(sim);  out = attn v
// where q_hat / k_hat are the l2-normalised, per-channel-scaled projections (produced by the
// L2NORM epilogue of the projection GEMM), plus the autograd backward of the same ops.
//
// Why warp-level mma.sync (m16n8k16 bf16) and not tcgen05 here: with dim_head = 32 a 128x576 score
// block needs 73.7k exponentials (16/clk/SM on the SFU = 4.6k cycles) but only 9.4 MFLOP of
// tensor work (1.15k cycles at tcgen05 rate, ~2.3k with mma.sync). The kernel is SFU-bound either
// way; the tcgen05 variant would add TMEM<->register round trips per KV block for no gain.
// ("attention roofline".)
//
// Token addressing is strided so that both factorised stacks run on the single canonical
// [b,t,h,w,D] layout: row(seq, i) = (seq / seq_inner) * seq_outer_stride + (seq % seq_inner) + i * tok_stride.
//   spatial : seq=(b,t), i=(h,w): seq_inner=1, seq_outer_stride=S, tok_stride=1
//   temporal: seq=(b,h,w), i=t  : seq_inner=S, seq_outer_stride=T*S, tok_stride=S
#include <stdlib.h>
#include "common.cuh"
#include "ptx.cuh"
#include "rng.cuh"
#include "../../include/ctclip_b200.h"

namespace ctb {

constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct AttnGeom {
  int n, heads, seq_inner;
  long long seq_outer_stride, tok_stride;
  __device__ __forceinline__ long long row(int seq, int i) const {
    return (long long)(seq / seq_inner) * seq_outer_stride + (seq % seq_inner) + (long long)i * tok_stride;
  }
};

template <int WPG>
__device__ __forceinline__ void group_sync(int group) {
  if (WPG == 1) __syncwarp();
  else asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(WPG * 32) : "memory");
}

// DH = dim_head: 32 for the CTViT stacks (run_train.py:25), 64 for the BERT text tower.
// KROW = DH + 8: padded row (bf16 elements) of row-major [token][d] tiles -> conflict-free fragment loads.
#define KROW (DH + 8)

// Row-major tile loader: dst[n_pad][KROW] <- src rows (64 B each), zero-filled beyond n.
template <int DH>
__device__ __forceinline__ void load_rows(__nv_bfloat16* dst, const __nv_bfloat16* src, long long ld, int head,
                                          const AttnGeom& g, int seq, int n_pad, int tid, int nthreads) {
  constexpr int PARTS = DH / 8;
  for (int idx = tid; idx < n_pad * PARTS; idx += nthreads) {
    const int r = idx / PARTS, part = idx % PARTS;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (r < g.n) val = *reinterpret_cast<const uint4*>(src + g.row(seq, r) * ld + head * DH + part * 8);
    *reinterpret_cast<uint4*>(dst + r * KROW + part * 8) = val;
  }
}
// Transposed tile loader: dst[DH][tstride] <- src rows, zero-filled beyond n.
template <int DH>
__device__ __forceinline__ void load_rows_t(__nv_bfloat16* dst, int tstride, const __nv_bfloat16* src, long long ld,
                                            int head, const AttnGeom& g, int seq, int n_pad, int tid, int nthreads) {
  constexpr int PARTS = DH / 8;
  for (int idx = tid; idx < n_pad * PARTS; idx += nthreads) {
    const int r = idx / PARTS, part = idx % PARTS;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (r < g.n) val = *reinterpret_cast<const uint4*>(src + g.row(seq, r) * ld + head * DH + part * 8);
    const __nv_bfloat16* e = reinterpret_cast<const __nv_bfloat16*>(&val);
#pragma unroll
    for (int i = 0; i < 8; i++) dst[(part * 8 + i) * tstride + r] = e[i];
  }
}
// A-operand fragments (16 rows x DH) straight from global memory; rows >= n read as zero.
template <int DH>
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[DH / 16][4], const __nv_bfloat16* src, long long ld,
                                             int head, const AttnGeom& g, int seq, int r0, int lane) {
  const int gq = lane >> 2, t = lane & 3;
  const int ra = r0 + gq, rb = r0 + gq + 8;
  const __nv_bfloat16* pa = src + g.row(seq, ra < g.n ? ra : 0) * ld + head * DH;
  const __nv_bfloat16* pb = src + g.row(seq, rb < g.n ? rb : 0) * ld + head * DH;
#pragma unroll
  for (int kt = 0; kt < DH / 16; kt++) {
    a[kt][0] = ra < g.n ? *reinterpret_cast<const uint32_t*>(pa + kt * 16 + 2 * t) : 0u;
    a[kt][1] = rb < g.n ? *reinterpret_cast<const uint32_t*>(pb + kt * 16 + 2 * t) : 0u;
    a[kt][2] = ra < g.n ? *reinterpret_cast<const uint32_t*>(pa + kt * 16 + 2 * t + 8) : 0u;
    a[kt][3] = rb < g.n ? *reinterpret_cast<const uint32_t*>(pb + kt * 16 + 2 * t + 8) : 0u;
  }
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}

// C[16 x 8*NT] = A[16 x DH] * rows(tile)[key0 .. key0+8*NT)^T  with tile row-major [key][KROW].
// B fragments via ldmatrix.x4: one instruction yields (b0,b1) of two consecutive 16-wide k-steps of one n-tile.
template <int NT, int DH, bool FULL = false>
__device__ __forceinline__ void qk_block(float (&s)[NT][4], const uint32_t (&a)[DH / 16][4],
                                         const __nv_bfloat16* tile, int key0, int lane, int nt_valid = NT) {
#pragma unroll
  for (int nt = 0; nt < NT; nt++) {
    s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
    if (!FULL && nt >= nt_valid) continue;  // warp-uniform; tiles beyond the padded sequence stay zero
    const __nv_bfloat16* base = tile + (key0 + nt * 8 + (lane & 7)) * KROW + (lane >> 3) * 8;
#pragma unroll
    for (int kp = 0; kp < DH / 32; kp++) {
      uint32_t b[4];
      ldsm_x4(b, base + kp * 32);
      mma_16816(s[nt], a[2 * kp], b[0], b[1]);
      mma_16816(s[nt], a[2 * kp + 1], b[2], b[3]);
    }
  }
}
// acc[16 x DH] += P[16 x 8*NT] * X[key0 .. key0+8*NT, :]  with X row-major [key][KROW] (the same tile layout as above):
// ldmatrix.x4.trans delivers the (k = key, n = d) B fragments straight from the row-major tile, so no transposed copy of
// V / K / Q / dO is kept in shared memory.
template <int NT, int DH, bool FULL = false>
__device__ __forceinline__ void pv_block(float (&acc)[DH / 8][4], const float (&p)[NT][4],
                                         const __nv_bfloat16* tile, int key0, int lane, int nt_valid = NT) {
#pragma unroll
  for (int kk = 0; kk < NT / 2; kk++) {
    if (!FULL && 2 * kk >= nt_valid) continue;
    uint32_t a[4];
    a[0] = pack_bf16x2(p[2 * kk][0], p[2 * kk][1]);
    a[1] = pack_bf16x2(p[2 * kk][2], p[2 * kk][3]);
    a[2] = pack_bf16x2(p[2 * kk + 1][0], p[2 * kk + 1][1]);
    a[3] = pack_bf16x2(p[2 * kk + 1][2], p[2 * kk + 1][3]);
    const int mi = lane >> 3;
    const __nv_bfloat16* base = tile + (key0 + kk * 16 + (mi & 1) * 8 + (lane & 7)) * KROW + (mi >> 1) * 8;
#pragma unroll
    for (int dp = 0; dp < DH / 16; dp++) {
      uint32_t b[4];
      ldsm_x4_trans(b, base + dp * 16);
      mma_16816(acc[2 * dp], a, b[0], b[1]);
      mma_16816(acc[2 * dp + 1], a, b[2], b[3]);
    }
  }
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// One 16 x 8*NT block of raw dot products -> log2-domain logits  s*sc2 + bias*log2(e);  columns >= n, masked keys and
// tiles >= nt_valid become -inf (=> probability 0). brow_a / brow_b: bias rows of the two fragment rows (nullptr: none).
// bfrag (optional): this thread's NT*4 bias values in MMA-fragment order (bf16, 8 bytes per n-tile) -- the layout
// ctclip_cpb_expand writes so that a warp reads its whole 16 x 64 bias block with four fully coalesced 16-byte loads
// per lane (the natural [h,i,j] layout costs 16 loads per lane at 25 % sector efficiency).
template <int NT, bool MASK>
__device__ __forceinline__ void logits_tile(float (&s)[NT][4], float sc2, const __nv_bfloat16* brow_a,
                                            const __nv_bfloat16* brow_b, int c0, int t, int n, bool cols_full, bool pair_ok,
                                            const int* sMask, int nt_valid, const uint2* bfrag = nullptr) {
  uint2 bf[NT];
  if (bfrag != nullptr) {
#pragma unroll
    for (int i = 0; i < NT / 2; i++) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(bfrag) + i);
      bf[2 * i] = make_uint2(u.x, u.y);
      bf[2 * i + 1] = make_uint2(u.z, u.w);
    }
  }
#pragma unroll
  for (int nt = 0; nt < NT; nt++) {
    if (nt >= nt_valid) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = -INFINITY;
      continue;
    }
    const int c = c0 + nt * 8 + 2 * t;
    const bool ok0 = cols_full || c < n, ok1 = cols_full || c + 1 < n;
    float ba0 = 0.f, ba1 = 0.f, bb0 = 0.f, bb1 = 0.f;
    if (bfrag != nullptr) {   // fragment tables already hold bias * log2(e)
      const float2 fa = unpack_bf16x2(bf[nt].x), fb = unpack_bf16x2(bf[nt].y);
      ba0 = fa.x; ba1 = fa.y; bb0 = fb.x; bb1 = fb.y;
    } else if (brow_a != nullptr) {
      if (ok1 && pair_ok) {
        const float2 f = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(brow_a + c));
        ba0 = f.x; ba1 = f.y;
      } else {
        if (ok0) ba0 = __bfloat162float(brow_a[c]);
        if (ok1) ba1 = __bfloat162float(brow_a[c + 1]);
      }
    }
    if (bfrag == nullptr && brow_b != nullptr) {
      if (ok1 && pair_ok) {
        const float2 f = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(brow_b + c));
        bb0 = f.x; bb1 = f.y;
      } else {
        if (ok0) bb0 = __bfloat162float(brow_b[c]);
        if (ok1) bb1 = __bfloat162float(brow_b[c + 1]);
      }
    }
    const float bsc = (bfrag != nullptr) ? 1.0f : kLog2e;   // warp-uniform
    s[nt][0] = fmaf(s[nt][0], sc2, ba0 * bsc);
    s[nt][1] = fmaf(s[nt][1], sc2, ba1 * bsc);
    s[nt][2] = fmaf(s[nt][2], sc2, bb0 * bsc);
    s[nt][3] = fmaf(s[nt][3], sc2, bb1 * bsc);
    bool m0 = !ok0, m1 = !ok1;
    if (MASK) {
      if (ok0 && sMask[c]) m0 = true;
      if (ok1 && sMask[c + 1]) m1 = true;
    }
    if (m0) s[nt][0] = s[nt][2] = -INFINITY;
    if (m1) s[nt][1] = s[nt][3] = -INFINITY;
  }
}

// Branch-free variant for blocks that lie completely inside the sequence and carry no key mask (every block of the
// 576-token spatial sequences): logits = s*sc2 (+ fragment-ordered bias, which ctclip_cpb_expand_frag stores already
// multiplied by log2 e).
template <int NT>
__device__ __forceinline__ void logits_tile_full(float (&s)[NT][4], float sc2, const uint2* bfrag) {
  if (bfrag != nullptr) {
#pragma unroll
    for (int i = 0; i < NT / 2; i++) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(bfrag) + i);
      const float2 a0 = unpack_bf16x2(u.x), b0 = unpack_bf16x2(u.y), a1 = unpack_bf16x2(u.z), b1 = unpack_bf16x2(u.w);
      s[2 * i][0] = fmaf(s[2 * i][0], sc2, a0.x);
      s[2 * i][1] = fmaf(s[2 * i][1], sc2, a0.y);
      s[2 * i][2] = fmaf(s[2 * i][2], sc2, b0.x);
      s[2 * i][3] = fmaf(s[2 * i][3], sc2, b0.y);
      s[2 * i + 1][0] = fmaf(s[2 * i + 1][0], sc2, a1.x);
      s[2 * i + 1][1] = fmaf(s[2 * i + 1][1], sc2, a1.y);
      s[2 * i + 1][2] = fmaf(s[2 * i + 1][2], sc2, b1.x);
      s[2 * i + 1][3] = fmaf(s[2 * i + 1][3], sc2, b1.y);
    }
  } else {
#pragma unroll
    for (int nt = 0; nt < NT; nt++) {
      s[nt][0] *= sc2; s[nt][1] *= sc2; s[nt][2] *= sc2; s[nt][3] *= sc2;
    }
  }
}

// Same with the bias fragments already in registers (loaded one key block ahead: the L2 latency of the table used to be
// the top stall of the forward / dQ / dKV kernels).
template <int NT>
__device__ __forceinline__ void logits_tile_full_regs(float (&s)[NT][4], float sc2, const uint4 (&b)[NT / 2]) {
#pragma unroll
  for (int i = 0; i < NT / 2; i++) {
    const float2 a0 = unpack_bf16x2(b[i].x), b0 = unpack_bf16x2(b[i].y), a1 = unpack_bf16x2(b[i].z), b1 = unpack_bf16x2(b[i].w);
    s[2 * i][0] = fmaf(s[2 * i][0], sc2, a0.x);
    s[2 * i][1] = fmaf(s[2 * i][1], sc2, a0.y);
    s[2 * i][2] = fmaf(s[2 * i][2], sc2, b0.x);
    s[2 * i][3] = fmaf(s[2 * i][3], sc2, b0.y);
    s[2 * i + 1][0] = fmaf(s[2 * i + 1][0], sc2, a1.x);
    s[2 * i + 1][1] = fmaf(s[2 * i + 1][1], sc2, a1.y);
    s[2 * i + 1][2] = fmaf(s[2 * i + 1][2], sc2, b1.x);
    s[2 * i + 1][3] = fmaf(s[2 * i + 1][3], sc2, b1.y);
  }
}


// Attention-probability dropout (BERT): keep flags of the element pair (row, col), (row, col + 1) of item `item` (col even:
// both elements sit in the same Philox group of 4). Returns the multipliers (0 or 1/(1-p)).
struct AttnDrop {
  unsigned long long seed, offset;
  uint32_t thresh;
  float inv_keep;
  int n;
  __device__ __forceinline__ float2 pair(long long item, int row, int col) const {
    const unsigned long long idx = ((unsigned long long)item * n + row) * n + col;
    uint32_t w[4];
    philox4(seed, offset + (idx >> 2), w);
    const int e = (int)(idx & 3);
    float2 r;
    r.x = (w[e] >= thresh) ? inv_keep : 0.f;
    r.y = (w[(e + 1) & 3] >= thresh && e < 3) ? inv_keep : 0.f;   // e is 0 or 2 when n is even and col is even
    return r;
  }
  __device__ __forceinline__ float one(long long item, int row, int col) const {
    const unsigned long long idx = ((unsigned long long)item * n + row) * n + col;
    uint32_t w[4];
    philox4(seed, offset + (idx >> 2), w);
    return (w[idx & 3] >= thresh) ? inv_keep : 0.f;
  }
};
__device__ __forceinline__ AttnDrop make_attn_drop(const ctclip_attn_args& a) {
  AttnDrop d;
  d.seed = a.dropout_seed; d.offset = a.dropout_offset;
  d.thresh = dropout_threshold(a.dropout_p);
  d.inv_keep = 1.f / (1.f - a.dropout_p);
  d.n = a.n;
  return d;
}

__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int DH, int WPG, int GROUPS, bool MASK, bool DROP = false>
__global__ void __launch_bounds__(WPG* GROUPS * 32, 2) attn_fwd_kernel(ctclip_attn_args a) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const AttnGeom g{a.n, a.heads, a.seq_inner, a.seq_outer_stride, a.tok_stride};
  const int n_pad = (a.n + 15) & ~15;
  const int tstride = n_pad + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = warp / WPG, wig = warp % WPG;
  const int gq = lane >> 2, t = lane & 3;
  const long long item = (long long)blockIdx.x * GROUPS + group;
  const size_t group_bytes = (size_t)(2 * n_pad * KROW) * 2 + (MASK ? (size_t)n_pad * 4 : 0);
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem_attn + group * group_bytes);
  __nv_bfloat16* sV = sK + n_pad * KROW;
  int* sMask = reinterpret_cast<int*>(sV + n_pad * KROW);   // 1 = key is masked out (only when MASK)
  const bool active = item < (long long)a.num_seqs * a.heads;
  const int head = active ? (int)(item % a.heads) : 0;
  const int seq = active ? (int)(item / a.heads) : 0;
  const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(a.q);
  const __nv_bfloat16* k = reinterpret_cast<const __nv_bfloat16*>(a.k);
  const __nv_bfloat16* v = reinterpret_cast<const __nv_bfloat16*>(a.v);
  const __nv_bfloat16* bias = reinterpret_cast<const __nv_bfloat16*>(a.bias);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.o);
  if (active) {
    load_rows<DH>(sK, k, a.ldk, head, g, seq, n_pad, wig * 32 + lane, WPG * 32);
    load_rows<DH>(sV, v, a.ldv, head, g, seq, n_pad, wig * 32 + lane, WPG * 32);
    if (MASK)
      for (int i = wig * 32 + lane; i < n_pad; i += WPG * 32)
        sMask[i] = (i < a.n && a.key_mask[(long long)seq * a.n + i] != 0) ? 0 : 1;
  }
  group_sync<WPG>(group);
  if (!active) return;
  const float sc2 = a.scale * kLog2e;
  const int row_tiles = n_pad / 16;
  const int kblocks = (n_pad + 63) / 64;
  const uint2* bias_frag = reinterpret_cast<const uint2*>(a.bias_frag);
  for (int rt = wig; rt < row_tiles; rt += WPG) {
    const int r0 = rt * 16;
    uint32_t qa[DH / 16][4];
    load_a_frags<DH>(qa, q, a.ldq, head, g, seq, r0, lane);
    const int ra = r0 + gq, rb = r0 + gq + 8;
    const bool pair_ok = (a.n & 1) == 0;
    const __nv_bfloat16* brow_a = (bias != nullptr && ra < a.n) ? bias + ((long long)head * a.n + ra) * a.n : nullptr;
    const __nv_bfloat16* brow_b = (bias != nullptr && rb < a.n) ? bias + ((long long)head * a.n + rb) * a.n : nullptr;
    float m_a = -INFINITY, m_b = -INFINITY, l_a = 0.f, l_b = 0.f;
    float oacc[DH / 8][4];
#pragma unroll
    for (int dt = 0; dt < DH / 8; dt++) oacc[dt][0] = oacc[dt][1] = oacc[dt][2] = oacc[dt][3] = 0.f;
    const bool use_frag = !MASK && bias_frag != nullptr;
    const uint2* bfr_row = use_frag ? bias_frag + (((long long)head * row_tiles + rt) * kblocks * 32 + lane) * 8 : nullptr;
    uint4 bcur[4];
    if (use_frag && 64 <= a.n) {
#pragma unroll
      for (int i = 0; i < 4; i++) bcur[i] = __ldg(reinterpret_cast<const uint4*>(bfr_row) + i);
    }
    for (int key0 = 0; key0 < n_pad; key0 += 64) {
      float s[8][4];
      const int rem = n_pad - key0;  // multiple of 16
      const int ntv = rem >= 64 ? 8 : rem / 8;
      const uint2* bfr = bias_frag ? bias_frag + ((((long long)head * row_tiles + rt) * kblocks + (key0 >> 6)) * 32 + lane) * 8
                                   : nullptr;
      const bool full = !MASK && (key0 + 64 <= a.n) && (bias == nullptr || bias_frag != nullptr);   // warp-uniform
      const bool nfull = use_frag && (key0 + 128 <= a.n);   // the next key block is complete too
      if (full) {
        qk_block<8, DH, true>(s, qa, sK, key0, lane);
        if (use_frag) logits_tile_full_regs<8>(s, sc2, bcur);
        else logits_tile_full<8>(s, sc2, nullptr);
        if (nfull) {   // refill the (now dead) fragment registers for the next block: the exps and PV MMAs below hide the L2 latency
#pragma unroll
          for (int i = 0; i < 4; i++) bcur[i] = __ldg(reinterpret_cast<const uint4*>(bfr + 256) + i);
        }
      } else {
        qk_block<8, DH>(s, qa, sK, key0, lane, ntv);
        logits_tile<8, MASK>(s, sc2, brow_a, brow_b, key0, t, a.n, key0 + 64 <= a.n, pair_ok, sMask, ntv, bfr);
      }
      float bm_a = -INFINITY, bm_b = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; nt++) {
        bm_a = fmaxf(bm_a, fmaxf(s[nt][0], s[nt][1]));
        bm_b = fmaxf(bm_b, fmaxf(s[nt][2], s[nt][3]));
      }
      bm_a = quad_max(bm_a);
      bm_b = quad_max(bm_b);
      const float mn_a = fmaxf(m_a, bm_a), mn_b = fmaxf(m_b, bm_b);
      // a fully masked prefix keeps the running max at -inf: use 0 as the reference there (all terms are exp2(-inf) = 0)
      const float rf_a = (mn_a == -INFINITY) ? 0.f : mn_a, rf_b = (mn_b == -INFINITY) ? 0.f : mn_b;
      const float corr_a = fast_exp2(m_a - rf_a), corr_b = fast_exp2(m_b - rf_b);
      m_a = mn_a;
      m_b = mn_b;
      float ps_a = 0.f, ps_b = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; nt++) {
        s[nt][0] = fast_exp2(s[nt][0] - rf_a);
        s[nt][1] = fast_exp2(s[nt][1] - rf_a);
        s[nt][2] = fast_exp2(s[nt][2] - rf_b);
        s[nt][3] = fast_exp2(s[nt][3] - rf_b);
        ps_a += s[nt][0] + s[nt][1];
        ps_b += s[nt][2] + s[nt][3];
      }
      l_a = l_a * corr_a + ps_a;
      l_b = l_b * corr_b + ps_b;
#pragma unroll
      for (int dt = 0; dt < DH / 8; dt++) {
        oacc[dt][0] *= corr_a; oacc[dt][1] *= corr_a;
        oacc[dt][2] *= corr_b; oacc[dt][3] *= corr_b;
      }
      if (DROP) {   // the row sums above use the un-dropped probabilities (softmax first, dropout second)
        const AttnDrop dr = make_attn_drop(a);
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
          const int c = key0 + nt * 8 + 2 * t;
          if (c < a.n) {
            const float2 ka = dr.pair(item, ra < a.n ? ra : 0, c), kb = dr.pair(item, rb < a.n ? rb : 0, c);
            s[nt][0] *= ka.x; s[nt][1] *= ka.y;
            s[nt][2] *= kb.x; s[nt][3] *= kb.y;
          }
        }
      }
      // masked / beyond-n_pad keys carry p == 0 and V^T rows are zero-filled there
      if (full) pv_block<8, DH, true>(oacc, s, sV, key0, lane);
      else pv_block<8, DH>(oacc, s, sV, key0, lane, ntv);
    }
    l_a = quad_sum(l_a);
    l_b = quad_sum(l_b);
    const float inv_a = 1.f / l_a, inv_b = 1.f / l_b;
    if (ra < a.n) {
      __nv_bfloat16* orow = o + g.row(seq, ra) * a.ldo + head * DH + 2 * t;
#pragma unroll
      for (int dt = 0; dt < DH / 8; dt++)
        *reinterpret_cast<uint32_t*>(orow + dt * 8) = pack_bf16x2(oacc[dt][0] * inv_a, oacc[dt][1] * inv_a);
      if (t == 0 && a.lse != nullptr) a.lse[g.row(seq, ra) * a.heads + head] = m_a + log2f(l_a);
    }
    if (rb < a.n) {
      __nv_bfloat16* orow = o + g.row(seq, rb) * a.ldo + head * DH + 2 * t;
#pragma unroll
      for (int dt = 0; dt < DH / 8; dt++)
        *reinterpret_cast<uint32_t*>(orow + dt * 8) = pack_bf16x2(oacc[dt][2] * inv_b, oacc[dt][3] * inv_b);
      if (t == 0 && a.lse != nullptr) a.lse[g.row(seq, rb) * a.heads + head] = m_b + log2f(l_b);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward, part 1 (query-row parallel): dq_hat.   dlogits = P * (dP - delta), dq = scale * dlogits K
// ------------------------------------------------------------------------------------------------
template <int DH, int WPG, int GROUPS, bool MASK, bool DROP = false>
__global__ void __launch_bounds__(WPG* GROUPS * 32, (WPG * GROUPS <= 8) ? 2 : 1) attn_bwd_dq_kernel(ctclip_attn_args a) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const AttnGeom g{a.n, a.heads, a.seq_inner, a.seq_outer_stride, a.tok_stride};
  const int n_pad = (a.n + 15) & ~15;
  const int tstride = n_pad + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = warp / WPG, wig = warp % WPG;
  const int gq = lane >> 2, t = lane & 3;
  const long long item = (long long)blockIdx.x * GROUPS + group;
  const size_t group_bytes = (size_t)(2 * n_pad * KROW) * 2 + (MASK ? (size_t)n_pad * 4 : 0);
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem_attn + group * group_bytes);
  __nv_bfloat16* sV = sK + n_pad * KROW;
  int* sMask = reinterpret_cast<int*>(sV + n_pad * KROW);
  const bool active = item < (long long)a.num_seqs * a.heads;
  const int head = active ? (int)(item % a.heads) : 0;
  const int seq = active ? (int)(item / a.heads) : 0;
  const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(a.q);
  const __nv_bfloat16* k = reinterpret_cast<const __nv_bfloat16*>(a.k);
  const __nv_bfloat16* v = reinterpret_cast<const __nv_bfloat16*>(a.v);
  const __nv_bfloat16* dO = reinterpret_cast<const __nv_bfloat16*>(a.d_o);
  const __nv_bfloat16* bias = reinterpret_cast<const __nv_bfloat16*>(a.bias);
  __nv_bfloat16* dq = reinterpret_cast<__nv_bfloat16*>(a.dq);
  __nv_bfloat16* ds_out = (a.dbias != nullptr) ? reinterpret_cast<__nv_bfloat16*>(a.ds_scratch) : nullptr;
  if (active) {
    const int tid = wig * 32 + lane;
    load_rows<DH>(sK, k, a.ldk, head, g, seq, n_pad, tid, WPG * 32);
    load_rows<DH>(sV, v, a.ldv, head, g, seq, n_pad, tid, WPG * 32);
    if (MASK)
      for (int i = tid; i < n_pad; i += WPG * 32)
        sMask[i] = (i < a.n && a.key_mask[(long long)seq * a.n + i] != 0) ? 0 : 1;
  }
  group_sync<WPG>(group);
  if (!active) return;
  const float sc2 = a.scale * kLog2e;
  const int row_tiles = n_pad / 16;
  const int kblocks = (n_pad + 63) / 64;
  const uint2* bias_frag = reinterpret_cast<const uint2*>(a.bias_frag);
  for (int rt = wig; rt < row_tiles; rt += WPG) {
    const int r0 = rt * 16;
    uint32_t qa[DH / 16][4], da[DH / 16][4];
    load_a_frags<DH>(qa, q, a.ldq, head, g, seq, r0, lane);
    load_a_frags<DH>(da, dO, a.ldo, head, g, seq, r0, lane);
    const int ra = r0 + gq, rb = r0 + gq + 8;
    // rows beyond n: lse = +inf makes their probabilities exactly 0
    float lse_a = INFINITY, lse_b = INFINITY, del_a = 0.f, del_b = 0.f;
    if (ra < a.n) { lse_a = a.lse[g.row(seq, ra) * a.heads + head]; del_a = a.delta[g.row(seq, ra) * a.heads + head]; }
    if (rb < a.n) { lse_b = a.lse[g.row(seq, rb) * a.heads + head]; del_b = a.delta[g.row(seq, rb) * a.heads + head]; }
    const bool pair_ok = (a.n & 1) == 0;
    const __nv_bfloat16* brow_a = (bias != nullptr && ra < a.n) ? bias + ((long long)head * a.n + ra) * a.n : nullptr;
    const __nv_bfloat16* brow_b = (bias != nullptr && rb < a.n) ? bias + ((long long)head * a.n + rb) * a.n : nullptr;
    float dqa[DH / 8][4];
#pragma unroll
    for (int dt = 0; dt < DH / 8; dt++) dqa[dt][0] = dqa[dt][1] = dqa[dt][2] = dqa[dt][3] = 0.f;
    const bool use_frag = !MASK && bias_frag != nullptr;
    uint4 bcur[2];
    if (use_frag && 32 <= a.n) {
      const uint4* p0 = reinterpret_cast<const uint4*>(bias_frag + (((long long)head * row_tiles + rt) * kblocks * 32 + lane) * 8);
      bcur[0] = __ldg(p0);
      bcur[1] = __ldg(p0 + 1);
    }
    for (int key0 = 0; key0 < n_pad; key0 += 32) {
      float s[4][4], dp[4][4];
      const int rem = n_pad - key0;
      const int ntv = rem >= 32 ? 4 : 2;
      const uint2* bfr = bias_frag ? bias_frag + ((((long long)head * row_tiles + rt) * kblocks + (key0 >> 6)) * 32 + lane) * 8 +
                                         ((key0 >> 5) & 1) * 4
                                   : nullptr;
      const bool full = !MASK && (key0 + 32 <= a.n) && (bias == nullptr || bias_frag != nullptr);   // warp-uniform
      const bool nfull = use_frag && (key0 + 64 <= a.n);
      if (full) {
        qk_block<4, DH, true>(s, qa, sK, key0, lane);
        qk_block<4, DH, true>(dp, da, sV, key0, lane);
        if (use_frag) logits_tile_full_regs<4>(s, sc2, bcur);
        else logits_tile_full<4>(s, sc2, nullptr);
        if (nfull) {   // fragments of the next 32-key block: second half of this 64-key table block, or the next table block
          const uint4* pn = reinterpret_cast<const uint4*>(bfr + (((key0 >> 5) & 1) ? 256 - 4 : 4));
          bcur[0] = __ldg(pn);
          bcur[1] = __ldg(pn + 1);
        }
      } else {
        qk_block<4, DH>(s, qa, sK, key0, lane, ntv);
        qk_block<4, DH>(dp, da, sV, key0, lane, ntv);
        logits_tile<4, MASK>(s, sc2, brow_a, brow_b, key0, t, a.n, key0 + 32 <= a.n, pair_ok, sMask, ntv, bfr);
      }
      if (DROP) {   // dP is the gradient w.r.t. the DROPPED probabilities: d softmax-output = keep/(1-p) * dP
        const AttnDrop dr = make_attn_drop(a);
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
          const int c = key0 + nt * 8 + 2 * t;
          if (c < a.n) {
            const float2 ka = dr.pair(item, ra < a.n ? ra : 0, c), kb = dr.pair(item, rb < a.n ? rb : 0, c);
            dp[nt][0] *= ka.x; dp[nt][1] *= ka.y;
            dp[nt][2] *= kb.x; dp[nt][3] *= kb.y;
          }
        }
      }
#pragma unroll
      for (int nt = 0; nt < 4; nt++) {
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const float p = fast_exp2(s[nt][e] - ((e < 2) ? lse_a : lse_b));           // -inf logits -> 0
          s[nt][e] = p * (dp[nt][e] - ((e < 2) ? del_a : del_b));                    // d logits (the softmax scale is applied to dq below)
        }
      }
      if (ds_out != nullptr) {   // d logits == d bias: spill them (bf16) for the sequence reduction that replaces the third pass
        __nv_bfloat16* da_ = ds_out + ((long long)item * a.n + ra) * a.n + key0 + 2 * t;
        __nv_bfloat16* db_ = ds_out + ((long long)item * a.n + rb) * a.n + key0 + 2 * t;
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
          const int c = key0 + nt * 8 + 2 * t;
          if (c + 1 < a.n) {   // n is even on this path (checked by the launcher): pairs never straddle the row end
            if (ra < a.n) *reinterpret_cast<uint32_t*>(da_ + nt * 8) = pack_bf16x2(s[nt][0], s[nt][1]);
            if (rb < a.n) *reinterpret_cast<uint32_t*>(db_ + nt * 8) = pack_bf16x2(s[nt][2], s[nt][3]);
          }
        }
      }
      if (full) pv_block<4, DH, true>(dqa, s, sK, key0, lane);
      else pv_block<4, DH>(dqa, s, sK, key0, lane, ntv);
    }
    if (ra < a.n) {
      __nv_bfloat16* orow = dq + g.row(seq, ra) * a.ld_dq + head * DH + 2 * t;
#pragma unroll
      for (int dt = 0; dt < DH / 8; dt++)
        *reinterpret_cast<uint32_t*>(orow + dt * 8) = pack_bf16x2(dqa[dt][0] * a.scale, dqa[dt][1] * a.scale);
    }
    if (rb < a.n) {
      __nv_bfloat16* orow = dq + g.row(seq, rb) * a.ld_dq + head * DH + 2 * t;
#pragma unroll
      for (int dt = 0; dt < DH / 8; dt++)
        *reinterpret_cast<uint32_t*>(orow + dt * 8) = pack_bf16x2(dqa[dt][2] * a.scale, dqa[dt][3] * a.scale);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward, part 2 (key-row parallel): dk_hat, dv.  Works on S^T = K Q^T so that P^T / dS^T come out
// of the MMA in the register layout the next MMA needs as its A operand.
// ------------------------------------------------------------------------------------------------
template <int DH, int WPG, int GROUPS, bool MASK, bool DROP = false>
__global__ void __launch_bounds__(WPG* GROUPS * 32) attn_bwd_dkv_kernel(ctclip_attn_args a) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const AttnGeom g{a.n, a.heads, a.seq_inner, a.seq_outer_stride, a.tok_stride};
  const int n_pad = (a.n + 15) & ~15;
  const int tstride = n_pad + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = warp / WPG, wig = warp % WPG;
  const int gq = lane >> 2, t = lane & 3;
  const long long item = (long long)blockIdx.x * GROUPS + group;
  const size_t group_bytes = (size_t)(2 * n_pad * KROW) * 2 + (size_t)n_pad * 8;
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_attn + group * group_bytes);
  __nv_bfloat16* sDO = sQ + n_pad * KROW;
  float* sLse = reinterpret_cast<float*>(sDO + n_pad * KROW);
  float* sDel = sLse + n_pad;
  const bool active = item < (long long)a.num_seqs * a.heads;
  const int head = active ? (int)(item % a.heads) : 0;
  const int seq = active ? (int)(item / a.heads) : 0;
  const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(a.q);
  const __nv_bfloat16* k = reinterpret_cast<const __nv_bfloat16*>(a.k);
  const __nv_bfloat16* v = reinterpret_cast<const __nv_bfloat16*>(a.v);
  const __nv_bfloat16* dO = reinterpret_cast<const __nv_bfloat16*>(a.d_o);
  const __nv_bfloat16* biasT = reinterpret_cast<const __nv_bfloat16*>(a.bias_t);
  __nv_bfloat16* dk = reinterpret_cast<__nv_bfloat16*>(a.dk);
  __nv_bfloat16* dv = reinterpret_cast<__nv_bfloat16*>(a.dv);
  if (active) {
    const int tid = wig * 32 + lane;
    load_rows<DH>(sQ, q, a.ldq, head, g, seq, n_pad, tid, WPG * 32);
    load_rows<DH>(sDO, dO, a.ldo, head, g, seq, n_pad, tid, WPG * 32);
    for (int i = tid; i < n_pad; i += WPG * 32) {
      sLse[i] = i < a.n ? a.lse[g.row(seq, i) * a.heads + head] : 0.f;
      sDel[i] = i < a.n ? a.delta[g.row(seq, i) * a.heads + head] : 0.f;
    }
  }
  group_sync<WPG>(group);
  if (!active) return;
  const float sc2 = a.scale * kLog2e;
  const int key_tiles = n_pad / 16;
  const int qblocks = (n_pad + 63) / 64;
  const uint2* bias_t_frag = reinterpret_cast<const uint2*>(a.bias_t_frag);
  for (int kt_ = wig; kt_ < key_tiles; kt_ += WPG) {
    const int k0 = kt_ * 16;
    uint32_t ka[DH / 16][4], va[DH / 16][4];
    load_a_frags<DH>(ka, k, a.ldk, head, g, seq, k0, lane);
    load_a_frags<DH>(va, v, a.ldv, head, g, seq, k0, lane);
    const int ka_ = k0 + gq, kb_ = k0 + gq + 8;  // key rows owned by this thread
    const bool pair_ok = (a.n & 1) == 0;
    const __nv_bfloat16* brow_a = (biasT != nullptr && ka_ < a.n) ? biasT + ((long long)head * a.n + ka_) * a.n : nullptr;
    const __nv_bfloat16* brow_b = (biasT != nullptr && kb_ < a.n) ? biasT + ((long long)head * a.n + kb_) * a.n : nullptr;
    bool keep_a = ka_ < a.n, keep_b = kb_ < a.n;
    if (MASK) {
      if (keep_a) keep_a = a.key_mask[(long long)seq * a.n + ka_] != 0;
      if (keep_b) keep_b = a.key_mask[(long long)seq * a.n + kb_] != 0;
    }
    float dka[DH / 8][4], dva[DH / 8][4];
#pragma unroll
    for (int dt = 0; dt < DH / 8; dt++) {
      dka[dt][0] = dka[dt][1] = dka[dt][2] = dka[dt][3] = 0.f;
      dva[dt][0] = dva[dt][1] = dva[dt][2] = dva[dt][3] = 0.f;
    }
    const bool use_frag = bias_t_frag != nullptr;
    uint4 bcur[2];
    if (use_frag && 32 <= a.n) {
      const uint4* p0 = reinterpret_cast<const uint4*>(bias_t_frag + (((long long)head * key_tiles + kt_) * qblocks * 32 + lane) * 8);
      bcur[0] = __ldg(p0);
      bcur[1] = __ldg(p0 + 1);
    }
    for (int q0 = 0; q0 < n_pad; q0 += 32) {
      float s[4][4], dp[4][4];
      const int rem = n_pad - q0;
      const int ntv = rem >= 32 ? 4 : 2;
      const bool full = (q0 + 32 <= a.n) && (biasT == nullptr || bias_t_frag != nullptr);   // warp-uniform
      const bool nfull = use_frag && (q0 + 64 <= a.n);
      if (full) {
        qk_block<4, DH, true>(s, ka, sQ, q0, lane);
        qk_block<4, DH, true>(dp, va, sDO, q0, lane);
      } else {
        qk_block<4, DH>(s, ka, sQ, q0, lane, ntv);
        qk_block<4, DH>(dp, va, sDO, q0, lane, ntv);
      }
      // columns of this block are queries: mask columns >= n; masked / out-of-range key rows via keep_a / keep_b
      const uint2* bfr = bias_t_frag ? bias_t_frag + ((((long long)head * key_tiles + kt_) * qblocks + (q0 >> 6)) * 32 + lane) * 8 +
                                           ((q0 >> 5) & 1) * 4
                                     : nullptr;
      if (full) {
        if (use_frag) logits_tile_full_regs<4>(s, sc2, bcur);
        else logits_tile_full<4>(s, sc2, nullptr);
        if (nfull) {
          const uint4* pn = reinterpret_cast<const uint4*>(bfr + (((q0 >> 5) & 1) ? 256 - 4 : 4));
          bcur[0] = __ldg(pn);
          bcur[1] = __ldg(pn + 1);
        }
      } else {
        logits_tile<4, false>(s, sc2, brow_a, brow_b, q0, t, a.n, q0 + 32 <= a.n, pair_ok, nullptr, ntv, bfr);
      }
      float ds[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; nt++) {
        const int qr = q0 + nt * 8 + 2 * t;              // < n_pad whenever nt < ntv (sLse/sDel are n_pad long, zero-filled)
        float2 l2 = make_float2(0.f, 0.f), d2 = make_float2(0.f, 0.f);
        if (nt < ntv) {
          l2 = *reinterpret_cast<const float2*>(sLse + qr);
          d2 = *reinterpret_cast<const float2*>(sDel + qr);
        }
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const bool keep = (e < 2) ? keep_a : keep_b;
          const float p = keep ? fast_exp2(s[nt][e] - ((e & 1) ? l2.y : l2.x)) : 0.f;
          float dmul = 1.f;
          if (DROP) {   // element (query qr + (e&1), key ka_ / kb_) of the probability tensor
            const int qq = qr + (e & 1), kk = (e < 2) ? ka_ : kb_;
            dmul = (qq < a.n && kk < a.n) ? make_attn_drop(a).one(item, qq, kk) : 0.f;
          }
          s[nt][e] = p * dmul;                                                     // (dropped) P^T: dV = (P o M)^T dO
          ds[nt][e] = p * (dp[nt][e] * dmul - ((e & 1) ? d2.y : d2.x));            // dS^T w.r.t. the logits (scale applied to dk below)
        }
      }
      if (full) {
        pv_block<4, DH, true>(dva, s, sDO, q0, lane);
        pv_block<4, DH, true>(dka, ds, sQ, q0, lane);
      } else {
        pv_block<4, DH>(dva, s, sDO, q0, lane, ntv);
        pv_block<4, DH>(dka, ds, sQ, q0, lane, ntv);
      }
    }
    if (ka_ < a.n) {
      __nv_bfloat16* r1 = dk + g.row(seq, ka_) * a.ld_dk + head * DH + 2 * t;
      __nv_bfloat16* r2 = dv + g.row(seq, ka_) * a.ld_dv + head * DH + 2 * t;
#pragma unroll
      for (int dt = 0; dt < DH / 8; dt++) {
        *reinterpret_cast<uint32_t*>(r1 + dt * 8) = pack_bf16x2(dka[dt][0] * a.scale, dka[dt][1] * a.scale);
        *reinterpret_cast<uint32_t*>(r2 + dt * 8) = pack_bf16x2(dva[dt][0], dva[dt][1]);
      }
    }
    if (kb_ < a.n) {
      __nv_bfloat16* r1 = dk + g.row(seq, kb_) * a.ld_dk + head * DH + 2 * t;
      __nv_bfloat16* r2 = dv + g.row(seq, kb_) * a.ld_dv + head * DH + 2 * t;
#pragma unroll
      for (int dt = 0; dt < DH / 8; dt++) {
        *reinterpret_cast<uint32_t*>(r1 + dt * 8) = pack_bf16x2(dka[dt][2] * a.scale, dka[dt][3] * a.scale);
        *reinterpret_cast<uint32_t*>(r2 + dt * 8) = pack_bf16x2(dva[dt][2], dva[dt][3]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward, part 3 (spatial stack only): dbias[h,i,j] += sum_seq dlogits_seq[h,i,j].
// CTA = (head, 64 query rows, 64 keys, chunk of sequences); 4 warps x 16 rows. The K/V tiles of consecutive
// sequences stream through a 2-stage cp.async ring and the per-row operands (q_hat, dO fragments, lse, delta) of the
// next sequence are prefetched into registers while the current one is being processed; partial sums are merged with
// one fp32 red.add per element per chunk.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src, bool valid) {
  const uint32_t d = smem_u32(smem_dst);
  const int bytes = valid ? 16 : 0;   // src-size 0 => 16 bytes of zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(128, 3) attn_bwd_dbias_kernel(ctclip_attn_args a, int seq_chunks) {
  constexpr int DH = 32;
  constexpr int BROW = 72;  // padded bias-tile row (bf16)
  __shared__ __align__(16) __nv_bfloat16 sK[2][64 * KROW];
  __shared__ __align__(16) __nv_bfloat16 sV[2][64 * KROW];
  __shared__ __align__(16) __nv_bfloat16 sB[64 * BROW];
  const AttnGeom g{a.n, a.heads, a.seq_inner, a.seq_outer_stride, a.tok_stride};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, t = lane & 3;
  const int head = blockIdx.z % a.heads;
  const int chunk = blockIdx.z / a.heads;
  const int per = (a.num_seqs + seq_chunks - 1) / seq_chunks;
  const int seq_begin = chunk * per;
  const int seq_end = min(a.num_seqs, seq_begin + per);
  if (seq_begin >= seq_end) return;
  const int rblk = blockIdx.y * 64;
  const int r0 = rblk + warp * 16;
  const int key0 = blockIdx.x * 64;
  const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(a.q);
  const __nv_bfloat16* k = reinterpret_cast<const __nv_bfloat16*>(a.k);
  const __nv_bfloat16* v = reinterpret_cast<const __nv_bfloat16*>(a.v);
  const __nv_bfloat16* dO = reinterpret_cast<const __nv_bfloat16*>(a.d_o);
  const __nv_bfloat16* bias = reinterpret_cast<const __nv_bfloat16*>(a.bias);
  const float sc2 = a.scale * kLog2e;
  const int ra = r0 + gq, rb = r0 + gq + 8;
  // bias tile (64 x 64) of this CTA, zero where out of range
  for (int idx = threadIdx.x; idx < 64 * 64; idx += 128) {
    const int r = idx >> 6, c = idx & 63;
    float bv = 0.f;
    if (bias != nullptr && rblk + r < a.n && key0 + c < a.n)
      bv = __bfloat162float(bias[((long long)head * a.n + rblk + r) * a.n + key0 + c]);
    sB[r * BROW + c] = __float2bfloat16(bv);
  }
  auto issue_tile = [&](int seq, int buf) {
    for (int idx = threadIdx.x; idx < 64 * 4; idx += 128) {
      const int r = idx >> 2, part = idx & 3;
      const bool ok = key0 + r < a.n;
      const long long row = ok ? g.row(seq, key0 + r) : 0;
      cp_async_16(&sK[buf][r * KROW + part * 8], k + row * a.ldk + head * DH + part * 8, ok);
      cp_async_16(&sV[buf][r * KROW + part * 8], v + row * a.ldv + head * DH + part * 8, ok);
    }
    cp_async_commit();
  };
  float acc[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; nt++) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
  uint32_t qa[DH / 16][4], da[DH / 16][4], qn[DH / 16][4], dn[DH / 16][4];
  float lse_a, lse_b, del_a, del_b, nlse_a, nlse_b, ndel_a, ndel_b;
  auto load_rows_ops = [&](int seq, uint32_t (&fq)[DH / 16][4], uint32_t (&fd)[DH / 16][4], float& la, float& lb, float& dla,
                           float& dlb) {
    load_a_frags<DH>(fq, q, a.ldq, head, g, seq, r0, lane);
    load_a_frags<DH>(fd, dO, a.ldo, head, g, seq, r0, lane);
    la = lb = INFINITY;   // rows beyond n contribute probability 0
    dla = dlb = 0.f;
    if (ra < a.n) { la = a.lse[g.row(seq, ra) * a.heads + head]; dla = a.delta[g.row(seq, ra) * a.heads + head]; }
    if (rb < a.n) { lb = a.lse[g.row(seq, rb) * a.heads + head]; dlb = a.delta[g.row(seq, rb) * a.heads + head]; }
  };
  issue_tile(seq_begin, 0);
  load_rows_ops(seq_begin, qa, da, lse_a, lse_b, del_a, del_b);
  for (int seq = seq_begin; seq < seq_end; seq++) {
    const int buf = (seq - seq_begin) & 1;
    cp_async_wait<0>();
    __syncthreads();                       // tile `buf` landed for everyone; everyone finished with tile `buf^1`
    if (seq + 1 < seq_end) {
      issue_tile(seq + 1, buf ^ 1);
      load_rows_ops(seq + 1, qn, dn, nlse_a, nlse_b, ndel_a, ndel_b);
    }
    float s[8][4], dp[8][4];
    qk_block<8, DH>(s, qa, sK[buf], 0, lane);
    qk_block<8, DH>(dp, da, sV[buf], 0, lane);
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
      const int c = nt * 8 + 2 * t;
      const float2 b_a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(&sB[(warp * 16 + gq) * BROW + c]));
      const float2 b_b = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(&sB[(warp * 16 + gq + 8) * BROW + c]));
      const bool k0ok = key0 + c < a.n, k1ok = key0 + c + 1 < a.n;
      const float p0 = k0ok ? fast_exp2(fmaf(s[nt][0], sc2, b_a.x * kLog2e) - lse_a) : 0.f;
      const float p1 = k1ok ? fast_exp2(fmaf(s[nt][1], sc2, b_a.y * kLog2e) - lse_a) : 0.f;
      const float p2 = k0ok ? fast_exp2(fmaf(s[nt][2], sc2, b_b.x * kLog2e) - lse_b) : 0.f;
      const float p3 = k1ok ? fast_exp2(fmaf(s[nt][3], sc2, b_b.y * kLog2e) - lse_b) : 0.f;
      acc[nt][0] += p0 * (dp[nt][0] - del_a);
      acc[nt][1] += p1 * (dp[nt][1] - del_a);
      acc[nt][2] += p2 * (dp[nt][2] - del_b);
      acc[nt][3] += p3 * (dp[nt][3] - del_b);
    }
    if (seq + 1 < seq_end) {
#pragma unroll
      for (int kt = 0; kt < DH / 16; kt++)
#pragma unroll
        for (int e = 0; e < 4; e++) { qa[kt][e] = qn[kt][e]; da[kt][e] = dn[kt][e]; }
      lse_a = nlse_a; lse_b = nlse_b; del_a = ndel_a; del_b = ndel_b;
    }
  }
#pragma unroll
  for (int nt = 0; nt < 8; nt++) {
    const int key = key0 + nt * 8 + 2 * t;
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const int kk = key + (e & 1), rr = (e < 2) ? ra : rb;
      if (rr < a.n && kk < a.n) atomicAdd(&a.dbias[((long long)head * a.n + rr) * a.n + kk], acc[nt][e]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Short-sequence kernels (n <= 32, dim_head 32, no bias, no key mask): the temporal stack, 24 tokens per sequence,
// 36 864 (sequence, head) items per layer at configs[1]. These are pure HBM-streaming problems (4.6 KB in / 1.5 KB out
// per item forward, 6.3 KB in / 4.5 KB out backward; < 1 us of math per item), so the kernels are PERSISTENT: one warp
// owns an item at a time, its Q/K/V(/dO) tiles arrive through a private two-stage cp.async ring in shared memory so
// that the loads of the next item are in flight while the current one is computed, and the backward produces dQ, dK
// and dV in ONE pass (the general path's three kernels re-read everything and spent one exposed DRAM round trip per
// item with nothing else in flight: 0.19 / 0.57 ms per layer where the traffic needs 0.04 / 0.07 ms).
// Tile layout: [32 rows][KROW = 40] bf16 (rows >= n stay zero: zero-filled once, cp.async never touches them).
// ------------------------------------------------------------------------------------------------
constexpr int SH_KROW = 40;
constexpr int SH_TILE = 32 * SH_KROW;   // elements
constexpr int SH_FWD_WARPS = 12;
constexpr int SH_BWD_WARPS = 10;

__device__ __forceinline__ void sh_cp16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
// rows [0, n) x 64 bytes of one head slice -> tile (4 x 16-byte pieces per row)
__device__ __forceinline__ void sh_load_tile(__nv_bfloat16* tile, const __nv_bfloat16* src, long long ld, int head,
                                             const AttnGeom& g, int seq, int lane) {
  for (int idx = lane; idx < g.n * 4; idx += 32) {
    const int r = idx >> 2, part = idx & 3;
    sh_cp16(tile + r * SH_KROW + part * 8, src + g.row(seq, r) * ld + head * 32 + part * 8);
  }
}
// A-operand fragments (16 rows x 32) of a row-major smem tile via ldmatrix.x4
__device__ __forceinline__ void sh_a_frags(uint32_t (&a)[2][4], const __nv_bfloat16* tile, int r0, int lane) {
  const __nv_bfloat16* base = tile + (r0 + (lane & 7) + ((lane >> 3) & 1) * 8) * SH_KROW + (lane >> 4) * 8;
  ldsm_x4(a[0], base);
  ldsm_x4(a[1], base + 16);
}

__global__ void __launch_bounds__(SH_FWD_WARPS * 32, 1) attn_short_fwd_kernel(ctclip_attn_args a) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  constexpr int DH = 32;
  const AttnGeom g{a.n, a.heads, a.seq_inner, a.seq_outer_stride, a.tok_stride};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, t = lane & 3;
  __nv_bfloat16* ring = reinterpret_cast<__nv_bfloat16*>(smem_attn) + (size_t)warp * (2 * 3 * SH_TILE);   // [2][q,k,v]
  for (int i = lane; i < 2 * 3 * SH_TILE / 8; i += 32) reinterpret_cast<uint4*>(ring)[i] = make_uint4(0, 0, 0, 0);
  __syncwarp();
  const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(a.q);
  const __nv_bfloat16* k = reinterpret_cast<const __nv_bfloat16*>(a.k);
  const __nv_bfloat16* v = reinterpret_cast<const __nv_bfloat16*>(a.v);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.o);
  const long long total = (long long)a.num_seqs * a.heads;
  const long long stride = (long long)gridDim.x * SH_FWD_WARPS;
  const float sc2 = a.scale * kLog2e;
  long long item = (long long)blockIdx.x * SH_FWD_WARPS + warp;
  int stage = 0;
  if (item < total) {
    const int head = (int)(item % a.heads), seq = (int)(item / a.heads);
    sh_load_tile(ring, q, a.ldq, head, g, seq, lane);
    sh_load_tile(ring + SH_TILE, k, a.ldk, head, g, seq, lane);
    sh_load_tile(ring + 2 * SH_TILE, v, a.ldv, head, g, seq, lane);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (; item < total; item += stride, stage ^= 1) {
    const long long nxt = item + stride;
    if (nxt < total) {
      __nv_bfloat16* nb = ring + (stage ^ 1) * 3 * SH_TILE;
      const int head = (int)(nxt % a.heads), seq = (int)(nxt / a.heads);
      sh_load_tile(nb, q, a.ldq, head, g, seq, lane);
      sh_load_tile(nb + SH_TILE, k, a.ldk, head, g, seq, lane);
      sh_load_tile(nb + 2 * SH_TILE, v, a.ldv, head, g, seq, lane);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncwarp();
    const int head = (int)(item % a.heads), seq = (int)(item / a.heads);
    const __nv_bfloat16* sQ = ring + stage * 3 * SH_TILE;
    const __nv_bfloat16* sK = sQ + SH_TILE;
    const __nv_bfloat16* sV = sK + SH_TILE;
#pragma unroll 1
    for (int r0 = 0; r0 < a.n; r0 += 16) {
      uint32_t qa[2][4];
      sh_a_frags(qa, sQ, r0, lane);
      float sc[4][4];
      qk_block<4, DH, true>(sc, qa, sK, 0, lane);
      float m_a = -INFINITY, m_b = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 4; nt++) {
        const int c = nt * 8 + 2 * t;
        sc[nt][0] = (c < a.n) ? sc[nt][0] * sc2 : -INFINITY;
        sc[nt][1] = (c + 1 < a.n) ? sc[nt][1] * sc2 : -INFINITY;
        sc[nt][2] = (c < a.n) ? sc[nt][2] * sc2 : -INFINITY;
        sc[nt][3] = (c + 1 < a.n) ? sc[nt][3] * sc2 : -INFINITY;
        m_a = fmaxf(m_a, fmaxf(sc[nt][0], sc[nt][1]));
        m_b = fmaxf(m_b, fmaxf(sc[nt][2], sc[nt][3]));
      }
      m_a = quad_max(m_a);
      m_b = quad_max(m_b);
      float l_a = 0.f, l_b = 0.f;
#pragma unroll
      for (int nt = 0; nt < 4; nt++) {
        sc[nt][0] = fast_exp2(sc[nt][0] - m_a);
        sc[nt][1] = fast_exp2(sc[nt][1] - m_a);
        sc[nt][2] = fast_exp2(sc[nt][2] - m_b);
        sc[nt][3] = fast_exp2(sc[nt][3] - m_b);
        l_a += sc[nt][0] + sc[nt][1];
        l_b += sc[nt][2] + sc[nt][3];
      }
      l_a = quad_sum(l_a);
      l_b = quad_sum(l_b);
      float oacc[DH / 8][4];
#pragma unroll
      for (int dt = 0; dt < DH / 8; dt++) oacc[dt][0] = oacc[dt][1] = oacc[dt][2] = oacc[dt][3] = 0.f;
      pv_block<4, DH, true>(oacc, sc, sV, 0, lane);
      const float inv_a = 1.f / l_a, inv_b = 1.f / l_b;
      const int ra = r0 + gq, rb = r0 + gq + 8;
      if (ra < a.n) {
        const long long row = g.row(seq, ra);
        __nv_bfloat16* orow = o + row * a.ldo + head * DH + 2 * t;
#pragma unroll
        for (int dt = 0; dt < DH / 8; dt++)
          *reinterpret_cast<uint32_t*>(orow + dt * 8) = pack_bf16x2(oacc[dt][0] * inv_a, oacc[dt][1] * inv_a);
        if (t == 0 && a.lse != nullptr) a.lse[row * a.heads + head] = m_a + log2f(l_a);
      }
      if (rb < a.n) {
        const long long row = g.row(seq, rb);
        __nv_bfloat16* orow = o + row * a.ldo + head * DH + 2 * t;
#pragma unroll
        for (int dt = 0; dt < DH / 8; dt++)
          *reinterpret_cast<uint32_t*>(orow + dt * 8) = pack_bf16x2(oacc[dt][2] * inv_b, oacc[dt][3] * inv_b);
        if (t == 0 && a.lse != nullptr) a.lse[row * a.heads + head] = m_b + log2f(l_b);
      }
    }
    __syncwarp();   // every lane is done with this stage before the next iteration refills it
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// backward: dq_hat, dk_hat, dv of one (sequence, head) item per warp iteration, delta computed by attn_delta_kernel
__global__ void __launch_bounds__(SH_BWD_WARPS * 32, 1) attn_short_bwd_kernel(ctclip_attn_args a) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  constexpr int DH = 32;
  const AttnGeom g{a.n, a.heads, a.seq_inner, a.seq_outer_stride, a.tok_stride};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, t = lane & 3;
  constexpr int WARP_ELEMS = 2 * 4 * SH_TILE + 128;   // [2][q,k,v,dO] tiles + 64 floats (lse, delta) as 128 bf16 slots
  __nv_bfloat16* ring = reinterpret_cast<__nv_bfloat16*>(smem_attn) + (size_t)warp * WARP_ELEMS;
  float* sLse = reinterpret_cast<float*>(ring + 2 * 4 * SH_TILE);   // [32]
  float* sDel = sLse + 32;                                          // [32]
  for (int i = lane; i < 2 * 4 * SH_TILE / 8; i += 32) reinterpret_cast<uint4*>(ring)[i] = make_uint4(0, 0, 0, 0);
  __syncwarp();
  const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(a.q);
  const __nv_bfloat16* k = reinterpret_cast<const __nv_bfloat16*>(a.k);
  const __nv_bfloat16* v = reinterpret_cast<const __nv_bfloat16*>(a.v);
  const __nv_bfloat16* dO = reinterpret_cast<const __nv_bfloat16*>(a.d_o);
  __nv_bfloat16* dq = reinterpret_cast<__nv_bfloat16*>(a.dq);
  __nv_bfloat16* dk = reinterpret_cast<__nv_bfloat16*>(a.dk);
  __nv_bfloat16* dv = reinterpret_cast<__nv_bfloat16*>(a.dv);
  const long long total = (long long)a.num_seqs * a.heads;
  const long long stride = (long long)gridDim.x * SH_BWD_WARPS;
  const float sc2 = a.scale * kLog2e;
  long long item = (long long)blockIdx.x * SH_BWD_WARPS + warp;
  int stage = 0;
  float lse_n = INFINITY, del_n = 0.f;   // this lane's row of the NEXT item (row = lane)
  auto issue = [&](long long it_, int st_) {
    __nv_bfloat16* nb = ring + st_ * 4 * SH_TILE;
    const int head = (int)(it_ % a.heads), seq = (int)(it_ / a.heads);
    sh_load_tile(nb, q, a.ldq, head, g, seq, lane);
    sh_load_tile(nb + SH_TILE, k, a.ldk, head, g, seq, lane);
    sh_load_tile(nb + 2 * SH_TILE, v, a.ldv, head, g, seq, lane);
    sh_load_tile(nb + 3 * SH_TILE, dO, a.ldo, head, g, seq, lane);
    lse_n = INFINITY;   // rows >= n: probability exactly 0
    del_n = 0.f;
    if (lane < a.n) {
      const long long row = g.row(seq, lane);
      lse_n = __ldg(a.lse + row * a.heads + head);
      del_n = __ldg(a.delta + row * a.heads + head);
    }
  };
  if (item < total) issue(item, 0);
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (; item < total; item += stride, stage ^= 1) {
    sLse[lane] = lse_n;
    sDel[lane] = del_n;
    const long long nxt = item + stride;
    if (nxt < total) issue(nxt, stage ^ 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncwarp();
    const int head = (int)(item % a.heads), seq = (int)(item / a.heads);
    const __nv_bfloat16* sQ = ring + stage * 4 * SH_TILE;
    const __nv_bfloat16* sK = sQ + SH_TILE;
    const __nv_bfloat16* sV = sK + SH_TILE;
    const __nv_bfloat16* sDO = sV + SH_TILE;
    // ---- query-row parallel: dq = scale * (P o (dP - delta)) K
#pragma unroll 1
    for (int r0 = 0; r0 < a.n; r0 += 16) {
      uint32_t qa[2][4], da[2][4];
      sh_a_frags(qa, sQ, r0, lane);
      sh_a_frags(da, sDO, r0, lane);
      float sc[4][4], dp[4][4];
      qk_block<4, DH, true>(sc, qa, sK, 0, lane);
      qk_block<4, DH, true>(dp, da, sV, 0, lane);
      const int ra = r0 + gq, rb = r0 + gq + 8;
      const float lse_a = sLse[ra], lse_b = sLse[rb], del_a = sDel[ra], del_b = sDel[rb];
#pragma unroll
      for (int nt = 0; nt < 4; nt++) {
        const int c = nt * 8 + 2 * t;
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const bool ok = (c + (e & 1)) < a.n;
          const float p = ok ? fast_exp2(fmaf(sc[nt][e], sc2, -((e < 2) ? lse_a : lse_b))) : 0.f;
          sc[nt][e] = p * (dp[nt][e] - ((e < 2) ? del_a : del_b));
        }
      }
      float dqa[DH / 8][4];
#pragma unroll
      for (int dt = 0; dt < DH / 8; dt++) dqa[dt][0] = dqa[dt][1] = dqa[dt][2] = dqa[dt][3] = 0.f;
      pv_block<4, DH, true>(dqa, sc, sK, 0, lane);
      if (ra < a.n) {
        __nv_bfloat16* orow = dq + g.row(seq, ra) * a.ld_dq + head * DH + 2 * t;
#pragma unroll
        for (int dt = 0; dt < DH / 8; dt++)
          *reinterpret_cast<uint32_t*>(orow + dt * 8) = pack_bf16x2(dqa[dt][0] * a.scale, dqa[dt][1] * a.scale);
      }
      if (rb < a.n) {
        __nv_bfloat16* orow = dq + g.row(seq, rb) * a.ld_dq + head * DH + 2 * t;
#pragma unroll
        for (int dt = 0; dt < DH / 8; dt++)
          *reinterpret_cast<uint32_t*>(orow + dt * 8) = pack_bf16x2(dqa[dt][2] * a.scale, dqa[dt][3] * a.scale);
      }
    }
    // ---- key-row parallel on S^T = K Q^T: dv = P^T dO, dk = scale * dS^T Q
#pragma unroll 1
    for (int k0 = 0; k0 < a.n; k0 += 16) {
      uint32_t ka[2][4], va[2][4];
      sh_a_frags(ka, sK, k0, lane);
      sh_a_frags(va, sV, k0, lane);
      float st[4][4], dpt[4][4], ds[4][4];
      qk_block<4, DH, true>(st, ka, sQ, 0, lane);
      qk_block<4, DH, true>(dpt, va, sDO, 0, lane);
      const int ka_ = k0 + gq, kb_ = k0 + gq + 8;
      const bool keep_a = ka_ < a.n, keep_b = kb_ < a.n;
#pragma unroll
      for (int nt = 0; nt < 4; nt++) {
        const int qr = nt * 8 + 2 * t;   // query index of this column pair; sLse = +inf beyond n
        const float2 l2 = *reinterpret_cast<const float2*>(sLse + qr);
        const float2 d2 = *reinterpret_cast<const float2*>(sDel + qr);
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const bool keep = (e < 2) ? keep_a : keep_b;
          const float p = keep ? fast_exp2(fmaf(st[nt][e], sc2, -((e & 1) ? l2.y : l2.x))) : 0.f;
          st[nt][e] = p;
          ds[nt][e] = p * (dpt[nt][e] - ((e & 1) ? d2.y : d2.x));
        }
      }
      float dka[DH / 8][4], dva[DH / 8][4];
#pragma unroll
      for (int dt = 0; dt < DH / 8; dt++) {
        dka[dt][0] = dka[dt][1] = dka[dt][2] = dka[dt][3] = 0.f;
        dva[dt][0] = dva[dt][1] = dva[dt][2] = dva[dt][3] = 0.f;
      }
      pv_block<4, DH, true>(dva, st, sDO, 0, lane);
      pv_block<4, DH, true>(dka, ds, sQ, 0, lane);
      if (keep_a) {
        const long long row = g.row(seq, ka_);
        __nv_bfloat16* r1 = dk + row * a.ld_dk + head * DH + 2 * t;
        __nv_bfloat16* r2 = dv + row * a.ld_dv + head * DH + 2 * t;
#pragma unroll
        for (int dt = 0; dt < DH / 8; dt++) {
          *reinterpret_cast<uint32_t*>(r1 + dt * 8) = pack_bf16x2(dka[dt][0] * a.scale, dka[dt][1] * a.scale);
          *reinterpret_cast<uint32_t*>(r2 + dt * 8) = pack_bf16x2(dva[dt][0], dva[dt][1]);
        }
      }
      if (keep_b) {
        const long long row = g.row(seq, kb_);
        __nv_bfloat16* r1 = dk + row * a.ld_dk + head * DH + 2 * t;
        __nv_bfloat16* r2 = dv + row * a.ld_dv + head * DH + 2 * t;
#pragma unroll
        for (int dt = 0; dt < DH / 8; dt++) {
          *reinterpret_cast<uint32_t*>(r1 + dt * 8) = pack_bf16x2(dka[dt][2] * a.scale, dka[dt][3] * a.scale);
          *reinterpret_cast<uint32_t*>(r2 + dt * 8) = pack_bf16x2(dva[dt][2], dva[dt][3]);
        }
      }
    }
    __syncwarp();   // stage and sLse / sDel are free for the next iteration
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// dbias[h, i, j] += sum over sequences of the d logits the dQ kernel spilled (bf16 [num_seqs*heads][n*n]); 4 elements/thread
__global__ void __launch_bounds__(256) attn_dbias_reduce_kernel(const __nv_bfloat16* __restrict__ ds, float* __restrict__ dbias,
                                                                int num_seqs, int heads, long long nn) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long quads = nn / 4;
  if (q >= (long long)heads * quads) return;
  const int h = (int)(q / quads);
  const long long e = (q - (long long)h * quads) * 4;
  const __nv_bfloat16* p = ds + (long long)h * nn + e;
  const long long stride = (long long)heads * nn;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
  for (int sq = 0; sq < num_seqs; sq++) {
    const uint2 u = __ldcs(reinterpret_cast<const uint2*>(p + (long long)sq * stride));
    const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y);
    acc.x += a0.x; acc.y += a0.y; acc.z += a1.x; acc.w += a1.y;
  }
  float4* d = reinterpret_cast<float4*>(dbias + (long long)h * nn + e);
  float4 o = *d;
  o.x += acc.x; o.y += acc.y; o.z += acc.z; o.w += acc.w;
  *d = o;
}

// delta[row, head] = sum_d dO[row, head, d] * O[row, head, d]
template <int DH>
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o,
                                  long long ldo, float* __restrict__ delta, long long rows, int heads) {
  constexpr int PARTS = DH / 8;
  const long long idx = (long long)blockIdx.x * (blockDim.x / PARTS) + (threadIdx.x / PARTS);  // (row, head)
  const int part = threadIdx.x % PARTS;
  float s = 0.f;
  if (idx < rows * heads) {
    const long long row = idx / heads;
    const int head = (int)(idx % heads);
    const uint4 uo = *reinterpret_cast<const uint4*>(o + row * ldo + head * DH + part * 8);
    const uint4 ud = *reinterpret_cast<const uint4*>(d_o + row * ldo + head * DH + part * 8);
    const uint32_t* po = reinterpret_cast<const uint32_t*>(&uo);
    const uint32_t* pd = reinterpret_cast<const uint32_t*>(&ud);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const float2 x = unpack_bf16x2(po[i]), y = unpack_bf16x2(pd[i]);
      s += x.x * y.x + x.y * y.y;
    }
  }
  s = quad_sum(s);
  if (PARTS == 8) s += __shfl_xor_sync(0xffffffffu, s, 4);
  if (part == 0 && idx < rows * heads) delta[idx] = s;
}

// Backward of  x_hat = x / max(||x||, eps) * scale  per (row, head):  given d(x_hat) and raw x:
//   u = scale * dxh;  dx = (u - xn * (xn . u)) / ||x||  with xn = x/||x||;   dscale[d] += dxh[d] * xn[d]
// HEADS > 0: compile-time head count (8 on the CTViT path, 12 would be BERT-like); 0: run-time. The item index is 32-bit
// (the 64-bit division + modulo per 48-byte item used to dominate the instruction count of this kernel: 44 % of HBM).
template <int HEADS>
__global__ void __launch_bounds__(256) l2norm_bwd_kernel(const __nv_bfloat16* __restrict__ dxh, long long ld_dxh,
                                                        const __nv_bfloat16* __restrict__ xraw, long long ld_x,
                                                        const float* __restrict__ scale, __nv_bfloat16* __restrict__ dx,
                                                        long long ld_dx, float* __restrict__ dscale, unsigned n_items,
                                                        int heads_rt) {
  constexpr int DH = 32;
  const unsigned heads = HEADS > 0 ? (unsigned)HEADS : (unsigned)heads_rt;
  __shared__ float sds[DH];
  if (threadIdx.x < DH) sds[threadIdx.x] = 0.f;
  __syncthreads();
  const int part = threadIdx.x & 3;
  const unsigned qmask = 0xFu << (threadIdx.x & 28);   // the four lanes of this item (the tail of the item loop diverges per quad)
  float sc[8], dsc[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { sc[i] = scale[part * 8 + i]; dsc[i] = 0.f; }
  const unsigned stride = gridDim.x * (blockDim.x >> 2);
  for (unsigned idx = blockIdx.x * (blockDim.x >> 2) + (threadIdx.x >> 2); idx < n_items; idx += stride) {
    const unsigned row = idx / heads;
    const unsigned head = idx - row * heads;
    const long long col = (long long)head * DH + part * 8;
    const uint4 ug = *reinterpret_cast<const uint4*>(dxh + (long long)row * ld_dxh + col);
    const uint4 ux = *reinterpret_cast<const uint4*>(xraw + (long long)row * ld_x + col);
    const uint32_t* pg = reinterpret_cast<const uint32_t*>(&ug);
    const uint32_t* px = reinterpret_cast<const uint32_t*>(&ux);
    float gg[8], xx[8];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const float2 a = unpack_bf16x2(pg[i]), b = unpack_bf16x2(px[i]);
      gg[2 * i] = a.x; gg[2 * i + 1] = a.y;
      xx[2 * i] = b.x; xx[2 * i + 1] = b.y;
      ss += b.x * b.x + b.y * b.y;
    }
    ss += __shfl_xor_sync(qmask, ss, 1);
    ss += __shfl_xor_sync(qmask, ss, 2);
    const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      xx[i] *= inv;                 // xn
      dsc[i] += gg[i] * xx[i];
      gg[i] *= sc[i];               // u
      dot += gg[i] * xx[i];
    }
    dot += __shfl_xor_sync(qmask, dot, 1);
    dot += __shfl_xor_sync(qmask, dot, 2);
    uint4 out;
    uint32_t* po = reinterpret_cast<uint32_t*>(&out);
#pragma unroll
    for (int i = 0; i < 4; i++)
      po[i] = pack_bf16x2((gg[2 * i] - xx[2 * i] * dot) * inv, (gg[2 * i + 1] - xx[2 * i + 1] * dot) * inv);
    *reinterpret_cast<uint4*>(dx + (long long)row * ld_dx + col) = out;
  }
#pragma unroll
  for (int i = 0; i < 8; i++) atomicAdd(&sds[part * 8 + i], dsc[i]);
  __syncthreads();
  if (threadIdx.x < DH) atomicAdd(dscale + threadIdx.x, sds[threadIdx.x]);
}

}  // namespace ctb

using namespace ctb;

// tcgen05 / TMEM path (attention_tc.cu), selected by args->cpb_table
int ctb_attn_fwd_tc(const ctclip_attn_args* a, cudaStream_t stream);
int ctb_attn_bwd_tc(const ctclip_attn_args* a, cudaStream_t stream);

static int attn_check(const ctclip_attn_args* a, const char* who) {
  CTB_CHECK_ARG(a != nullptr, "%s: null args", who);
  CTB_CHECK_ARG(a->dim_head == 32 || a->dim_head == 64, "%s: dim_head must be 32 or 64 (got %d)", who, a->dim_head);
  CTB_CHECK_ARG(a->n > 0 && a->heads > 0 && a->num_seqs > 0 && a->seq_inner > 0, "%s: bad geometry", who);
  CTB_CHECK_ARG(a->q && a->k && a->v, "%s: null q/k/v", who);
  CTB_CHECK_ARG(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0, "%s: q/k/v rows must be 16B aligned", who);
  CTB_CHECK_ARG(a->key_mask == nullptr || a->bias == nullptr, "%s: key_mask and bias are not combined on this path", who);
  CTB_CHECK_ARG(a->dropout_p >= 0.f && a->dropout_p < 1.f, "%s: dropout_p must be in [0, 1)", who);
  return CTCLIP_OK;
}

template <typename Kern>
static int launch_grouped(Kern kern, const ctclip_attn_args* a, size_t group_bytes, int wpg, int groups,
                          cudaStream_t stream) {
  const size_t smem = group_bytes * groups;
  CTB_CHECK_ARG(smem <= 227 * 1024, "attention: sequence of %d tokens needs %zu B of shared memory (> 227 KB)", a->n, smem);
  CTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long items = (long long)a->num_seqs * a->heads;
  const int grid = (int)((items + groups - 1) / groups);
  kern<<<grid, wpg * groups * 32, smem, stream>>>(*a);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}

// which: 0 = forward, 1 = backward dq, 2 = backward dk/dv.  Group shapes: tiny sequences pack 8 (seq, head) pairs per
// CTA with one warp each, mid-size (BERT, n <= 256) 2 pairs x 4 warps, long (spatial, n = 576/1024) one pair per CTA.
template <int DH, bool MASK>
static int attn_dispatch(int which, const ctclip_attn_args* a, cudaStream_t stream) {
  const int n_pad = (a->n + 15) & ~15;
  if (a->dropout_p > 0.f) {   // attention-probability dropout (BERT text tower: dim_head 64, even sequence lengths)
    if (DH == 64 && a->n % 2 == 0) {
      const size_t mk2 = MASK ? (size_t)n_pad * 4 : 0;
      const size_t g_f = (size_t)(2 * n_pad * (DH + 8)) * 2 + mk2, g_kv = (size_t)(2 * n_pad * (DH + 8)) * 2 + (size_t)n_pad * 8;
      if (a->n <= 64) {
        if (which == 0) return launch_grouped(attn_fwd_kernel<DH, 1, 8, MASK, true>, a, g_f, 1, 8, stream);
        if (which == 1) return launch_grouped(attn_bwd_dq_kernel<DH, 1, 8, MASK, true>, a, g_f, 1, 8, stream);
        return launch_grouped(attn_bwd_dkv_kernel<DH, 1, 8, MASK, true>, a, g_kv, 1, 8, stream);
      }
      if (a->n <= 256) {
        if (which == 0) return launch_grouped(attn_fwd_kernel<DH, 4, 2, MASK, true>, a, g_f, 4, 2, stream);
        if (which == 1) return launch_grouped(attn_bwd_dq_kernel<DH, 4, 2, MASK, true>, a, g_f, 4, 2, stream);
        return launch_grouped(attn_bwd_dkv_kernel<DH, 4, 2, MASK, true>, a, g_kv, 4, 2, stream);
      }
      if (which == 0) return launch_grouped(attn_fwd_kernel<DH, 8, 1, MASK, true>, a, g_f, 8, 1, stream);
      if (which == 1) return launch_grouped(attn_bwd_dq_kernel<DH, 8, 1, MASK, true>, a, g_f, 8, 1, stream);
      return launch_grouped(attn_bwd_dkv_kernel<DH, 12, 1, MASK, true>, a, g_kv, 12, 1, stream);
    }
    set_error("attention: dropout_p > 0 is implemented for dim_head 64 and even n (got dh %d, n %d)", DH, a->n);
    return CTCLIP_ERR_UNSUPPORTED;
  }
  const size_t mk = MASK ? (size_t)n_pad * 4 : 0;
  const size_t gb_fwd = (size_t)(2 * n_pad * (DH + 8)) * 2 + mk;
  const size_t gb_dq = (size_t)(2 * n_pad * (DH + 8)) * 2 + mk;
  const size_t gb_dkv = (size_t)(2 * n_pad * (DH + 8)) * 2 + (size_t)n_pad * 8;
  if (a->n <= 64) {
    if (which == 0) return launch_grouped(attn_fwd_kernel<DH, 1, 8, MASK>, a, gb_fwd, 1, 8, stream);
    if (which == 1) return launch_grouped(attn_bwd_dq_kernel<DH, 1, 8, MASK>, a, gb_dq, 1, 8, stream);
    return launch_grouped(attn_bwd_dkv_kernel<DH, 1, 8, MASK>, a, gb_dkv, 1, 8, stream);
  }
  if (a->n <= 256) {
    if (which == 0) return launch_grouped(attn_fwd_kernel<DH, 4, 2, MASK>, a, gb_fwd, 4, 2, stream);
    if (which == 1) return launch_grouped(attn_bwd_dq_kernel<DH, 4, 2, MASK>, a, gb_dq, 4, 2, stream);
    return launch_grouped(attn_bwd_dkv_kernel<DH, 4, 2, MASK>, a, gb_dkv, 4, 2, stream);
  }
  if (which == 0) return launch_grouped(attn_fwd_kernel<DH, 8, 1, MASK>, a, gb_fwd, 8, 1, stream);   // 128 regs, 2 CTAs / SM
  if (which == 1) return launch_grouped(attn_bwd_dq_kernel<DH, 8, 1, MASK>, a, gb_dq, 8, 1, stream);   // 2 CTAs / SM
  return launch_grouped(attn_bwd_dkv_kernel<DH, 12, 1, MASK>, a, gb_dkv, 12, 1, stream);
}

static int attn_route(int which, const ctclip_attn_args* a, cudaStream_t stream) {
  const bool m = a->key_mask != nullptr;
  if (a->dim_head == 32) return m ? attn_dispatch<32, true>(which, a, stream) : attn_dispatch<32, false>(which, a, stream);
  return m ? attn_dispatch<64, true>(which, a, stream) : attn_dispatch<64, false>(which, a, stream);
}

// n <= 32, dim_head 32, no bias / mask: persistent streaming kernels (the temporal stack)
static bool attn_is_short(const ctclip_attn_args* a) {
  static const int off = getenv("CTCLIP_ATTN_NO_SHORT") ? atoi(getenv("CTCLIP_ATTN_NO_SHORT")) : 0;   // debug knob
  return !off && a->n <= 32 && a->dim_head == 32 && a->bias == nullptr && a->bias_frag == nullptr && a->key_mask == nullptr;
}
static int attn_short_grid(const ctclip_attn_args* a, int warps) {
  const long long items = (long long)a->num_seqs * a->heads;
  const long long ctas = (items + warps - 1) / warps;
  return (int)(ctas < num_sms() ? ctas : num_sms());
}

extern "C" int ctclip_attn_fwd(const ctclip_attn_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = attn_check(a, "attn_fwd")) return rc;
  CTB_CHECK_ARG(a->o != nullptr && a->ldo % 8 == 0, "attn_fwd: bad o");
  if (a->cpb_table != nullptr) return ctb_attn_fwd_tc(a, stream);
  if (attn_is_short(a)) {
    const size_t smem = (size_t)SH_FWD_WARPS * 2 * 3 * SH_TILE * 2;
    CTB_CUDA(cudaFuncSetAttribute(attn_short_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_short_fwd_kernel<<<attn_short_grid(a, SH_FWD_WARPS), SH_FWD_WARPS * 32, smem, stream>>>(*a);
    CTB_LAUNCH_CHECK();
    return CTCLIP_OK;
  }
  return attn_route(0, a, stream);
}

extern "C" int ctclip_attn_bwd(const ctclip_attn_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = attn_check(a, "attn_bwd")) return rc;
  CTB_CHECK_ARG(a->o && a->d_o && a->lse && a->delta && a->dq && a->dk && a->dv, "attn_bwd: null pointer");
  CTB_CHECK_ARG(a->bias == nullptr || a->bias_t != nullptr || a->bias_t_frag != nullptr,
                "attn_bwd: bias needs its transposed copy (bias_t or bias_t_frag)");
  CTB_CHECK_ARG((a->bias_frag == nullptr) == (a->bias_t_frag == nullptr), "attn_bwd: bias_frag and bias_t_frag come together");
  CTB_CHECK_ARG(a->dbias == nullptr || a->dim_head == 32, "attn_bwd: dbias is implemented for dim_head 32 only");
  const long long rows = a->total_rows;
  CTB_CHECK_ARG(rows > 0, "attn_bwd: total_rows must be set");
  {
    const long long items = rows * a->heads;
    const __nv_bfloat16* o = reinterpret_cast<const __nv_bfloat16*>(a->o);
    const __nv_bfloat16* d_o = reinterpret_cast<const __nv_bfloat16*>(a->d_o);
    if (a->dim_head == 32)
      attn_delta_kernel<32><<<(int)((items + 63) / 64), 256, 0, stream>>>(o, d_o, a->ldo, a->delta, rows, a->heads);
    else
      attn_delta_kernel<64><<<(int)((items + 31) / 32), 256, 0, stream>>>(o, d_o, a->ldo, a->delta, rows, a->heads);
    CTB_LAUNCH_CHECK();
  }
  if (a->cpb_table != nullptr) return ctb_attn_bwd_tc(a, stream);
  if (attn_is_short(a) && a->dbias == nullptr) {
    const size_t smem = (size_t)SH_BWD_WARPS * (2 * 4 * SH_TILE + 128) * 2;
    CTB_CUDA(cudaFuncSetAttribute(attn_short_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_short_bwd_kernel<<<attn_short_grid(a, SH_BWD_WARPS), SH_BWD_WARPS * 32, smem, stream>>>(*a);
    CTB_LAUNCH_CHECK();
    return CTCLIP_OK;
  }
  // the d-logits spill stores bf16 PAIRS: only with an even sequence length (else the recomputing dbias kernel is used)
  ctclip_attn_args a_nospill = *a;
  a_nospill.ds_scratch = nullptr;
  const ctclip_attn_args* a_dq = (a->ds_scratch != nullptr && a->n % 2 == 0) ? a : &a_nospill;
  if (int rc = attn_route(1, a_dq, stream)) return rc;
  if (int rc = attn_route(2, a, stream)) return rc;
  if (a->dbias != nullptr && a->ds_scratch != nullptr && a->n % 2 == 0) {
    // the dQ kernel spilled its d logits: reduce them over the sequences (replaces the recomputing third pass)
    const long long nn = (long long)a->n * a->n;
    const long long threads = (long long)a->heads * (nn / 4);
    attn_dbias_reduce_kernel<<<(int)((threads + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(a->ds_scratch), a->dbias, a->num_seqs, a->heads, nn);
    CTB_LAUNCH_CHECK();
  } else if (a->dbias != nullptr) {
    // sequence chunks: enough CTAs for ~2 waves at 3 CTAs/SM
    const int tiles = ((a->n + 63) / 64) * ((a->n + 63) / 64) * a->heads;
    int chunks = (6 * num_sms() + tiles - 1) / tiles;
    if (chunks < 1) chunks = 1;
    if (chunks > a->num_seqs) chunks = a->num_seqs;
    dim3 grid((a->n + 63) / 64, (a->n + 63) / 64, a->heads * chunks);
    attn_bwd_dbias_kernel<<<grid, 128, 0, stream>>>(*a, chunks);
    CTB_LAUNCH_CHECK();
  }
  return CTCLIP_OK;
}

extern "C" int ctclip_l2norm_bwd(const void* dxh, int64_t ld_dxh, const void* xraw, int64_t ld_x, const float* scale,
                                 void* dx, int64_t ld_dx, float* dscale, int64_t rows, int32_t heads, int32_t dim_head,
                                 void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(dim_head == 32, "l2norm_bwd: dim_head must be 32");
  CTB_CHECK_ARG(dxh && xraw && scale && dx && dscale && rows > 0 && heads > 0, "l2norm_bwd: bad args");
  CTB_CHECK_ARG(ld_dxh % 8 == 0 && ld_x % 8 == 0 && ld_dx % 8 == 0, "l2norm_bwd: rows must be 16B aligned");
  CTB_CHECK_ARG(rows * heads < (1ll << 31), "l2norm_bwd: rows*heads must fit 31 bits");
  long long ctas = (rows * heads + 63) / 64;
  const long long cap = (long long)num_sms() * 8;
  if (ctas > cap) ctas = cap;
  const unsigned n_items = (unsigned)(rows * heads);
#define L2N_LAUNCH(H)                                                                                                   \
  l2norm_bwd_kernel<H><<<(int)ctas, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(dxh), ld_dxh,              \
                                                     reinterpret_cast<const __nv_bfloat16*>(xraw), ld_x, scale,         \
                                                     reinterpret_cast<__nv_bfloat16*>(dx), ld_dx, dscale, n_items, heads)
  if (heads == 8) L2N_LAUNCH(8);
  else if (heads == 12) L2N_LAUNCH(12);
  else L2N_LAUNCH(0);
#undef L2N_LAUNCH
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}
