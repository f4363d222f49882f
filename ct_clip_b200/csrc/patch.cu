// Tubelet im2col + per-patch standardisation (HBM-bound).
//
// Replaces ctvit.py:171-172: Rearrange('b c (t pt) (h p1) (w p2) -> b t h w (c pt p1 p2)') followed by
// nn.LayerNorm(P) -- minus its affine (gamma, beta), which is folded into the patch Linear
// (ctvit.py:173) at weight-preparation time:  W' = W*diag(gamma), b' = W beta + b.
// Output: x_hat bf16 [b*T*H*W, P], the A operand of the patch GEMM (and of its weight-gradient GEMM).
//
// One CTA per (b, t, h) patch row: every (c, pt, p1) image row is read fully coalesced (int16 HU,
// consumed as x/1000 -- scripts/data.py:122-125 -- or fp32); pass 1 accumulates per-patch moments,
// pass 2 re-reads the slab (L2-resident, <= 384 KB) and scatters standardised bf16 pairs.
#include <stdlib.h>
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/ctclip_b200.h"

namespace ctb {

template <bool kInt16>
__device__ __forceinline__ float2 load_pair(const void* base, long long idx, float scale) {
  if (kInt16) {
    const short2 s = *reinterpret_cast<const short2*>(reinterpret_cast<const short*>(base) + idx);
    return make_float2((float)s.x * scale, (float)s.y * scale);
  } else {
    return *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(base) + idx);
  }
}

template <bool kInt16>
__global__ void __launch_bounds__(256) patchify_kernel(ctclip_patchify_args a) {
  extern __shared__ __align__(16) float sm[];  // [Wt][2] moments, [Wt][2] mean/rstd, then the [p1][W] frame slab
  const int Wt = a.W / a.p2, Ht = a.H / a.p1, Tt = a.F / a.pt;
  const int h = blockIdx.x % Ht;
  const int t = (blockIdx.x / Ht) % Tt;
  const int b = blockIdx.x / (Ht * Tt);
  const int rows = a.C * a.pt * a.p1;
  const int pairs = a.W / 2;
  const int P = a.C * a.pt * a.p1 * a.p2;
  for (int i = threadIdx.x; i < 2 * Wt; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const long long plane = (long long)a.H * a.W;
  // pass 1: moments
  for (int j = threadIdx.x; j < pairs; j += blockDim.x) {
    float s = 0.f, q = 0.f;
    for (int r = 0; r < rows; r++) {
      const int p1 = r % a.p1, pt = (r / a.p1) % a.pt, c = r / (a.p1 * a.pt);
      const long long idx = (((long long)b * a.C + c) * a.F + (t * a.pt + pt)) * plane +
                            (long long)(h * a.p1 + p1) * a.W + 2 * j;
      const float2 v = load_pair<kInt16>(a.video, idx, a.scale);
      s += v.x + v.y;
      q += v.x * v.x + v.y * v.y;
    }
    const int w = (2 * j) / a.p2;
    atomicAdd(&sm[2 * w], s);
    atomicAdd(&sm[2 * w + 1], q);
  }
  __syncthreads();
  for (int w = threadIdx.x; w < Wt; w += blockDim.x) {
    const float mean = sm[2 * w] / P;
    const float var = fmaxf(sm[2 * w + 1] / P - mean * mean, 0.f);
    sm[2 * Wt + 2 * w] = mean;
    sm[2 * Wt + 2 * w + 1] = rsqrtf(var + a.eps);
  }
  __syncthreads();
  // pass 2: standardise + write. One (c, pt) frame slab (p1 rows x W columns) is staged in shared memory, then every
  // warp writes whole patches: p1*p2 CONTIGUOUS output elements per (patch, c, pt) -> fully coalesced bf16x2 stores
  // (the direct scatter wrote 40-byte fragments at 50 % sector efficiency).
  float* slab = sm + 4 * Wt;                       // [p1][W] fp32
  const long long m_base = (((long long)b * Tt + t) * Ht + h) * Wt;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int patch_pairs = (a.p1 * a.p2) / 2;
  for (int cf = 0; cf < a.C * a.pt; cf++) {
    const int pt = cf % a.pt, c = cf / a.pt;
    __syncthreads();
    for (int i = threadIdx.x; i < a.p1 * pairs; i += blockDim.x) {
      const int p1 = i / pairs, j = i % pairs;
      const long long idx = (((long long)b * a.C + c) * a.F + (t * a.pt + pt)) * plane + (long long)(h * a.p1 + p1) * a.W + 2 * j;
      const float2 v = load_pair<kInt16>(a.video, idx, a.scale);
      *reinterpret_cast<float2*>(slab + p1 * a.W + 2 * j) = v;
    }
    __syncthreads();
    for (int w = warp; w < Wt; w += nwarps) {
      const float mean = sm[2 * Wt + 2 * w], rstd = sm[2 * Wt + 2 * w + 1];
      __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(a.xhat) + (m_base + w) * (long long)a.ld_out + (long long)cf * a.p1 * a.p2;
      for (int k = lane; k < patch_pairs; k += 32) {
        const int e = 2 * k, p1 = e / a.p2, x2 = e % a.p2;      // p2 is even: a pair never straddles two rows
        const float2 v = *reinterpret_cast<const float2*>(slab + p1 * a.W + w * a.p2 + x2);
        *reinterpret_cast<uint32_t*>(orow + e) = pack_bf16x2((v.x - mean) * rstd, (v.y - mean) * rstd);
      }
    }
  }
}

// v3: same algorithm and arithmetic order per patch as patchify_kernel, with the index arithmetic strength-reduced: the v2
// kernel spent ~8000 instructions per thread on runtime divisions / modulos (i / pairs, e / p2, r % p1 ...) and ran at 17 %
// of the copy bandwidth. Here every loop is a plain (row, pair) nest or uses per-lane offsets computed once per CTA.
// Requires pairs = W/2 <= blockDim.x (W <= 512) and p1*p2/2 <= 7*32 pairs per patch plane (else the v2 kernel is used).
template <bool kInt16>
__global__ void __launch_bounds__(256) patchify3_kernel(ctclip_patchify_args a) {
  extern __shared__ __align__(16) float sm[];  // [Wt][2] moments, [Wt][2] mean/rstd, then the [p1][W] frame slab
  const int Wt = a.W / a.p2, Ht = a.H / a.p1, Tt = a.F / a.pt;
  const int h = blockIdx.x % Ht;
  const int t = (blockIdx.x / Ht) % Tt;
  const int b = blockIdx.x / (Ht * Tt);
  const int pairs = a.W / 2;
  const int P = a.C * a.pt * a.p1 * a.p2;
  const int tid = threadIdx.x;
  for (int i = tid; i < 2 * Wt; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const long long plane = (long long)a.H * a.W;
  const bool col_ok = tid < pairs;
  // pass 1: moments; thread = one column pair, rows walked with pointer increments (4 rows in flight)
  {
    float s = 0.f, q = 0.f;
    if (col_ok) {
      for (int c = 0; c < a.C; c++)
        for (int pt = 0; pt < a.pt; pt++) {
          const long long base = (((long long)b * a.C + c) * a.F + (t * a.pt + pt)) * plane + (long long)(h * a.p1) * a.W + 2 * tid;
          int p1 = 0;
          for (; p1 + 4 <= a.p1; p1 += 4) {
            const float2 v0 = load_pair<kInt16>(a.video, base + (long long)(p1 + 0) * a.W, a.scale);
            const float2 v1 = load_pair<kInt16>(a.video, base + (long long)(p1 + 1) * a.W, a.scale);
            const float2 v2 = load_pair<kInt16>(a.video, base + (long long)(p1 + 2) * a.W, a.scale);
            const float2 v3 = load_pair<kInt16>(a.video, base + (long long)(p1 + 3) * a.W, a.scale);
            s += v0.x + v0.y; q += v0.x * v0.x + v0.y * v0.y;
            s += v1.x + v1.y; q += v1.x * v1.x + v1.y * v1.y;
            s += v2.x + v2.y; q += v2.x * v2.x + v2.y * v2.y;
            s += v3.x + v3.y; q += v3.x * v3.x + v3.y * v3.y;
          }
          for (; p1 < a.p1; p1++) {
            const float2 v = load_pair<kInt16>(a.video, base + (long long)p1 * a.W, a.scale);
            s += v.x + v.y; q += v.x * v.x + v.y * v.y;
          }
        }
      const int w = (2 * tid) / a.p2;
      atomicAdd(&sm[2 * w], s);
      atomicAdd(&sm[2 * w + 1], q);
    }
  }
  __syncthreads();
  for (int w = tid; w < Wt; w += blockDim.x) {
    const float mean = sm[2 * w] / P;
    const float var = fmaxf(sm[2 * w + 1] / P - mean * mean, 0.f);
    sm[2 * Wt + 2 * w] = mean;
    sm[2 * Wt + 2 * w + 1] = rsqrtf(var + a.eps);
  }
  __syncthreads();
  // pass 2: per (c, pt) frame: slab rows -> shared memory (one pair per thread and row), then whole patches out
  float* slab = sm + 4 * Wt;                       // [p1][W] fp32
  const long long m_base = (((long long)b * Tt + t) * Ht + h) * Wt;
  const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  const int patch_pairs = (a.p1 * a.p2) / 2;
  // slab offset (relative to the patch's first column) of this lane's i-th output pair: computed once
  int soff[7];
#pragma unroll
  for (int i = 0; i < 7; i++) {
    const int e = 2 * (lane + 32 * i);
    soff[i] = (e < 2 * patch_pairs) ? (e / a.p2) * a.W + (e % a.p2) : -1;
  }
  for (int cf = 0; cf < a.C * a.pt; cf++) {
    const int pt = cf % a.pt, c = cf / a.pt;
    const long long base = (((long long)b * a.C + c) * a.F + (t * a.pt + pt)) * plane + (long long)(h * a.p1) * a.W + 2 * tid;
    __syncthreads();
    if (col_ok)
      for (int p1 = 0; p1 < a.p1; p1++) {
        const float2 v = load_pair<kInt16>(a.video, base + (long long)p1 * a.W, a.scale);
        *reinterpret_cast<float2*>(slab + p1 * a.W + 2 * tid) = v;
      }
    __syncthreads();
    for (int w = warp; w < Wt; w += nwarps) {
      const float mean = sm[2 * Wt + 2 * w], rstd = sm[2 * Wt + 2 * w + 1];
      __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(a.xhat) + (m_base + w) * (long long)a.ld_out + (long long)cf * a.p1 * a.p2;
      const float* sw = slab + w * a.p2;
#pragma unroll
      for (int i = 0; i < 7; i++) {
        if (soff[i] >= 0) {
          const float2 v = *reinterpret_cast<const float2*>(sw + soff[i]);
          *reinterpret_cast<uint32_t*>(orow + 2 * (lane + 32 * i)) = pack_bf16x2((v.x - mean) * rstd, (v.y - mean) * rstd);
        }
      }
    }
  }
}

}  // namespace ctb

using namespace ctb;

extern "C" int ctclip_patchify(const ctclip_patchify_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(a && a->video && a->xhat, "patchify: null pointer");
  CTB_CHECK_ARG(a->B > 0 && a->C > 0 && a->F > 0 && a->H > 0 && a->W > 0, "patchify: bad volume shape");
  CTB_CHECK_ARG(a->F % a->pt == 0 && a->H % a->p1 == 0 && a->W % a->p2 == 0, "patchify: volume not divisible by patch");
  CTB_CHECK_ARG(a->p2 % 2 == 0 && a->W % 2 == 0, "patchify: patch width must be even (pair-vectorised)");
  CTB_CHECK_ARG(a->ld_out % 2 == 0 && a->ld_out >= a->C * a->pt * a->p1 * a->p2, "patchify: bad ld_out");
  CTB_CHECK_ARG(a->dtype == 0 || a->dtype == 1, "patchify: dtype must be 0 (f32) or 1 (int16)");
  const int Wt = a->W / a->p2;
  const int grid = a->B * (a->F / a->pt) * (a->H / a->p1);
  const size_t smem = sizeof(float) * (4 * Wt + (size_t)a->p1 * a->W);
  // measured on B200 at configs[1] (tools/small_probe.py): 1596 us (v2) -> 672 us (v3), outputs equal to one bf16 ulp (the
  // per-patch moments are reduced with shared-memory atomics in both). CTCLIP_PATCHIFY_V3=0 selects the v2 kernel.
  static const int v3 = getenv("CTCLIP_PATCHIFY_V3") ? atoi(getenv("CTCLIP_PATCHIFY_V3")) : 1;   // debug knob, read once
  const bool v3_ok = v3 && a->W / 2 <= 256 && (a->p1 * a->p2) / 2 <= 7 * 32;
  if (v3_ok) {
    if (a->dtype == 1) patchify3_kernel<true><<<grid, 256, smem, stream>>>(*a);
    else patchify3_kernel<false><<<grid, 256, smem, stream>>>(*a);
  } else {
    if (a->dtype == 1) patchify_kernel<true><<<grid, 256, smem, stream>>>(*a);
    else patchify_kernel<false><<<grid, 256, smem, stream>>>(*a);
  }
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}
