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
  if (a->dtype == 1) patchify_kernel<true><<<grid, 256, smem, stream>>>(*a);
  else patchify_kernel<false><<<grid, 256, smem, stream>>>(*a);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}
