// LayerNorm forward / backward over rows of an fp32 [M, D] tensor (HBM-bound; one warp per row,
// 128-bit coalesced accesses, row kept in registers, two-pass statistics in fp32).
//
// Replaces F.layer_norm at attention.py:35 (LayerNorm with zero beta buffer), attention.py:47
// (nn.LayerNorm in FeedForward), ctvit.py:174 (LayerNorm(dim) after the patch Linear),
// attention.py:333 (norm_out) and BERT's LayerNorms. In the pre-norm layers the affine (gamma,
// beta) is folded into the following Linear's weights at weight-preparation time, so the kernel
// emits the standardised row x_hat (bf16, the GEMM operand saved for backward) and rstd.
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/ctclip_b200.h"

namespace ctb {

template <int NCH>  // D = NCH * 128
__global__ void __launch_bounds__(256) ln_fwd_kernel(ctclip_ln_fwd_args a) {
  const int warps_per_cta = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = NCH * 128;
  for (long long row = (long long)blockIdx.x * warps_per_cta + warp; row < a.M;
       row += (long long)gridDim.x * warps_per_cta) {
    const float* xr = a.x + row * D;
    float4 v[NCH];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; i++) {
      v[i] = *reinterpret_cast<const float4*>(xr + i * 128 + lane * 4);
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
    const float mean = warp_sum(s) / D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; i++) {
      const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
      q += dx * dx + dy * dy + dz * dz + dw * dw;
    }
    const float rstd = rsqrtf(warp_sum(q) / D + a.eps);
    if (a.rstd_out != nullptr && lane == 0) a.rstd_out[row] = rstd;
#pragma unroll
    for (int i = 0; i < NCH; i++) {
      const int c = i * 128 + lane * 4;
      if (a.raw_bf16 != nullptr) {
        uint2 u = make_uint2(pack_bf16x2(v[i].x, v[i].y), pack_bf16x2(v[i].z, v[i].w));
        *reinterpret_cast<uint2*>(a.raw_bf16 + row * D + c) = u;
      }
      float4 h = make_float4((v[i].x - mean) * rstd, (v[i].y - mean) * rstd, (v[i].z - mean) * rstd,
                             (v[i].w - mean) * rstd);
      if (a.xhat_bf16 != nullptr) {
        uint2 u = make_uint2(pack_bf16x2(h.x, h.y), pack_bf16x2(h.z, h.w));
        *reinterpret_cast<uint2*>(a.xhat_bf16 + row * D + c) = u;
      }
      if (a.y_f32 != nullptr || a.y_bf16 != nullptr) {
        const float4 g = *reinterpret_cast<const float4*>(a.gamma + c);
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.beta != nullptr) b = *reinterpret_cast<const float4*>(a.beta + c);
        float4 y = make_float4(h.x * g.x + b.x, h.y * g.y + b.y, h.z * g.z + b.z, h.w * g.w + b.w);
        if (a.y_f32 != nullptr) *reinterpret_cast<float4*>(a.y_f32 + row * D + c) = y;
        if (a.y_bf16 != nullptr) {
          uint2 u = make_uint2(pack_bf16x2(y.x, y.y), pack_bf16x2(y.z, y.w));
          *reinterpret_cast<uint2*>(a.y_bf16 + row * D + c) = u;
        }
      }
    }
  }
}

// Backward. g = upstream gradient w.r.t. the LN output (if gamma != NULL: dxhat = g*gamma, and
// dgamma += sum g*xhat, dbeta += sum g) or directly w.r.t. x_hat (gamma == NULL).
//   dx = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat*xhat))  (+ dres_in) (+ add_bf16)
template <int NCH>
__global__ void __launch_bounds__(256) ln_bwd_kernel(ctclip_ln_bwd_args a) {
  const int warps_per_cta = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = NCH * 128;
  float4 accg[NCH], accb[NCH];
#pragma unroll
  for (int i = 0; i < NCH; i++) {
    accg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    accb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const bool want_param = (a.dgamma != nullptr) || (a.dbeta != nullptr);
  for (long long row = (long long)blockIdx.x * warps_per_cta + warp; row < a.M;
       row += (long long)gridDim.x * warps_per_cta) {
    float4 g[NCH], h[NCH], rin[NCH];
    uint2 radd[NCH];
    float s1 = 0.f, s2 = 0.f;
    // every load of the row is issued before the first reduction: the residual-gradient reads used to start only after
    // the two warp reductions, i.e. a second exposed DRAM round trip per row (52-56 % of the copy bandwidth)
#pragma unroll
    for (int i = 0; i < NCH; i++) {
      const int c = i * 128 + lane * 4;
      if (a.dres_in != nullptr) rin[i] = *reinterpret_cast<const float4*>(a.dres_in + row * D + c);
      if (a.add_bf16 != nullptr) radd[i] = *reinterpret_cast<const uint2*>(a.add_bf16 + row * D + c);
    }
    const float rstd = a.rstd[row];
#pragma unroll
    for (int i = 0; i < NCH; i++) {
      const int c = i * 128 + lane * 4;
      if (a.g_f32 != nullptr) {
        g[i] = *reinterpret_cast<const float4*>(a.g_f32 + row * D + c);
      } else {
        const uint2 u = *reinterpret_cast<const uint2*>(a.g_bf16 + row * D + c);
        const float2 p0 = unpack_bf16x2(u.x), p1 = unpack_bf16x2(u.y);
        g[i] = make_float4(p0.x, p0.y, p1.x, p1.y);
      }
      const uint2 uh = *reinterpret_cast<const uint2*>(a.xhat + row * D + c);
      const float2 h0 = unpack_bf16x2(uh.x), h1 = unpack_bf16x2(uh.y);
      h[i] = make_float4(h0.x, h0.y, h1.x, h1.y);
      if (want_param) {
        accg[i].x += g[i].x * h[i].x; accg[i].y += g[i].y * h[i].y;
        accg[i].z += g[i].z * h[i].z; accg[i].w += g[i].w * h[i].w;
        accb[i].x += g[i].x; accb[i].y += g[i].y; accb[i].z += g[i].z; accb[i].w += g[i].w;
      }
      if (a.gamma != nullptr) {
        const float4 gm = *reinterpret_cast<const float4*>(a.gamma + c);
        g[i].x *= gm.x; g[i].y *= gm.y; g[i].z *= gm.z; g[i].w *= gm.w;
      }
      s1 += g[i].x + g[i].y + g[i].z + g[i].w;
      s2 += g[i].x * h[i].x + g[i].y * h[i].y + g[i].z * h[i].z + g[i].w * h[i].w;
    }
    const float c1 = warp_sum(s1) / D;
    const float c2 = warp_sum(s2) / D;
#pragma unroll
    for (int i = 0; i < NCH; i++) {
      const int c = i * 128 + lane * 4;
      float4 dx = make_float4(rstd * (g[i].x - c1 - h[i].x * c2), rstd * (g[i].y - c1 - h[i].y * c2),
                              rstd * (g[i].z - c1 - h[i].z * c2), rstd * (g[i].w - c1 - h[i].w * c2));
      if (a.dres_in != nullptr) {
        const float4 r = rin[i];
        dx.x += r.x; dx.y += r.y; dx.z += r.z; dx.w += r.w;
      }
      if (a.add_bf16 != nullptr) {
        const uint2 u = radd[i];
        const float2 p0 = unpack_bf16x2(u.x), p1 = unpack_bf16x2(u.y);
        dx.x += p0.x; dx.y += p0.y; dx.z += p1.x; dx.w += p1.y;
      }
      if (a.dx_f32 != nullptr) *reinterpret_cast<float4*>(a.dx_f32 + row * D + c) = dx;
      if (a.dx_bf16 != nullptr) {
        uint2 u = make_uint2(pack_bf16x2(dx.x, dx.y), pack_bf16x2(dx.z, dx.w));
        *reinterpret_cast<uint2*>(a.dx_bf16 + row * D + c) = u;
      }
    }
  }
  if (want_param) {
    // reduce the per-warp column partials across the CTA, then one atomic per column per CTA
    __shared__ float red[8][NCH * 128];
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {
      float* dst = pass == 0 ? a.dgamma : a.dbeta;
      if (dst == nullptr) continue;
      __syncthreads();
#pragma unroll
      for (int i = 0; i < NCH; i++) {
        const float4 vv = pass == 0 ? accg[i] : accb[i];
        *reinterpret_cast<float4*>(&red[warp][i * 128 + lane * 4]) = vv;
      }
      __syncthreads();
      for (int c = threadIdx.x; c < D; c += blockDim.x) {
        float t = 0.f;
        for (int w = 0; w < warps_per_cta; w++) t += red[w][c];
        atomicAdd(dst + c, t);
      }
    }
  }
}

}  // namespace ctb

using namespace ctb;

extern "C" int ctclip_ln_fwd(const ctclip_ln_fwd_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(a && a->x && a->M > 0, "ln_fwd: bad args");
  CTB_CHECK_ARG(a->D % 128 == 0 && a->D >= 128 && a->D <= 1024, "ln_fwd: D=%d must be a multiple of 128 in [128,1024]", a->D);
  CTB_CHECK_ARG((a->y_f32 == nullptr && a->y_bf16 == nullptr) || a->gamma != nullptr, "ln_fwd: affine output needs gamma");
  const int wpc = 8;
  long long ctas = (a->M + wpc - 1) / wpc;
  const long long cap = (long long)num_sms() * 16;
  if (ctas > cap) ctas = cap;
  switch (a->D / 128) {
#define CASE(n) case n: ln_fwd_kernel<n><<<(int)ctas, wpc * 32, 0, stream>>>(*a); break;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
  }
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}

extern "C" int ctclip_ln_bwd(const ctclip_ln_bwd_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(a && a->xhat && a->rstd && a->M > 0, "ln_bwd: bad args");
  CTB_CHECK_ARG((a->g_f32 != nullptr) != (a->g_bf16 != nullptr), "ln_bwd: exactly one of g_f32 / g_bf16");
  CTB_CHECK_ARG(a->D % 128 == 0 && a->D >= 128 && a->D <= 1024, "ln_bwd: D=%d must be a multiple of 128 in [128,1024]", a->D);
  const int wpc = 8;
  long long ctas = (a->M + wpc - 1) / wpc;
  // parameter-gradient partials are flushed once per CTA: keep the grid to ~2 CTAs/SM worth
  const long long cap = (a->dgamma || a->dbeta) ? (long long)num_sms() * 4 : (long long)num_sms() * 16;
  if (ctas > cap) ctas = cap;
  switch (a->D / 128) {
#define CASE(n) case n: ln_bwd_kernel<n><<<(int)ctas, wpc * 32, 0, stream>>>(*a); break;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
  }
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}
