// PEG: depthwise causal 3x3x3 convolution + residual on the fp32 token stream (HBM/L1-bound stencil).
//
// Replaces attention.py:63-84 (PEG.forward, causal=True: F.pad (1,1),(1,1),(2,0) + nn.Conv3d(groups=dim))
// and the `peg(x) + x` residual at attention.py:324, plus their backward.
//
// Tokens live in ONE canonical layout [b, t, h, w, D] for the whole encoder (no re-layout copies,
// ctvit.py:291/297/301/305). The reference calls PEG on whatever memory order the current stack
// uses and *reshapes* it to (b, T, H, W) (attention.py:70), so:
//   spatial stack : memory order (b,t,h,w)  -> conv grid coordinate f == canonical token index;
//   temporal stack: memory order (b,h,w,t) re-read as (T,H,W) (SURVEY trap T1): the conv-grid flat
//                   index f = (ih*W + iw)*T + it, canonical token = (it*H + ih)*W + iw.
// The kernel walks the conv grid (a0,a1,a2) in (T,H,W) shape and maps every access through
// canon(f); a0 is the causal axis (taps a0-2, a0-1, a0).
//
// Thread = one channel pair; CTA = `lines` consecutive (a0,a1) lines of one volume; sliding window
// along a2 so every input element is loaded 9x (not 27x) from L1/L2.
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/ctclip_b200.h"

namespace ctb {

struct PegGeom {
  int T, H, W, D, temporal;
};

__device__ __forceinline__ long long peg_canon(const PegGeom& g, int a0, int a1, int a2) {
  const int f = (a0 * g.H + a1) * g.W + a2;
  if (!g.temporal) return f;
  const int it = f % g.T;
  const int iw = (f / g.T) % g.W;
  const int ih = f / (g.T * g.W);
  return ((long long)it * g.H + ih) * g.W + iw;
}

// MODE 0: y = x + conv(x) (+bias)           (forward)
// MODE 1: dx = dy + conv^T(dy)              (backward data; taps mirrored, no bias)
template <int MODE>
__global__ void __launch_bounds__(384) peg_conv_kernel(ctclip_peg_args a) {
  const PegGeom g{a.T, a.H, a.W, a.D, a.temporal};
  const int c = 2 * threadIdx.x;  // channel pair
  if (c >= a.D) return;
  const int lines_total = a.T * a.H;
  const int line0 = blockIdx.x * a.lines;
  const int b = blockIdx.y;
  const float* xin = a.x + (long long)b * a.T * a.H * a.W * a.D;
  float* yout = a.y + (long long)b * a.T * a.H * a.W * a.D;
  __nv_bfloat16* ybf = a.y_bf16 ? reinterpret_cast<__nv_bfloat16*>(a.y_bf16) + (long long)b * a.T * a.H * a.W * a.D : nullptr;

  float2 wt[27];
#pragma unroll
  for (int k = 0; k < 27; k++) {
    const int kk = (MODE == 0) ? k : (26 - k);  // mirrored taps for the transposed conv
    wt[k] = make_float2(a.weight[(long long)c * 27 + kk], a.weight[(long long)(c + 1) * 27 + kk]);
  }
  float2 bias = make_float2(0.f, 0.f);
  if (MODE == 0 && a.bias != nullptr) bias = make_float2(a.bias[c], a.bias[c + 1]);

  for (int li = 0; li < a.lines; li++) {
    const int line = line0 + li;
    if (line >= lines_total) break;
    const int a0 = line / a.H, a1 = line % a.H;
    // tap k0 reads a0 + k0 - 2 (forward) or a0 + k0 (mirrored: original offset 2-k0' with k0'=2-k0)
    float2 win[9][3];
    long long rowbase_valid[9];
#pragma unroll
    for (int r = 0; r < 9; r++) {
      const int k0 = r / 3, k1 = r % 3;
      const int n0 = (MODE == 0) ? (a0 + k0 - 2) : (a0 + k0);
      const int n1 = a1 + k1 - 1;
      rowbase_valid[r] = (n0 >= 0 && n0 < a.T && n1 >= 0 && n1 < a.H) ? ((long long)n0 << 32 | (unsigned)n1) : -1;
      win[r][0] = make_float2(0.f, 0.f);  // position a2-1 = -1 (padding)
      win[r][1] = make_float2(0.f, 0.f);
      if (rowbase_valid[r] >= 0)
        win[r][1] = *reinterpret_cast<const float2*>(xin + peg_canon(g, n0, n1, 0) * a.D + c);
    }
    for (int a2 = 0; a2 < a.W; a2++) {
      float2 acc = bias;
#pragma unroll
      for (int r = 0; r < 9; r++) {
        float2 nx = make_float2(0.f, 0.f);
        if (rowbase_valid[r] >= 0 && a2 + 1 < a.W) {
          const int n0 = (int)(rowbase_valid[r] >> 32), n1 = (int)(rowbase_valid[r] & 0xffffffff);
          nx = *reinterpret_cast<const float2*>(xin + peg_canon(g, n0, n1, a2 + 1) * a.D + c);
        }
        win[r][2] = nx;
#pragma unroll
        for (int k2 = 0; k2 < 3; k2++) {
          acc.x = fmaf(wt[r * 3 + k2].x, win[r][k2].x, acc.x);
          acc.y = fmaf(wt[r * 3 + k2].y, win[r][k2].y, acc.y);
        }
        win[r][0] = win[r][1];
        win[r][1] = win[r][2];
      }
      const long long tok = peg_canon(g, a0, a1, a2);
      const float2 ctr = *reinterpret_cast<const float2*>(xin + tok * a.D + c);
      acc.x += ctr.x;
      acc.y += ctr.y;
      *reinterpret_cast<float2*>(yout + tok * a.D + c) = acc;
      if (ybf != nullptr) *reinterpret_cast<uint32_t*>(ybf + tok * a.D + c) = pack_bf16x2(acc.x, acc.y);
    }
  }
}

// Weight / bias gradient: dw[c][k] += sum_p dy[p] * x[p + off(k)], db[c] += sum_p dy[p].
__global__ void __launch_bounds__(384) peg_wgrad_kernel(ctclip_peg_args a) {
  const PegGeom g{a.T, a.H, a.W, a.D, a.temporal};
  const int c = 2 * threadIdx.x;
  if (c >= a.D) return;
  const int lines_total = a.T * a.H;
  const int line0 = blockIdx.x * a.lines;
  const int b = blockIdx.y;
  const float* xin = a.x + (long long)b * a.T * a.H * a.W * a.D;    // forward input
  const float* dy = a.dy + (long long)b * a.T * a.H * a.W * a.D;    // upstream gradient
  float2 acc[27];
#pragma unroll
  for (int k = 0; k < 27; k++) acc[k] = make_float2(0.f, 0.f);
  float2 accb = make_float2(0.f, 0.f);
  for (int li = 0; li < a.lines; li++) {
    const int line = line0 + li;
    if (line >= lines_total) break;
    const int a0 = line / a.H, a1 = line % a.H;
    float2 win[9][3];
    long long rv[9];
#pragma unroll
    for (int r = 0; r < 9; r++) {
      const int k0 = r / 3, k1 = r % 3;
      const int n0 = a0 + k0 - 2, n1 = a1 + k1 - 1;
      rv[r] = (n0 >= 0 && n0 < a.T && n1 >= 0 && n1 < a.H) ? ((long long)n0 << 32 | (unsigned)n1) : -1;
      win[r][0] = make_float2(0.f, 0.f);
      win[r][1] = make_float2(0.f, 0.f);
      if (rv[r] >= 0) win[r][1] = *reinterpret_cast<const float2*>(xin + peg_canon(g, n0, n1, 0) * a.D + c);
    }
    for (int a2 = 0; a2 < a.W; a2++) {
      const float2 d = *reinterpret_cast<const float2*>(dy + peg_canon(g, a0, a1, a2) * a.D + c);
      accb.x += d.x;
      accb.y += d.y;
#pragma unroll
      for (int r = 0; r < 9; r++) {
        float2 nx = make_float2(0.f, 0.f);
        if (rv[r] >= 0 && a2 + 1 < a.W) {
          const int n0 = (int)(rv[r] >> 32), n1 = (int)(rv[r] & 0xffffffff);
          nx = *reinterpret_cast<const float2*>(xin + peg_canon(g, n0, n1, a2 + 1) * a.D + c);
        }
        win[r][2] = nx;
#pragma unroll
        for (int k2 = 0; k2 < 3; k2++) {
          acc[r * 3 + k2].x = fmaf(d.x, win[r][k2].x, acc[r * 3 + k2].x);
          acc[r * 3 + k2].y = fmaf(d.y, win[r][k2].y, acc[r * 3 + k2].y);
        }
        win[r][0] = win[r][1];
        win[r][1] = win[r][2];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 27; k++) {
    atomicAdd(a.dweight + (long long)c * 27 + k, acc[k].x);
    atomicAdd(a.dweight + (long long)(c + 1) * 27 + k, acc[k].y);
  }
  if (a.dbias != nullptr) {
    atomicAdd(a.dbias + c, accb.x);
    atomicAdd(a.dbias + c + 1, accb.y);
  }
}

}  // namespace ctb

using namespace ctb;

static int peg_check(const ctclip_peg_args* a, const char* who) {
  CTB_CHECK_ARG(a && a->x, "%s: null x", who);
  CTB_CHECK_ARG(a->B > 0 && a->T > 0 && a->H > 0 && a->W > 0, "%s: bad grid", who);
  CTB_CHECK_ARG(a->D % 2 == 0 && a->D <= 768, "%s: D must be even and <= 768", who);
  CTB_CHECK_ARG(a->lines >= 1, "%s: lines must be >= 1", who);
  return CTCLIP_OK;
}

extern "C" int ctclip_peg_fwd(const ctclip_peg_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = peg_check(a, "peg_fwd")) return rc;
  CTB_CHECK_ARG(a->y && a->weight, "peg_fwd: null y/weight");
  dim3 grid(ceil_div(a->T * a->H, a->lines), a->B);
  peg_conv_kernel<0><<<grid, a->D / 2, 0, stream>>>(*a);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}

// x = upstream gradient dy (fp32), y = dx out (fp32), y_bf16 = optional bf16 copy of dx
extern "C" int ctclip_peg_bwd_data(const ctclip_peg_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = peg_check(a, "peg_bwd_data")) return rc;
  CTB_CHECK_ARG(a->y && a->weight, "peg_bwd_data: null y/weight");
  dim3 grid(ceil_div(a->T * a->H, a->lines), a->B);
  peg_conv_kernel<1><<<grid, a->D / 2, 0, stream>>>(*a);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}

// x = forward input, dy = upstream gradient; dweight [D,27] and dbias [D] are accumulated (atomics)
extern "C" int ctclip_peg_bwd_weight(const ctclip_peg_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = peg_check(a, "peg_bwd_weight")) return rc;
  CTB_CHECK_ARG(a->dy && a->dweight, "peg_bwd_weight: null dy/dweight");
  dim3 grid(ceil_div(a->T * a->H, a->lines), a->B);
  peg_wgrad_kernel<<<grid, a->D / 2, 0, stream>>>(*a);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}
