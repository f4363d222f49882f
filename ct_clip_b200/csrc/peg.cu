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
// See the v2 kernel comment below for the tiling (rolling 3-plane shared-memory buffer along the causal axis).
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

// ------------------------------------------------------------------------------------------------
// v2 kernels. CTA = (32-channel block, tile of A1T=8 conv-grid lines along a1, range of a0 planes, volume).
// The CTA walks the causal axis a0 with a ROLLING 3-plane buffer in shared memory ([slot][a1 halo][a2 halo][32 ch] fp32):
// every step loads ONE new plane (halo factor (A1T+2)(W+2)/(A1T*W) ~ 1.35x of compulsory traffic, 128 B coalesced
// segments), then each thread (= one line x one channel) slides a 3-wide register window along a2 (9 LDS per output,
// 27 taps). Warps read 32 consecutive channels -> conflict-free LDS and 128 B coalesced stores.
// ------------------------------------------------------------------------------------------------
constexpr int A1T = 8;   // lines per CTA tile
constexpr int CB = 32;   // channels per CTA

struct PegTile {
  int a1_0, c0, b, p_begin, p_end;  // a0 plane range [p_begin, p_end) produced by this CTA
};

// load plane `a0` (rows a1_0-1 .. a1_0+A1T, cols -1 .. W) of a [tokens, D] fp32 tensor into one smem slot; zero outside
__device__ __forceinline__ void peg_load_plane(float* slot, const float* __restrict__ src, const PegGeom& g, int a0, int a1_0) {
  const int a2h = g.W + 2;
  const int n_tok = (A1T + 2) * a2h;
  for (int idx = threadIdx.x; idx < n_tok * (CB / 4); idx += blockDim.x) {
    const int tok = idx / (CB / 4), q = idx % (CB / 4);
    const int r1 = tok / a2h, r2 = tok % a2h;
    const int a1 = a1_0 - 1 + r1, a2 = r2 - 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a0 >= 0 && a0 < g.T && a1 >= 0 && a1 < g.H && a2 >= 0 && a2 < g.W)
      v = *reinterpret_cast<const float4*>(src + peg_canon(g, a0, a1, a2) * g.D + q * 4);
    *reinterpret_cast<float4*>(slot + (size_t)tok * CB + q * 4) = v;
  }
}

// MODE 0: y = x + conv(x) + bias (forward; taps a0-2..a0)     MODE 1: dx = dy + conv^T(dy) (taps a0..a0+2, mirrored)
template <int MODE>
__global__ void __launch_bounds__(256, 2) peg_conv2_kernel(ctclip_peg_args a, int planes_per_cta) {
  extern __shared__ __align__(16) float peg_sm[];
  const PegGeom g{a.T, a.H, a.W, a.D, a.temporal};
  const int a2h = a.W + 2;
  const size_t slot_elems = (size_t)(A1T + 2) * a2h * CB;
  const int n_a1t = (a.H + A1T - 1) / A1T;
  const int c0 = (blockIdx.x % (a.D / CB)) * CB;
  const int a1_0 = ((blockIdx.x / (a.D / CB)) % n_a1t) * A1T;
  const int chunk = blockIdx.x / ((a.D / CB) * n_a1t);
  const int b = blockIdx.y;
  const int p_begin = chunk * planes_per_cta;
  const int p_end = min(a.T, p_begin + planes_per_cta);
  if (p_begin >= p_end) return;
  const long long vol = (long long)a.T * a.H * a.W * a.D;
  const float* xin = a.x + (long long)b * vol + c0;
  float* yout = a.y + (long long)b * vol + c0;
  __nv_bfloat16* ybf = a.y_bf16 ? reinterpret_cast<__nv_bfloat16*>(a.y_bf16) + (long long)b * vol + c0 : nullptr;
  const int lane = threadIdx.x & 31, line = threadIdx.x >> 5;  // channel, a1 line within the tile
  const int ch = c0 + lane;
  float wt[27];
#pragma unroll
  for (int k = 0; k < 27; k++) wt[k] = a.weight[(long long)ch * 27 + ((MODE == 0) ? k : 26 - k)];
  const float bias = (MODE == 0 && a.bias != nullptr) ? a.bias[ch] : 0.f;
  const int a1 = a1_0 + line;
  // prime the rolling buffer with the two planes preceding (MODE 0) / following (MODE 1) the first output plane
  const int first = (MODE == 0) ? p_begin : p_end - 1;
  const int dirn = (MODE == 0) ? 1 : -1;
  for (int d = 2; d >= 1; d--) {
    const int pl = first - dirn * d;  // MODE 0: first-2, first-1   MODE 1: first+2, first+1
    peg_load_plane(peg_sm + (size_t)(((pl % 3) + 3) % 3) * slot_elems, xin, g, pl, a1_0);
  }
  for (int step = 0; step < p_end - p_begin; step++) {
    const int a0 = first + dirn * step;
    __syncthreads();  // everyone is done reading the slot we are about to overwrite
    peg_load_plane(peg_sm + (size_t)(((a0 % 3) + 3) % 3) * slot_elems, xin, g, a0, a1_0);
    __syncthreads();
    if (a1 < a.H) {
      // tap k0 reads plane a0 + k0 - 2 (MODE 0) or a0 + k0 (MODE 1)
      const float* rowp[9];
#pragma unroll
      for (int r = 0; r < 9; r++) {
        const int k0 = r / 3, k1 = r % 3;
        const int pl = (MODE == 0) ? (a0 + k0 - 2) : (a0 + k0);
        rowp[r] = peg_sm + (size_t)(((pl % 3) + 3) % 3) * slot_elems + (size_t)((line + k1) * a2h) * CB + lane;
      }
      float win[9][3];
#pragma unroll
      for (int r = 0; r < 9; r++) {
        win[r][0] = rowp[r][0];       // a2 = -1 (zero padding)
        win[r][1] = rowp[r][CB];      // a2 = 0
      }
      for (int a2 = 0; a2 < a.W; a2++) {
        float acc = bias;
#pragma unroll
        for (int r = 0; r < 9; r++) {
          win[r][2] = rowp[r][(a2 + 2) * CB];
#pragma unroll
          for (int k2 = 0; k2 < 3; k2++) acc = fmaf(wt[r * 3 + k2], win[r][k2], acc);
          win[r][0] = win[r][1];
          win[r][1] = win[r][2];
        }
        // centre tap input = residual term: row r = (k0 = 2 | 0, k1 = 1), window position "a2" was shifted into win[.][0]
        const float ctr = (MODE == 0) ? win[7][0] : win[1][0];
        const long long tok = peg_canon(g, a0, a1, a2);
        const float out = acc + ctr;
        yout[tok * a.D + lane] = out;
        if (ybf != nullptr) ybf[tok * a.D + lane] = __float2bfloat16(out);
      }
    }
  }
}

// dw[c][k] += sum_p dy[p] * x[p + off(k)], db[c] += sum_p dy[p]   (x planes in the rolling buffer, dy straight from global)
__global__ void __launch_bounds__(256, 2) peg_wgrad2_kernel(ctclip_peg_args a, int planes_per_cta) {
  extern __shared__ __align__(16) float peg_sm[];
  const PegGeom g{a.T, a.H, a.W, a.D, a.temporal};
  const int a2h = a.W + 2;
  const size_t slot_elems = (size_t)(A1T + 2) * a2h * CB;
  const int n_a1t = (a.H + A1T - 1) / A1T;
  const int c0 = (blockIdx.x % (a.D / CB)) * CB;
  const int a1_0 = ((blockIdx.x / (a.D / CB)) % n_a1t) * A1T;
  const int chunk = blockIdx.x / ((a.D / CB) * n_a1t);
  const int b = blockIdx.y;
  const int p_begin = chunk * planes_per_cta;
  const int p_end = min(a.T, p_begin + planes_per_cta);
  if (p_begin >= p_end) return;
  const long long vol = (long long)a.T * a.H * a.W * a.D;
  const float* xin = a.x + (long long)b * vol + c0;
  const float* dy = a.dy + (long long)b * vol + c0;
  const int lane = threadIdx.x & 31, line = threadIdx.x >> 5;
  const int a1 = a1_0 + line;
  float acc[27];
#pragma unroll
  for (int k = 0; k < 27; k++) acc[k] = 0.f;
  float accb = 0.f;
  for (int d = 2; d >= 1; d--) {
    const int pl = p_begin - d;
    peg_load_plane(peg_sm + (size_t)(((pl % 3) + 3) % 3) * slot_elems, xin, g, pl, a1_0);
  }
  for (int a0 = p_begin; a0 < p_end; a0++) {
    __syncthreads();
    peg_load_plane(peg_sm + (size_t)(((a0 % 3) + 3) % 3) * slot_elems, xin, g, a0, a1_0);
    __syncthreads();
    if (a1 < a.H) {
      const float* rowp[9];
#pragma unroll
      for (int r = 0; r < 9; r++) {
        const int k0 = r / 3, k1 = r % 3;
        const int pl = a0 + k0 - 2;
        rowp[r] = peg_sm + (size_t)(((pl % 3) + 3) % 3) * slot_elems + (size_t)((line + k1) * a2h) * CB + lane;
      }
      float win[9][3];
#pragma unroll
      for (int r = 0; r < 9; r++) {
        win[r][0] = rowp[r][0];
        win[r][1] = rowp[r][CB];
      }
      for (int a2 = 0; a2 < a.W; a2++) {
        const float d = dy[peg_canon(g, a0, a1, a2) * a.D + lane];
        accb += d;
#pragma unroll
        for (int r = 0; r < 9; r++) {
          win[r][2] = rowp[r][(a2 + 2) * CB];
#pragma unroll
          for (int k2 = 0; k2 < 3; k2++) acc[r * 3 + k2] = fmaf(d, win[r][k2], acc[r * 3 + k2]);
          win[r][0] = win[r][1];
          win[r][1] = win[r][2];
        }
      }
    }
  }
  // reduce the A1T lines of this CTA (same channel = same lane) through shared memory, one atomic per (channel, tap)
  __syncthreads();
  float* red = peg_sm;  // [A1T][28][CB]
#pragma unroll
  for (int k = 0; k < 27; k++) red[(line * 28 + k) * CB + lane] = acc[k];
  red[(line * 28 + 27) * CB + lane] = accb;
  __syncthreads();
  for (int i = threadIdx.x; i < 28 * CB; i += blockDim.x) {
    const int k = i / CB, c = i % CB;
    float t = 0.f;
#pragma unroll
    for (int l = 0; l < A1T; l++) t += red[(l * 28 + k) * CB + c];
    if (k < 27) atomicAdd(a.dweight + (long long)(c0 + c) * 27 + k, t);
    else if (a.dbias != nullptr) atomicAdd(a.dbias + c0 + c, t);
  }
}

}  // namespace ctb

using namespace ctb;

static int peg_check(const ctclip_peg_args* a, const char* who) {
  CTB_CHECK_ARG(a && a->x, "%s: null x", who);
  CTB_CHECK_ARG(a->B > 0 && a->T > 0 && a->H > 0 && a->W > 0, "%s: bad grid", who);
  CTB_CHECK_ARG(a->D % 32 == 0, "%s: D must be a multiple of 32", who);
  return CTCLIP_OK;
}

// grid.x = channel blocks x a1 tiles x a0 chunks; a0 chunks sized so that the grid has >= ~3 CTAs per SM
static void peg_launch_shape(const ctclip_peg_args* a, dim3* grid, int* planes_per_cta, size_t* smem) {
  const int n_a1t = (a->H + A1T - 1) / A1T;
  const long long base = (long long)(a->D / CB) * n_a1t * a->B;
  int chunks = (int)((3LL * num_sms() + base - 1) / base);
  if (chunks < 1) chunks = 1;
  if (chunks > a->T) chunks = a->T;
  *planes_per_cta = (a->T + chunks - 1) / chunks;
  chunks = (a->T + *planes_per_cta - 1) / *planes_per_cta;
  *grid = dim3((unsigned)((a->D / CB) * n_a1t * chunks), (unsigned)a->B);
  *smem = sizeof(float) * 3 * (size_t)(A1T + 2) * (a->W + 2) * CB;
  const size_t red_bytes = sizeof(float) * A1T * 28 * CB;  // weight-gradient reduction scratch
  if (*smem < red_bytes) *smem = red_bytes;
}

template <typename Kern>
static int peg_launch(Kern kern, const ctclip_peg_args* a, cudaStream_t stream) {
  dim3 grid;
  int ppc;
  size_t smem;
  peg_launch_shape(a, &grid, &ppc, &smem);
  CTB_CHECK_ARG(smem <= 227 * 1024, "peg: token grid width %d needs %zu B of shared memory", a->W, smem);
  CTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, 256, smem, stream>>>(*a, ppc);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}

extern "C" int ctclip_peg_fwd(const ctclip_peg_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = peg_check(a, "peg_fwd")) return rc;
  CTB_CHECK_ARG(a->y && a->weight, "peg_fwd: null y/weight");
  return peg_launch(peg_conv2_kernel<0>, a, stream);
}

// x = upstream gradient dy (fp32), y = dx out (fp32), y_bf16 = optional bf16 copy of dx
extern "C" int ctclip_peg_bwd_data(const ctclip_peg_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = peg_check(a, "peg_bwd_data")) return rc;
  CTB_CHECK_ARG(a->y && a->weight, "peg_bwd_data: null y/weight");
  return peg_launch(peg_conv2_kernel<1>, a, stream);
}

// x = forward input, dy = upstream gradient; dweight [D,27] and dbias [D] are accumulated (atomics)
extern "C" int ctclip_peg_bwd_weight(const ctclip_peg_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = peg_check(a, "peg_bwd_weight")) return rc;
  CTB_CHECK_ARG(a->dy && a->dweight, "peg_bwd_weight: null dy/dweight");
  return peg_launch(peg_wgrad2_kernel, a, stream);
}
