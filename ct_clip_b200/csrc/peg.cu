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
  const int* table;   // optional precomputed canon(f) (avoids three integer divisions per access in the temporal stack)
};

__device__ __forceinline__ long long peg_canon(const PegGeom& g, int a0, int a1, int a2) {
  const int f = (a0 * g.H + a1) * g.W + a2;
  if (!g.temporal) return f;
  if (g.table != nullptr) return __ldg(g.table + f);
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
constexpr int A1T = 8;     // lines per CTA tile
constexpr int CB = 32;     // channels per CTA
constexpr int NSLOT = 4;   // ring of planes: 3 live + 1 being filled by cp.async while the current plane is computed

__device__ __forceinline__ int peg_slot(int pl) { return ((pl % NSLOT) + NSLOT) % NSLOT; }

__device__ __forceinline__ void peg_cp16(void* smem_dst, const void* gmem_src, bool valid) {
  const int bytes = valid ? 16 : 0;   // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void peg_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void peg_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// async load of plane `a0` (rows a1_0-1 .. a1_0+A1T, cols -1 .. W) of a [tokens, D] fp32 tensor into one ring slot
__device__ __forceinline__ void peg_load_plane_async(float* slot, const float* __restrict__ src, const PegGeom& g, int a0,
                                                     int a1_0) {
  const int a2h = g.W + 2;
  const int n_tok = (A1T + 2) * a2h;
  for (int idx = threadIdx.x; idx < n_tok * (CB / 4); idx += blockDim.x) {
    const int tok = idx / (CB / 4), q = idx % (CB / 4);
    const int r1 = tok / a2h, r2 = tok % a2h;
    const int a1 = a1_0 - 1 + r1, a2 = r2 - 1;
    const bool ok = a0 >= 0 && a0 < g.T && a1 >= 0 && a1 < g.H && a2 >= 0 && a2 < g.W;
    const float* p = ok ? src + peg_canon(g, a0, a1, a2) * g.D + q * 4 : src;
    peg_cp16(slot + (size_t)tok * CB + q * 4, p, ok);
  }
}
// async load of the upstream-gradient tile (A1T lines x W positions x CB channels) of plane a0
__device__ __forceinline__ void peg_load_dy_async(float* buf, const float* __restrict__ dy, const PegGeom& g, int a0, int a1_0) {
  const int n_tok = A1T * g.W;
  for (int idx = threadIdx.x; idx < n_tok * (CB / 4); idx += blockDim.x) {
    const int tok = idx / (CB / 4), q = idx % (CB / 4);
    const int a1 = a1_0 + tok / g.W, a2 = tok % g.W;
    const bool ok = a0 >= 0 && a0 < g.T && a1 < g.H;
    const float* p = ok ? dy + peg_canon(g, a0, a1, a2) * g.D + q * 4 : dy;
    peg_cp16(buf + (size_t)tok * CB + q * 4, p, ok);
  }
}

// One output position of the sliding window; the 3-wide window rotates through the register file (ROT) instead of
// being shifted, so no register moves are issued. Window slot (ROT+2)%3 receives the new right-hand element.
// Two independent accumulation chains keep the FMA pipe busy with only 8 warps per SM.
template <int ROT>
__device__ __forceinline__ float peg_step(float (&win)[9][3], const float* const (&rowp)[9], const float (&wt)[27],
                                          int off, float init) {
  constexpr int L = ROT % 3, C = (ROT + 1) % 3, R = (ROT + 2) % 3;
  float acc0 = init, acc1 = 0.f;
#pragma unroll
  for (int r = 0; r < 9; r++) win[r][R] = rowp[r][off];
#pragma unroll
  for (int r = 0; r < 9; r++) {
    float& acc = (r & 1) ? acc1 : acc0;
    acc = fmaf(wt[r * 3 + 0], win[r][L], acc);
    acc = fmaf(wt[r * 3 + 1], win[r][C], acc);
    acc = fmaf(wt[r * 3 + 2], win[r][R], acc);
  }
  return acc0 + acc1;
}
template <int ROT>
__device__ __forceinline__ void peg_wstep(float (&win)[9][3], const float* const (&rowp)[9], float (&acc)[27], int off,
                                          float d) {
  constexpr int L = ROT % 3, C = (ROT + 1) % 3, R = (ROT + 2) % 3;
#pragma unroll
  for (int r = 0; r < 9; r++) win[r][R] = rowp[r][off];
#pragma unroll
  for (int r = 0; r < 9; r++) {
    acc[r * 3 + 0] = fmaf(d, win[r][L], acc[r * 3 + 0]);
    acc[r * 3 + 1] = fmaf(d, win[r][C], acc[r * 3 + 1]);
    acc[r * 3 + 2] = fmaf(d, win[r][R], acc[r * 3 + 2]);
  }
}

// MODE 0: y = x + conv(x) + bias (forward; taps a0-2..a0)     MODE 1: dx = dy + conv^T(dy) (taps a0..a0+2, mirrored)
template <int MODE>
__global__ void __launch_bounds__(256, 1) peg_conv2_kernel(ctclip_peg_args a, int planes_per_cta) {
  extern __shared__ __align__(16) float peg_sm[];
  const PegGeom g{a.T, a.H, a.W, a.D, a.temporal, a.canon_table};
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
  // planes needed for output plane a0: a0-2, a0-1, a0 (MODE 0) / a0, a0+1, a0+2 (MODE 1). Walk a0 in direction `dirn`
  // so that only ONE new plane enters per step; it is fetched (cp.async) while the previous plane is being computed.
  const int first = (MODE == 0) ? p_begin : p_end - 1;
  const int dirn = (MODE == 0) ? 1 : -1;
  for (int d = 2; d >= 0; d--) {
    const int pl = first - dirn * d;
    peg_load_plane_async(peg_sm + (size_t)peg_slot(pl) * slot_elems, xin, g, pl, a1_0);
  }
  peg_commit();
  const int n_steps = p_end - p_begin;
  // canonical token index of every output of the current plane ([2][A1T][W], double-buffered by step parity): one
  // canon() per thread per plane instead of one per output inside the sliding loop
  int* s_tok = reinterpret_cast<int*>(peg_sm + NSLOT * slot_elems);
  for (int step = 0; step < n_steps; step++) {
    const int a0 = first + dirn * step;
    for (int i = threadIdx.x; i < A1T * a.W; i += blockDim.x) {
      const int l1 = a1_0 + i / a.W;
      s_tok[(step & 1) * A1T * a.W + i] = (l1 < a.H) ? (int)peg_canon(g, a0, l1, i % a.W) : 0;
    }
    peg_wait_all();
    __syncthreads();   // plane a0 landed for everyone; everyone finished computing plane a0 - dirn
    if (step + 1 < n_steps) {   // prefetch the next plane into the slot released by plane a0 - 3*dirn
      const int pl = a0 + dirn;
      peg_load_plane_async(peg_sm + (size_t)peg_slot(pl) * slot_elems, xin, g, pl, a1_0);
      peg_commit();
    }
    if (a1 < a.H) {
      const float* rowp[9];
#pragma unroll
      for (int r = 0; r < 9; r++) {
        const int k0 = r / 3, k1 = r % 3;
        const int pl = (MODE == 0) ? (a0 + k0 - 2) : (a0 + k0);
        rowp[r] = peg_sm + (size_t)peg_slot(pl) * slot_elems + (size_t)((line + k1) * a2h) * CB + lane;
      }
      float win[9][3];
#pragma unroll
      for (int r = 0; r < 9; r++) {
        win[r][0] = rowp[r][0];       // a2 = -1 (zero padding)
        win[r][1] = rowp[r][CB];      // a2 = 0
      }
      constexpr int CR = (MODE == 0) ? 7 : 1;   // row of the centre tap (k0 = 2 | 0, k1 = 1): its centre element = residual
      const int* tokl = s_tok + (step & 1) * A1T * a.W + line * a.W;
      auto emit = [&](int a2, float val) {
        const long long off = (long long)tokl[a2] * a.D + lane;
        yout[off] = val;
        if (ybf != nullptr) ybf[off] = __float2bfloat16(val);
      };
      for (int a2 = 0; a2 < a.W; a2 += 3) {
        const float o0 = peg_step<0>(win, rowp, wt, (a2 + 2) * CB, bias) + win[CR][1];
        emit(a2, o0);
        if (a2 + 1 < a.W) {
          const float o1 = peg_step<1>(win, rowp, wt, (a2 + 3) * CB, bias) + win[CR][2];
          emit(a2 + 1, o1);
        }
        if (a2 + 2 < a.W) {
          const float o2 = peg_step<2>(win, rowp, wt, (a2 + 4) * CB, bias) + win[CR][0];
          emit(a2 + 2, o2);
        }
      }
    }
  }
}

// dw[c][k] += sum_p dy[p] * x[p + off(k)], db[c] += sum_p dy[p]
// x planes in the 4-slot ring, the upstream-gradient tile of the same plane in a 2-slot ring, both prefetched with cp.async
__global__ void __launch_bounds__(256, 1) peg_wgrad2_kernel(ctclip_peg_args a, int planes_per_cta, int dy_double) {
  extern __shared__ __align__(16) float peg_sm[];
  const PegGeom g{a.T, a.H, a.W, a.D, a.temporal, a.canon_table};
  const int a2h = a.W + 2;
  const size_t slot_elems = (size_t)(A1T + 2) * a2h * CB;
  const size_t dy_elems = (size_t)A1T * a.W * CB;
  float* sdy = peg_sm + NSLOT * slot_elems;   // [2][A1T][W][CB]
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
  for (int d = 2; d >= 0; d--) {
    const int pl = p_begin - d;
    peg_load_plane_async(peg_sm + (size_t)peg_slot(pl) * slot_elems, xin, g, pl, a1_0);
  }
  peg_load_dy_async(sdy, dy, g, p_begin, a1_0);
  peg_commit();
  for (int a0 = p_begin; a0 < p_end; a0++) {
    const int par = dy_double ? ((a0 - p_begin) & 1) : 0;
    if (!dy_double && a0 > p_begin) {   // wide grids: a single gradient tile fits; fetch it without overlap
      __syncthreads();
      peg_load_dy_async(sdy, dy, g, a0, a1_0);
      peg_commit();
    }
    peg_wait_all();
    __syncthreads();
    if (a0 + 1 < p_end) {
      peg_load_plane_async(peg_sm + (size_t)peg_slot(a0 + 1) * slot_elems, xin, g, a0 + 1, a1_0);
      if (dy_double) peg_load_dy_async(sdy + (size_t)(par ^ 1) * dy_elems, dy, g, a0 + 1, a1_0);
      peg_commit();
    }
    if (a1 < a.H) {
      const float* rowp[9];
#pragma unroll
      for (int r = 0; r < 9; r++) {
        const int k0 = r / 3, k1 = r % 3;
        rowp[r] = peg_sm + (size_t)peg_slot(a0 + k0 - 2) * slot_elems + (size_t)((line + k1) * a2h) * CB + lane;
      }
      const float* dyl = sdy + (size_t)par * dy_elems + (size_t)(line * a.W) * CB + lane;
      float win[9][3];
#pragma unroll
      for (int r = 0; r < 9; r++) {
        win[r][0] = rowp[r][0];
        win[r][1] = rowp[r][CB];
      }
      for (int a2 = 0; a2 < a.W; a2 += 3) {
        const float d0 = dyl[a2 * CB];
        const float d1 = (a2 + 1 < a.W) ? dyl[(a2 + 1) * CB] : 0.f;
        const float d2 = (a2 + 2 < a.W) ? dyl[(a2 + 2) * CB] : 0.f;
        accb += d0 + d1 + d2;
        peg_wstep<0>(win, rowp, acc, (a2 + 2) * CB, d0);
        if (a2 + 1 < a.W) peg_wstep<1>(win, rowp, acc, (a2 + 3) * CB, d1);
        if (a2 + 2 < a.W) peg_wstep<2>(win, rowp, acc, (a2 + 4) * CB, d2);
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

// grid.x = channel blocks x a1 tiles x a0 chunks (1 CTA per SM: ~133-182 KB of shared memory). The a0 range is
// split into the chunk count that minimises  ceil(CTAs / SMs) * (planes per CTA + 2 priming planes).
static void peg_launch_shape(const ctclip_peg_args* a, int dy_bufs, dim3* grid, int* planes_per_cta, size_t* smem) {
  const int n_a1t = (a->H + A1T - 1) / A1T;
  const long long base = (long long)(a->D / CB) * n_a1t * a->B;
  int best_chunks = 1;
  long long best_cost = -1;
  for (int chunks = 1; chunks <= 6 && chunks <= a->T; chunks++) {
    const int ppc = (a->T + chunks - 1) / chunks;
    const int real_chunks = (a->T + ppc - 1) / ppc;
    const long long waves = (base * real_chunks + num_sms() - 1) / num_sms();
    const long long cost = waves * (ppc + 2);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_chunks = real_chunks; }
  }
  *planes_per_cta = (a->T + best_chunks - 1) / best_chunks;
  const int chunks = (a->T + *planes_per_cta - 1) / *planes_per_cta;
  *grid = dim3((unsigned)((a->D / CB) * n_a1t * chunks), (unsigned)a->B);
  *smem = sizeof(float) * NSLOT * (size_t)(A1T + 2) * (a->W + 2) * CB;
  if (dy_bufs == 0) *smem += sizeof(int) * 2 * (size_t)A1T * a->W;   // token-index table of the conv kernels
  *smem += sizeof(float) * dy_bufs * (size_t)A1T * a->W * CB;
  const size_t red_bytes = sizeof(float) * A1T * 28 * CB;  // weight-gradient reduction scratch
  if (*smem < red_bytes) *smem = red_bytes;
}

static int peg_launch_conv(int mode, const ctclip_peg_args* a, cudaStream_t stream) {
  dim3 grid;
  int ppc;
  size_t smem;
  peg_launch_shape(a, 0, &grid, &ppc, &smem);
  CTB_CHECK_ARG(smem <= 227 * 1024, "peg: token grid width %d needs %zu B of shared memory", a->W, smem);
  if (mode == 0) {
    CTB_CUDA(cudaFuncSetAttribute(peg_conv2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    peg_conv2_kernel<0><<<grid, 256, smem, stream>>>(*a, ppc);
  } else {
    CTB_CUDA(cudaFuncSetAttribute(peg_conv2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    peg_conv2_kernel<1><<<grid, 256, smem, stream>>>(*a, ppc);
  }
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}

static int peg_launch_wgrad(const ctclip_peg_args* a, cudaStream_t stream) {
  dim3 grid;
  int ppc;
  size_t smem;
  int dy_bufs = 2;
  peg_launch_shape(a, dy_bufs, &grid, &ppc, &smem);
  if (smem > 227 * 1024) {
    dy_bufs = 1;
    peg_launch_shape(a, dy_bufs, &grid, &ppc, &smem);
  }
  CTB_CHECK_ARG(smem <= 227 * 1024, "peg: token grid width %d needs %zu B of shared memory", a->W, smem);
  CTB_CUDA(cudaFuncSetAttribute(peg_wgrad2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  peg_wgrad2_kernel<<<grid, 256, smem, stream>>>(*a, ppc, dy_bufs == 2);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}

extern "C" int ctclip_peg_fwd(const ctclip_peg_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = peg_check(a, "peg_fwd")) return rc;
  CTB_CHECK_ARG(a->y && a->weight, "peg_fwd: null y/weight");
  return peg_launch_conv(0, a, stream);
}

// x = upstream gradient dy (fp32), y = dx out (fp32), y_bf16 = optional bf16 copy of dx
extern "C" int ctclip_peg_bwd_data(const ctclip_peg_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = peg_check(a, "peg_bwd_data")) return rc;
  CTB_CHECK_ARG(a->y && a->weight, "peg_bwd_data: null y/weight");
  return peg_launch_conv(1, a, stream);
}

// x = forward input, dy = upstream gradient; dweight [D,27] and dbias [D] are accumulated (atomics)
extern "C" int ctclip_peg_bwd_weight(const ctclip_peg_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = peg_check(a, "peg_bwd_weight")) return rc;
  CTB_CHECK_ARG(a->dy && a->dweight, "peg_bwd_weight: null dy/dweight");
  return peg_launch_wgrad(a, stream);
}
