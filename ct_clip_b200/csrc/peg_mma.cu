// PEG on the tensor pipe: the depthwise causal 3x3x3 convolution (attention.py:63-84) as bf16 MMAs with
// BLOCK-DIAGONAL weights, fp32 accumulation, fp32 residual.
//
// Why: the scalar / packed-fp32 stencil kernels (peg.cu) are issue-bound: 27 FMAs + 9 shared loads per output keep them at
// 3.7x the HBM time of the op however the loops are arranged (ncu: profiles/). A depthwise convolution has no reduction
// over channels, but for a group of 8 channels the 27 taps x 8 channels form the K dimension of
//     out[pos, c'] = sum_{tap, c} x[pos + off(tap), c] * Wbd[(tap, c), c'],      Wbd[(tap, c), c'] = w[c', tap] * (c == c')
// i.e. m16n8k16 MMAs whose A fragments are plain ldmatrix loads of the [token][channel] tile at a tap-shifted address
// (no im2col copy) and whose B fragments (two taps x 8 channels each, 7/8 zeros) live in registers. 15 ldmatrix + 15 HMMA
// produce 16 positions x 8 channels (128 outputs) instead of 3456 FMAs + 1152 shared loads.
// The weight gradient is the same contraction with positions as K:  dW[(tap, c), c'] = sum_pos x[pos + off, c] dy[pos, c'],
// of which the diagonal c == c' is kept; an extra all-ones "tap" row yields the bias gradient for free.
//
// STATUS: correct (tests/test_kernels_gpu.py::test_peg_fwd_bwd[mma]) but SLOWER than the packed-fp32 stencil of peg.cu on
// B200 (see peg_mma_supported below), so it is opt-in and the stencil stays the default path.
//
// Numerics: the convolution operands (x or dy, and w) are rounded to bf16, exactly what the reference's bf16 autocast does to
// nn.Conv3d (CTCLIPTrainer.py runs the model under accelerate's bf16 autocast); accumulation, bias and the residual
// `peg(x) + x` (attention.py:324) stay fp32.
//
// Tiling (shared with peg.cu): work unit = one plane-step of a COLUMN (volume, 8 conv-grid lines along a1, 32 channels);
// planes of the causal axis a0 roll through a 3-slot bf16 ring [token][40 bf16] (80-byte token pitch: conflict-free
// ldmatrix); fp32 planes arrive by cp.async into 3 staging buffers one step ahead and are converted after the compute
// phase. Output positions of a plane are addressed by the FLATTENED halo'd index o = line*(W+2) + col, so that every tap
// is a constant address offset o + k1*(W+2) + k2 and a plane is ceil(8*(W+2)/16) MMA row tiles (13 for W = 24); the two
// halo columns of every line produce garbage rows that are simply not stored. Persistent grid, equal contiguous ranges
// of (column, plane) steps per CTA.
#include <stdlib.h>
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/ctclip_b200.h"

namespace ctb {

constexpr int PM_A1T = 8;        // lines per column tile
constexpr int PM_CB = 32;        // channels per column (4 MMA channel groups of 8)
constexpr int PM_THREADS = 512;  // 16 warps: warp = (channel group = warp & 3, row-tile phase = warp >> 2)
constexpr int PM_WARPS = PM_THREADS / 32;
constexpr int PM_TOKB = 80;      // bytes per token in the bf16 ring (32 ch x 2 B + 16 B pad)
constexpr int PM_NSLOT = 3;      // bf16 ring: planes a0-2, a0-1, a0
constexpr int PM_NSTG = 4;       // fp32 staging buffers of the conv kernels: the plane being computed (its centre tap is the
                                 // fp32 residual) + three planes in flight (the step is far shorter than the DRAM latency)
constexpr int PM_NSTG_W = 2;     // weight-gradient kernel: two x planes + two dy tiles in flight
constexpr int PM_MAXIT = 7;      // loader items per thread and plane: (A1T+2)(W+2)*8/512 <= 7  <=>  W <= 42
constexpr int PM_PADTOK = 18;    // tokens past the last halo row that shifted ldmatrix rows may touch

struct PmGeom {
  int T, H, W, D, temporal;
  const int* table;
  int a2h;    // W + 2
  int n_tok;  // (A1T + 2) * a2h tokens of one halo'd plane tile
  int n_mt;   // MMA row tiles per plane = ceil(A1T * a2h / 16)
};

__device__ __forceinline__ int pm_canon(const PmGeom& g, int f) {
  if (!g.temporal) return f;
  if (g.table != nullptr) return __ldg(g.table + f);
  const int it = f % g.T, iw = (f / g.T) % g.W, ih = f / (g.T * g.W);
  return (it * g.H + ih) * g.W + iw;
}
__device__ __forceinline__ void pm_ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void pm_ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void pm_ldsm_x2_trans(uint32_t (&r)[2], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void pm_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void pm_cp16(uint32_t dst, const void* src, bool valid) {
  const int bytes = valid ? 16 : 0;   // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pm_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void pm_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void pm_wait_pending() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct PmCol {
  int b, a1_0, c0;
};
__device__ __forceinline__ PmCol pm_column(int col, int n_cb, int n_a1t) {
  PmCol c;
  c.c0 = (col % n_cb) * PM_CB;
  c.a1_0 = ((col / n_cb) % n_a1t) * PM_A1T;
  c.b = col / (n_cb * n_a1t);
  return c;
}
// loader table of the current column: item idx = tid + it*512 = (halo token idx/8, 16-byte quad idx%8);
// inpl[it] = in-plane conv-grid offset a1*W + a2, -1 = zero-filled halo, -2 = no such item
__device__ __forceinline__ void pm_loader_setup(int (&inpl)[PM_MAXIT], const PmGeom& g, int a1_0) {
#pragma unroll
  for (int it = 0; it < PM_MAXIT; it++) {
    const int idx = threadIdx.x + it * PM_THREADS;
    int v = -2;
    if (idx < g.n_tok * 8) {
      const int tok = idx >> 3;
      const int r1 = tok / g.a2h, r2 = tok - r1 * g.a2h;
      const int a1 = a1_0 - 1 + r1, a2 = r2 - 1;
      v = (a1 >= 0 && a1 < g.H && a2 >= 0 && a2 < g.W) ? a1 * g.W + a2 : -1;
    }
    inpl[it] = v;
  }
}
// cp.async plane a0 of a [tokens, D] fp32 tensor (src already offset to the column's first channel) into a staging buffer
__device__ __forceinline__ void pm_issue_plane(uint32_t stg, const float* __restrict__ src, const PmGeom& g, int a0,
                                               const int (&inpl)[PM_MAXIT]) {
  const bool pl_ok = a0 >= 0 && a0 < g.T;
  const int fbase = a0 * g.H * g.W;
  const float* s4 = src + (threadIdx.x & 7) * 4;
#pragma unroll
  for (int it = 0; it < PM_MAXIT; it++) {
    const int v = inpl[it];
    if (v != -2) {
      const bool ok = pl_ok && v >= 0;
      const float* p = s4;
      if (ok) p = s4 + (long long)pm_canon(g, fbase + v) * g.D;
      pm_cp16(stg + (threadIdx.x + it * PM_THREADS) * 16, p, ok);
    }
  }
}
// staging [token][32 fp32] -> ring slot [token][40 bf16]
__device__ __forceinline__ void pm_convert_plane(const float* stg, uint8_t* slot, const int (&inpl)[PM_MAXIT]) {
#pragma unroll
  for (int it = 0; it < PM_MAXIT; it++) {
    if (inpl[it] != -2) {
      const int idx = threadIdx.x + it * PM_THREADS;
      const float4 v = *reinterpret_cast<const float4*>(stg + (size_t)idx * 4);
      uint2 u;
      u.x = pack_bf16x2(v.x, v.y);
      u.y = pack_bf16x2(v.z, v.w);
      *reinterpret_cast<uint2*>(slot + (size_t)(idx >> 3) * PM_TOKB + (idx & 7) * 8) = u;
    }
  }
}
// canonical token of every flattened output index o of plane a0 (-1: halo column / line beyond H)
__device__ __forceinline__ void pm_write_tokens(int* s_tok, const PmGeom& g, int a0, int a1_0) {
  for (int o = threadIdx.x; o < g.n_mt * 16; o += PM_THREADS) {
    const int l = o / g.a2h, c = o - l * g.a2h;
    int tk = -1;
    if (l < PM_A1T && c < g.W && a1_0 + l < g.H) tk = pm_canon(g, (a0 * g.H + a1_0 + l) * g.W + c);
    s_tok[o] = tk;
  }
}

struct PmSmem {
  uint8_t* ring;     // [PM_NSLOT][slot_bytes]
  float* stg;        // [PM_NSTG][n_tok * 32]
  int* s_tok;        // [2][n_mt * 16]
  uint32_t slot_bytes;
};
__device__ __forceinline__ PmSmem pm_carve(uint8_t* base, const PmGeom& g, int nstg) {
  PmSmem s;
  s.slot_bytes = (uint32_t)(g.n_tok + PM_PADTOK) * PM_TOKB;
  s.ring = base;
  s.stg = reinterpret_cast<float*>(base + (size_t)PM_NSLOT * s.slot_bytes);
  s.s_tok = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(s.stg) + (size_t)nstg * g.n_tok * 128);
  return s;
}
__device__ __forceinline__ int pm_slot(int pl) { return ((pl % PM_NSLOT) + PM_NSLOT) % PM_NSLOT; }

// MODE 0: y = x + conv(x) + bias (taps a0-2..a0)      MODE 1: dx = dy + conv^T(dy) (taps a0..a0+2, mirrored weights)
// dbg (probing only, CTCLIP_PEG_DEBUG): bit 0 = no plane loads, bit 1 = no MMAs / output stores, bit 2 = no output stores,
// bit 3 = no fp32 -> bf16 conversion
template <int MODE>
__global__ void __launch_bounds__(PM_THREADS, 1) peg_mma_conv_kernel(ctclip_peg_args a, int steps_per_cta, int dbg) {
  extern __shared__ __align__(128) uint8_t pm_sm[];
  PmGeom g{a.T, a.H, a.W, a.D, a.temporal, a.canon_table, a.W + 2, (PM_A1T + 2) * (a.W + 2), (PM_A1T * (a.W + 2) + 15) / 16};
  const PmSmem sm = pm_carve(pm_sm, g, PM_NSTG);
  // TEAM scheduling: the n_cb channel blocks of one (volume, line tile, plane) are processed by n_cb CTAs (a team) at
  // about the same time, so that the 128-byte channel slices of a 2 KB token row are requested while its DRAM page is
  // open (one CTA per column at its own pace left every 128-byte read on a closed page: all kernel variants plateaued
  // at 1.4-1.8 TB/s). Team = blockIdx.x / n_cb, channel block = blockIdx.x % n_cb; a team owns a contiguous range of
  // (volume, line tile, plane) steps.
  const int n_a1t = (a.H + PM_A1T - 1) / PM_A1T, n_cb = a.D / PM_CB;
  const int total = n_a1t * a.B * a.T;
  const int team = blockIdx.x / n_cb, my_cb = blockIdx.x % n_cb;
  const int s_begin = team * steps_per_cta;
  const int s_end = min(total, s_begin + steps_per_cta);
  const long long vol = (long long)a.T * a.H * a.W * a.D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cg = warp & 3, mq = warp >> 2;
  const int gq = lane >> 2, t = lane & 3;
  // the pad tokens of every ring slot are never written by the converter: zero them once
  for (int i = threadIdx.x; i < PM_NSLOT * PM_PADTOK * (PM_TOKB / 16); i += PM_THREADS) {
    const int sl = i / (PM_PADTOK * (PM_TOKB / 16)), r = i % (PM_PADTOK * (PM_TOKB / 16));
    *reinterpret_cast<uint4*>(sm.ring + (size_t)sl * sm.slot_bytes + (size_t)g.n_tok * PM_TOKB + r * 16) = make_uint4(0, 0, 0, 0);
  }
  // per-lane A-fragment address offsets of the 5 k-steps of a plane: lanes 0-15 address tap 2j, lanes 16-31 tap 2j+1
  // (tap 9 = padding: zero weight, any valid address)
  uint32_t aoff[5];
#pragma unroll
  for (int j = 0; j < 5; j++) {
    const int tp = min(2 * j + (lane >> 4), 8);
    aoff[j] = (uint32_t)((tp / 3) * g.a2h + (tp % 3)) * PM_TOKB;
  }
  const uint32_t lane_off = (uint32_t)(lane & 15) * PM_TOKB + cg * 16;
  const uint32_t ring_u32 = smem_u32(sm.ring);
  const uint32_t stg_u32 = smem_u32(sm.stg);
  int inpl[PM_MAXIT];
  int par = 0;
  for (int s = s_begin; s < s_end;) {
    const int col = s / a.T;
    const int p_begin = s - col * a.T;
    const int p_end = min(a.T, p_begin + (s_end - s));
    const PmCol cc = pm_column(col * n_cb + my_cb, n_cb, n_a1t);
    const float* xin = a.x + (long long)cc.b * vol + cc.c0;
    const int chp = cc.c0 + cg * 8 + 2 * t;   // this lane's output channel pair
    float* yout = a.y + (long long)cc.b * vol + chp;
    __nv_bfloat16* ybf = a.y_bf16 ? reinterpret_cast<__nv_bfloat16*>(a.y_bf16) + (long long)cc.b * vol + chp : nullptr;
    // block-diagonal weight fragments: B[k = (tap sel, ch_in)][n = ch_out = gq]; this lane holds k = 2t, 2t+1 (+8)
    uint32_t bfr[15][2];
    {
      const int ch = cc.c0 + cg * 8 + gq;
#pragma unroll
      for (int k0 = 0; k0 < 3; k0++)
#pragma unroll
        for (int j = 0; j < 5; j++) {
          float wA = 0.f, wB = 0.f;
          const int kA = k0 * 9 + 2 * j, kB = kA + 1;
          wA = __ldg(a.weight + (long long)ch * 27 + ((MODE == 0) ? kA : 26 - kA));
          if (2 * j + 1 < 9) wB = __ldg(a.weight + (long long)ch * 27 + ((MODE == 0) ? kB : 26 - kB));
          bfr[k0 * 5 + j][0] = pack_bf16x2((2 * t == gq) ? wA : 0.f, (2 * t + 1 == gq) ? wA : 0.f);
          bfr[k0 * 5 + j][1] = pack_bf16x2((2 * t == gq) ? wB : 0.f, (2 * t + 1 == gq) ? wB : 0.f);
        }
    }
    float2 bias2 = make_float2(0.f, 0.f);
    if (MODE == 0 && a.bias != nullptr) bias2 = make_float2(__ldg(a.bias + chp), __ldg(a.bias + chp + 1));
    __syncthreads();   // previous column completely consumed (ring, staging, token tables)
    pm_loader_setup(inpl, g, cc.a1_0);
    const int first = (MODE == 0) ? p_begin : p_end - 1;
    const int dirn = (MODE == 0) ? 1 : -1;
    // staging buffer of the plane that enters at step i (i = -2, -1: the two extra planes of the first output): i mod 4
    // prime: the three planes of the first output plane in one round trip
#pragma unroll
    for (int d = 0; d < 3; d++)
      pm_issue_plane(stg_u32 + ((PM_NSTG - d) % PM_NSTG) * g.n_tok * 128, xin, g, first - dirn * d, inpl);
    pm_commit();
    pm_write_tokens(sm.s_tok + par * g.n_mt * 16, g, first, cc.a1_0);
    pm_wait_all();
    __syncthreads();
#pragma unroll
    for (int d = 0; d < 3; d++)
      pm_convert_plane(sm.stg + (size_t)((PM_NSTG - d) % PM_NSTG) * g.n_tok * 32,
                       sm.ring + (size_t)pm_slot(first - dirn * d) * sm.slot_bytes, inpl);
    __syncthreads();
    const int n_steps = p_end - p_begin;
    // three planes ahead, one cp.async group per plane (empty groups keep the group arithmetic uniform near the end)
#pragma unroll
    for (int d = 1; d <= 3; d++) {
      if (d < n_steps) pm_issue_plane(stg_u32 + (d % PM_NSTG) * g.n_tok * 128, xin, g, first + dirn * d, inpl);
      pm_commit();
    }
    for (int step = 0; step < n_steps; step++, par ^= 1) {
      const int a0 = first + dirn * step;
      const bool more = step + 1 < n_steps;
      if (step >= 1) {   // the fp32 copy of the previous plane is dead (its residual is written): refill it, 3 planes ahead
        if (step + 3 < n_steps && !(dbg & 1))
          pm_issue_plane(stg_u32 + ((step + 3) % PM_NSTG) * g.n_tok * 128, xin, g, a0 + 3 * dirn, inpl);
        pm_commit();
      }
      // fp32 residual = centre tap of the plane that entered at this step: read from its staging buffer (halo token
      // index of output o is o + a2h + 1), not from global memory (a dependent L2 round trip per row tile otherwise)
      const float* res_s = sm.stg + (size_t)(step % PM_NSTG) * g.n_tok * 32 + (size_t)(g.a2h + 1) * 32 + cg * 8 + 2 * t;
      // ---- compute plane a0
      uint32_t pb[3];
#pragma unroll
      for (int k0 = 0; k0 < 3; k0++) {
        const int pl = (MODE == 0) ? (a0 + k0 - 2) : (a0 + k0);
        pb[k0] = ring_u32 + pm_slot(pl) * sm.slot_bytes + lane_off;
      }
      const int* tokp = sm.s_tok + par * g.n_mt * 16;
      for (int mt = mq; mt < g.n_mt && !(dbg & 2); mt += PM_WARPS / 4) {
        const int o0 = mt * 16;
        int tka = tokp[o0 + gq], tkb = tokp[o0 + gq + 8];
        if (dbg & 4) tka = tkb = -1;
        const float2 ra = *reinterpret_cast<const float2*>(res_s + (size_t)(o0 + gq) * 32);
        const float2 rb = *reinterpret_cast<const float2*>(res_s + (size_t)(o0 + gq + 8) * 32);
        float acc[4] = {bias2.x, bias2.y, bias2.x, bias2.y};
#pragma unroll
        for (int k0 = 0; k0 < 3; k0++) {
          const uint32_t rbase = pb[k0] + o0 * PM_TOKB;
#pragma unroll
          for (int j = 0; j < 5; j++) {
            uint32_t af[4];
            pm_ldsm_x4(af, rbase + aoff[j]);
            pm_mma(acc, af, bfr[k0 * 5 + j][0], bfr[k0 * 5 + j][1]);
          }
        }
        if (tka >= 0) {
          const float2 v = make_float2(acc[0] + ra.x, acc[1] + ra.y);
          *reinterpret_cast<float2*>(yout + (long long)tka * a.D) = v;
          if (ybf != nullptr) *reinterpret_cast<uint32_t*>(ybf + (long long)tka * a.D) = pack_bf16x2(v.x, v.y);
        }
        if (tkb >= 0) {
          const float2 v = make_float2(acc[2] + rb.x, acc[3] + rb.y);
          *reinterpret_cast<float2*>(yout + (long long)tkb * a.D) = v;
          if (ybf != nullptr) *reinterpret_cast<uint32_t*>(ybf + (long long)tkb * a.D) = pack_bf16x2(v.x, v.y);
        }
      }
      if (more) {
        pm_wait_pending<2>();   // the plane of the next step has landed (the two younger groups may still be in flight)
        __syncthreads();        // everyone finished reading plane a0 - 2*dirn (its slot is refilled now)
        if (!(dbg & 8))
          pm_convert_plane(sm.stg + (size_t)((step + 1) % PM_NSTG) * g.n_tok * 32,
                           sm.ring + (size_t)pm_slot(a0 + dirn) * sm.slot_bytes, inpl);
        pm_write_tokens(sm.s_tok + (par ^ 1) * g.n_mt * 16, g, a0 + dirn, cc.a1_0);
        __syncthreads();
      }
    }
    s += n_steps;
  }
}

// dw[c][tap] += sum_pos x[pos + off(tap)][c] * dy[pos][c],  db[c] += sum_pos dy[pos][c]
__global__ void __launch_bounds__(PM_THREADS, 1) peg_mma_wgrad_kernel(ctclip_peg_args a, int steps_per_cta) {
  extern __shared__ __align__(128) uint8_t pm_sm[];
  PmGeom g{a.T, a.H, a.W, a.D, a.temporal, a.canon_table, a.W + 2, (PM_A1T + 2) * (a.W + 2), (PM_A1T * (a.W + 2) + 15) / 16};
  const PmSmem sm = pm_carve(pm_sm, g, PM_NSTG_W);
  // after the conv kernel's regions: dy tile bf16 [n_mt*16 tokens][80 B], two dy staging tiles fp32 [A1T*W][32] (contiguous:
  // together they also receive the third plane when a column is primed), 128 B of bf16 ones
  uint8_t* sdy = reinterpret_cast<uint8_t*>(sm.s_tok + 2 * g.n_mt * 16);
  float* dy_stg = reinterpret_cast<float*>(sdy + (size_t)g.n_mt * 16 * PM_TOKB);
  const int dy_elems = PM_A1T * g.W * 32;
  uint8_t* ones = reinterpret_cast<uint8_t*>(dy_stg + (size_t)2 * dy_elems);
  // TEAM scheduling: the n_cb channel blocks of one (volume, line tile, plane) are processed by n_cb CTAs (a team) at
  // about the same time, so that the 128-byte channel slices of a 2 KB token row are requested while its DRAM page is
  // open (one CTA per column at its own pace left every 128-byte read on a closed page: all kernel variants plateaued
  // at 1.4-1.8 TB/s). Team = blockIdx.x / n_cb, channel block = blockIdx.x % n_cb; a team owns a contiguous range of
  // (volume, line tile, plane) steps.
  const int n_a1t = (a.H + PM_A1T - 1) / PM_A1T, n_cb = a.D / PM_CB;
  const int total = n_a1t * a.B * a.T;
  const int team = blockIdx.x / n_cb, my_cb = blockIdx.x % n_cb;
  const int s_begin = team * steps_per_cta;
  const int s_end = min(total, s_begin + steps_per_cta);
  const long long vol = (long long)a.T * a.H * a.W * a.D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cg = warp & 3, mq = warp >> 2;
  const int gq = lane >> 2, t = lane & 3;
  for (int i = threadIdx.x; i < PM_NSLOT * PM_PADTOK * (PM_TOKB / 16); i += PM_THREADS) {
    const int sl = i / (PM_PADTOK * (PM_TOKB / 16)), r = i % (PM_PADTOK * (PM_TOKB / 16));
    *reinterpret_cast<uint4*>(sm.ring + (size_t)sl * sm.slot_bytes + (size_t)g.n_tok * PM_TOKB + r * 16) = make_uint4(0, 0, 0, 0);
  }
  for (int i = threadIdx.x; i < g.n_mt * 16 * (PM_TOKB / 16); i += PM_THREADS)   // halo columns of the dy tile stay zero
    reinterpret_cast<uint4*>(sdy)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x < 8) reinterpret_cast<uint4*>(ones)[threadIdx.x] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
  // per-lane A^T fragment offsets: matrix (lane >> 3): bit 0 = tap 2j / 2j+1, bit 1 = positions 0-7 / 8-15; row = lane & 7
  uint32_t aoff[5];
#pragma unroll
  for (int j = 0; j < 5; j++) {
    const int tp = min(2 * j + ((lane >> 3) & 1), 8);
    aoff[j] = (uint32_t)((tp / 3) * g.a2h + (tp % 3) + (lane >> 4) * 8 + (lane & 7)) * PM_TOKB + cg * 16;
  }
  const bool ones_lane = ((lane >> 3) & 1) != 0;   // lanes that address the padding tap of k-step 4 -> all-ones rows (bias grad)
  const uint32_t ones_u32 = smem_u32(ones);
  const uint32_t ring_u32 = smem_u32(sm.ring);
  const uint32_t stg_u32 = smem_u32(sm.stg);
  const uint32_t sdy_u32 = smem_u32(sdy) + (uint32_t)(lane & 15) * PM_TOKB + cg * 16;
  const uint32_t dystg_u32 = smem_u32(dy_stg);
  int inpl[PM_MAXIT];
  for (int s = s_begin; s < s_end;) {
    const int col = s / a.T;
    const int p_begin = s - col * a.T;
    const int p_end = min(a.T, p_begin + (s_end - s));
    const PmCol cc = pm_column(col * n_cb + my_cb, n_cb, n_a1t);
    const float* xin = a.x + (long long)cc.b * vol + cc.c0;
    const float* dyin = a.dy + (long long)cc.b * vol + cc.c0;
    float acc[15][4];
#pragma unroll
    for (int i = 0; i < 15; i++) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    __syncthreads();
    pm_loader_setup(inpl, g, cc.a1_0);
    // upstream-gradient tile loader: A1T lines x W tokens x 8 quads, fp32 -> staging tile `buf`
    auto issue_dy = [&](int a0, int buf) {
      const int n_valid = min(PM_A1T, g.H - cc.a1_0) * g.W;
      for (int idx = threadIdx.x; idx < PM_A1T * g.W * 8; idx += PM_THREADS) {
        const int tok = idx >> 3;
        const bool ok = tok < n_valid;
        const float* p = dyin + (idx & 7) * 4;
        if (ok) p += (long long)pm_canon(g, (a0 * g.H + cc.a1_0) * g.W + tok) * g.D;
        pm_cp16(dystg_u32 + buf * dy_elems * 4 + idx * 16, p, ok);
      }
    };
    auto convert_dy = [&](int buf) {
      const float* src = dy_stg + (size_t)buf * dy_elems;
      for (int idx = threadIdx.x; idx < PM_A1T * g.W * 8; idx += PM_THREADS) {
        const int tok = idx >> 3;
        const int l = tok / g.W, c = tok - l * g.W;
        const float4 v = *reinterpret_cast<const float4*>(src + (size_t)idx * 4);
        uint2 u;
        u.x = pack_bf16x2(v.x, v.y);
        u.y = pack_bf16x2(v.z, v.w);
        *reinterpret_cast<uint2*>(sdy + (size_t)(l * g.a2h + c) * PM_TOKB + (idx & 7) * 8) = u;
      }
    };
    const int n_steps = p_end - p_begin;
    // prime the ring in one round trip: planes p_begin-2, p_begin-1 -> the two x staging buffers, plane p_begin -> the
    // (contiguous, still unused) dy staging tiles
    pm_issue_plane(stg_u32, xin, g, p_begin - 2, inpl);
    pm_issue_plane(stg_u32 + g.n_tok * 128, xin, g, p_begin - 1, inpl);
    pm_issue_plane(dystg_u32, xin, g, p_begin, inpl);
    pm_commit();
    pm_wait_all();
    __syncthreads();
    pm_convert_plane(sm.stg, sm.ring + (size_t)pm_slot(p_begin - 2) * sm.slot_bytes, inpl);
    pm_convert_plane(sm.stg + (size_t)g.n_tok * 32, sm.ring + (size_t)pm_slot(p_begin - 1) * sm.slot_bytes, inpl);
    pm_convert_plane(dy_stg, sm.ring + (size_t)pm_slot(p_begin) * sm.slot_bytes, inpl);
    __syncthreads();
    issue_dy(p_begin, 0);
    pm_commit();
    pm_wait_all();
    __syncthreads();
    convert_dy(0);
    __syncthreads();
    // two planes ahead: plane q = p_begin + i uses x staging buffer i % 2 and dy staging tile i % 2, one group per plane
#pragma unroll
    for (int d = 1; d <= 2; d++) {
      if (d < n_steps) {
        pm_issue_plane(stg_u32 + (d % PM_NSTG_W) * g.n_tok * 128, xin, g, p_begin + d, inpl);
        issue_dy(p_begin + d, d % 2);
      }
      pm_commit();
    }
    for (int i = 0; i < n_steps; i++) {
      const int a0 = p_begin + i;
      const bool more = i + 1 < n_steps;
      if (i >= 1) {   // buffers i % 2 were converted at the end of the previous step: refill them with plane a0 + 2
        if (i + 2 < n_steps) {
          pm_issue_plane(stg_u32 + (i % PM_NSTG_W) * g.n_tok * 128, xin, g, a0 + 2, inpl);
          issue_dy(a0 + 2, i % 2);
        }
        pm_commit();
      }
      uint32_t pb[3];
#pragma unroll
      for (int k0 = 0; k0 < 3; k0++) pb[k0] = ring_u32 + pm_slot(a0 + k0 - 2) * sm.slot_bytes;
      for (int mt = mq; mt < g.n_mt; mt += PM_WARPS / 4) {
        const uint32_t o0b = (uint32_t)mt * 16 * PM_TOKB;
        uint32_t bfr[2];
        pm_ldsm_x2_trans(bfr, sdy_u32 + o0b);
#pragma unroll
        for (int k0 = 0; k0 < 3; k0++) {
#pragma unroll
          for (int j = 0; j < 5; j++) {
            uint32_t af[4];
            uint32_t addr = pb[k0] + o0b + aoff[j];
            if (j == 4 && ones_lane) addr = ones_u32 + (lane & 7) * 16;
            pm_ldsm_x4_trans(af, addr);
            pm_mma(acc[k0 * 5 + j], af, bfr[0], bfr[1]);
          }
        }
      }
      if (more) {
        pm_wait_pending<1>();   // plane a0 + 1 (x and dy) landed; plane a0 + 2 may still be in flight
        __syncthreads();
        pm_convert_plane(sm.stg + (size_t)((i + 1) % PM_NSTG_W) * g.n_tok * 32,
                         sm.ring + (size_t)pm_slot(a0 + 1) * sm.slot_bytes, inpl);
        convert_dy((i + 1) % 2);
        __syncthreads();
      }
    }
    // flush: rows (gq: tap 2j, gq+8: tap 2j+1) x cols (2t, 2t+1) -- the diagonal (channel gq == column) lives in the lanes
    // with t == gq >> 1, element gq & 1
    if (t == (gq >> 1)) {
      const int e = gq & 1;
      const int ch = cc.c0 + cg * 8 + gq;
#pragma unroll
      for (int k0 = 0; k0 < 3; k0++)
#pragma unroll
        for (int j = 0; j < 5; j++) {
          const int kA = k0 * 9 + 2 * j;
          const float vA = e ? acc[k0 * 5 + j][1] : acc[k0 * 5 + j][0];   // (no dynamic register-array indexing)
          const float vB = e ? acc[k0 * 5 + j][3] : acc[k0 * 5 + j][2];
          atomicAdd(a.dweight + (long long)ch * 27 + kA, vA);
          if (2 * j + 1 < 9) atomicAdd(a.dweight + (long long)ch * 27 + kA + 1, vB);
          else if (k0 == 2 && a.dbias != nullptr) atomicAdd(a.dbias + ch, vB);
        }
    }
    s += p_end - p_begin;
  }
}

static size_t pm_smem_bytes(const ctclip_peg_args* a, bool wgrad) {
  const int a2h = a->W + 2;
  const size_t n_tok = (size_t)(PM_A1T + 2) * a2h;
  const size_t n_mt = (PM_A1T * (size_t)a2h + 15) / 16;
  size_t b = (size_t)PM_NSLOT * (n_tok + PM_PADTOK) * PM_TOKB + (size_t)(wgrad ? PM_NSTG_W : PM_NSTG) * n_tok * 128 +
             2 * n_mt * 16 * sizeof(int);
  if (wgrad) b += n_mt * 16 * PM_TOKB + (size_t)2 * PM_A1T * a->W * 128 + 128;
  return b + 128;
}

// the tensor-core path needs D % 32 == 0, W <= 42 and the tiles above in <= 227 KB of shared memory
bool peg_mma_supported(const ctclip_peg_args* a, bool wgrad) {
  // Opt-in (ctclip_peg_args.lines == -2 or CTCLIP_PEG_MMA=1). Measured on B200 at configs[1] (tools/peg_probe.py, profiles/):
  // 305 us per forward launch against 238 us for the packed-fp32 stencil. The formulation re-reads every input element 27
  // times through ldmatrix (once per tap): 3.7 GB of shared-memory traffic per launch = 100 us at 128 B/clk/SM, plus a
  // 15-deep dependent HMMA chain per row tile; removing loads, conversion or stores changes little (probe knobs below).
  static const int on = getenv("CTCLIP_PEG_MMA") ? atoi(getenv("CTCLIP_PEG_MMA")) : 0;
  if (!(on || a->lines == -2)) return false;
  if (wgrad && 2 * PM_A1T * a->W < (PM_A1T + 2) * (a->W + 2)) return false;   // the dy staging tiles must hold one x plane (W >= 4)
  return a->D % PM_CB == 0 && a->W <= 42 && pm_smem_bytes(a, wgrad) <= 227 * 1024 &&
         (long long)a->B * a->T * a->H * a->W < (1ll << 31) / 64;
}

static void pm_launch_shape(const ctclip_peg_args* a, int* grid, int* steps_per_team) {
  const int n_a1t = (a->H + PM_A1T - 1) / PM_A1T;
  const int n_cb = a->D / PM_CB;
  const long long total = (long long)n_a1t * a->B * a->T;      // (volume, line tile, plane) steps, each done by n_cb CTAs
  long long teams = num_sms() / n_cb;
  if (teams < 1) teams = 1;
  if (teams > total) teams = total;
  *steps_per_team = (int)((total + teams - 1) / teams);
  teams = (total + *steps_per_team - 1) / *steps_per_team;
  *grid = (int)(teams * n_cb);
}

int peg_mma_launch_conv(int mode, const ctclip_peg_args* a, cudaStream_t stream) {
  int grid, spc;
  pm_launch_shape(a, &grid, &spc);
  const size_t smem = pm_smem_bytes(a, false);
  const int dbg = getenv("CTCLIP_PEG_DEBUG") ? atoi(getenv("CTCLIP_PEG_DEBUG")) : 0;
  if (mode == 0) {
    CTB_CUDA(cudaFuncSetAttribute(peg_mma_conv_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    peg_mma_conv_kernel<0><<<grid, PM_THREADS, smem, stream>>>(*a, spc, dbg);
  } else {
    CTB_CUDA(cudaFuncSetAttribute(peg_mma_conv_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    peg_mma_conv_kernel<1><<<grid, PM_THREADS, smem, stream>>>(*a, spc, dbg);
  }
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}

int peg_mma_launch_wgrad(const ctclip_peg_args* a, cudaStream_t stream) {
  int grid, spc;
  pm_launch_shape(a, &grid, &spc);
  const size_t smem = pm_smem_bytes(a, true);
  CTB_CUDA(cudaFuncSetAttribute(peg_mma_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  peg_mma_wgrad_kernel<<<grid, PM_THREADS, smem, stream>>>(*a, spc);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}

}  // namespace ctb
