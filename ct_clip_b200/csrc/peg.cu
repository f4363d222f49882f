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
// See the v4 kernel comment below for the tiling (rolling plane ring in shared memory along the causal axis).
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/ctclip_b200.h"

namespace ctb {

struct PegGeom {
  int T, H, W, D, temporal;
  const int* table;   // optional precomputed canon(f) (avoids three integer divisions per access in the temporal stack)
};

// canonical token of conv-grid flat index f of the temporal stack, without the lookup table
__device__ __forceinline__ long long peg_canon_f(const PegGeom& g, int f) {
  const int it = f % g.T;
  const int iw = (f / g.T) % g.W;
  const int ih = f / (g.T * g.W);
  return ((long long)it * g.H + ih) * g.W + iw;
}
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
// v4 kernels. Work unit = one plane-step of a COLUMN = (volume, tile of A1T=8 conv-grid lines along a1, 32-channel
// block); a column is walked along the causal axis a0 with a ROLLING ring of planes in shared memory
// ([slot][a1 halo][a2 halo][32 ch] fp32): every step loads ONE new plane with cp.async (halo factor
// (A1T+2)(W+2)/(A1T*W) ~ 1.35x of compulsory traffic, 128 B coalesced segments) while the previous one is computed.
// Each thread (= one line x one PAIR of channels x one half of the a2 range) slides a register window along a2 with
// packed fp32x2 arithmetic: 18 LDS.64 + 54 FFMA2 per two outputs x two channels (the scalar v2-v4 kernels were
// issue-bound at 80-95 instructions per output and 48 % issue utilisation with 8 warps per SM; ncu: profiles/). The
// window holds FOUR register columns, two outputs are computed per load group (four independent FFMA2 chains), and
// all shared-memory addresses are [running 32-bit offset + immediate]; the plane loader's (line, column) decomposition
// is done once per column and kept in registers.
// The grid is PERSISTENT: all (column, plane) steps are linearised and cut into gridDim.x equal contiguous ranges, so
// every SM gets the same number of plane-steps (the former one-CTA-per-column grid ran 2.6 waves = 3 wave times).
// ------------------------------------------------------------------------------------------------
constexpr int A1T = 8;          // lines per column tile
constexpr int CB = 32;          // channels per column
constexpr int NSLOT = 4;        // ring of planes: 3 live + 1 being filled by cp.async while the current plane is computed
constexpr int PEG_THREADS = 256;
constexpr int PEG_MAXIT = 14;   // cp.async items per thread and plane: (A1T+2)(W+2)(CB/4)/256 <= 14  <=>  W <= 42

__device__ __forceinline__ int peg_slot(int pl) { return ((pl % NSLOT) + NSLOT) % NSLOT; }

__device__ __forceinline__ void peg_cp16(void* smem_dst, const void* gmem_src, bool valid) {
  const int bytes = valid ? 16 : 0;   // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void peg_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void peg_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

struct PegCol {
  int b, a1_0, c0;
};
__device__ __forceinline__ PegCol peg_column(int col, int n_cb, int n_a1t) {
  PegCol c;
  c.c0 = (col % n_cb) * CB;
  c.a1_0 = ((col / n_cb) % n_a1t) * A1T;
  c.b = col / (n_cb * n_a1t);
  return c;
}

// Per-thread loader table of the current column: item idx = tid + it*256 copies 16 bytes (4 channels) of halo token
// idx/8 = (r1, r2); inpl[it] = in-plane conv-grid offset a1*W + a2 of that token, -1 for zero-filled halo, -2 for
// "beyond the slot".
__device__ __forceinline__ void peg_loader_setup(int (&inpl)[PEG_MAXIT], int nit, int n_items, int a2h, int a1_0, int H,
                                                 int W) {
#pragma unroll
  for (int it = 0; it < PEG_MAXIT; it++) {
    const int idx = threadIdx.x + it * PEG_THREADS;
    int v = -2;
    if (it < nit && idx < n_items) {
      const int tok = idx >> 3;
      const int r1 = tok / a2h, r2 = tok - r1 * a2h;
      const int a1 = a1_0 - 1 + r1, a2 = r2 - 1;
      v = (a1 >= 0 && a1 < H && a2 >= 0 && a2 < W) ? a1 * W + a2 : -1;
    }
    inpl[it] = v;
  }
}
// canonical tokens of plane a0 for this thread's loader items (temporal stack: one table lookup per item). Called ONE
// plane ahead of the copy that uses them, so the dependent LDG -> cp.async address chain is off the critical path.
__device__ __forceinline__ void peg_plane_tokens(int (&tok)[PEG_MAXIT], const PegGeom& g, int a0, const int (&inpl)[PEG_MAXIT],
                                                 int nit) {
  const bool pl_ok = a0 >= 0 && a0 < g.T;
  const int fbase = a0 * g.H * g.W;
#pragma unroll
  for (int it = 0; it < PEG_MAXIT; it++) {
    if (it < nit) {
      const int v = inpl[it];
      int tk = -1;
      if (pl_ok && v >= 0) {
        const int f = fbase + v;
        tk = !g.temporal ? f : (g.table != nullptr ? __ldg(g.table + f) : (int)peg_canon_f(g, f));
      }
      tok[it] = tk;
    }
  }
}
// async load of one plane (rows a1_0-1 .. a1_0+A1T, cols -1 .. W) of a [tokens, D] fp32 tensor into one ring slot;
// tok[] from peg_plane_tokens (-1: zero fill), src already offset to the column's first channel.
__device__ __forceinline__ void peg_load_plane5(float* slot, const float* __restrict__ src, int D, const int (&tok)[PEG_MAXIT],
                                                const int (&inpl)[PEG_MAXIT], int nit) {
  const uint32_t sbase = smem_u32(slot) + threadIdx.x * 16;
  const float* s4 = src + (threadIdx.x & 7) * 4;
#pragma unroll
  for (int it = 0; it < PEG_MAXIT; it++) {
    if (it < nit && inpl[it] != -2) {
      const int tk = tok[it];
      const bool ok = tk >= 0;
      const float* p = ok ? s4 + (long long)tk * D : s4;
      const int bytes = ok ? 16 : 0;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sbase + it * (PEG_THREADS * 16)), "l"(p), "r"(bytes)
                   : "memory");
    }
  }
}
// async load of the upstream-gradient tile (A1T lines x W positions x CB channels) of plane a0 (always inside [0,T))
__device__ __forceinline__ void peg_load_dy4(float* buf, const float* __restrict__ dy, const PegGeom& g, int a0, int a1_0) {
  const int n_tok = A1T * g.W;
  const int n_valid = min(A1T, g.H - a1_0) * g.W;
  const int fbase = (a0 * g.H + a1_0) * g.W;
  const int q4 = (threadIdx.x & 7) * 4;
  for (int idx = threadIdx.x; idx < n_tok * (CB / 4); idx += PEG_THREADS) {
    const int tok = idx >> 3;
    const bool ok = tok < n_valid;
    long long ctok = 0;
    if (ok) {
      const int f = fbase + tok;
      ctok = !g.temporal ? f : (g.table != nullptr ? __ldg(g.table + f) : (int)peg_canon_f(g, f));
    }
    peg_cp16(buf + (size_t)idx * 4, dy + ctok * g.D + q4, ok);
  }
}

// ---- packed fp32x2 arithmetic (Blackwell FFMA2: two FMAs per issued instruction) -----------------------------------
typedef unsigned long long f2_t;   // (lo, hi) = (channel c, channel c+1)
__device__ __forceinline__ f2_t f2_pack(float lo, float hi) {
  f2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float2 f2_unpack(f2_t v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ f2_t f2_fma(f2_t a, f2_t b, f2_t c) {
  f2_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
// shared-memory window loads: [32-bit running byte offset + immediate column offset], 8 bytes = one channel pair
#define PEG_LD2(smb, po, r, col) (*reinterpret_cast<const f2_t*>((smb) + (po)[r] + (col) * (CB * 4)))

// Thread layout of the compute phase (256 threads): pair = tid & 15 (channels 2*pair, 2*pair+1), line = (tid >> 4) & 7,
// seg = tid >> 7: the a2 range [0, W) is cut into PEG_NSEG segments so that every (line, channel pair) is worked on by
// PEG_NSEG threads. A half-warp reads / writes the 128 contiguous bytes of one token's 32 channels.
constexpr int PEG_NSEG = PEG_THREADS / (16 * A1T);   // 2

// Two outputs (a2, a2+1) of the sliding window per call: the window holds FOUR columns per row; the call loads the two
// columns (halo index a2+2, a2+3) into window slots (R0+2, R0+3) mod 4, output a2 uses slots (R0, R0+1, R0+2) and output
// a2+1 uses (R0+1, R0+2, R0+3): four independent FFMA2 chains (2 outputs x 2 partial sums).
template <int R0>
__device__ __forceinline__ void peg_out2(f2_t (&win)[9][4], const char* smb, const uint32_t (&po)[9], const f2_t (&wt)[27],
                                         f2_t init, f2_t& o0, f2_t& o1) {
  constexpr int A = R0 % 4, B = (R0 + 1) % 4, C = (R0 + 2) % 4, D = (R0 + 3) % 4;
#pragma unroll
  for (int r = 0; r < 9; r++) {
    win[r][C] = PEG_LD2(smb, po, r, R0 + 2);
    win[r][D] = PEG_LD2(smb, po, r, R0 + 3);
  }
  // Six independent FFMA2 chains are kept in flight in both passes (ptxas keeps the source order: with two chains the
  // kernel sat in fixed-latency "wait" stalls). Pass 1 only touches the two columns already in registers, so the 18
  // loads above have ~27 issue slots to land; pass 2 continues on the same accumulators.
  f2_t x0[2] = {init, 0ull}, x1[2] = {0ull, 0ull}, y0[2] = {init, 0ull}, y2[2] = {0ull, 0ull};
#pragma unroll
  for (int r = 0; r < 9; r++) {
    const int p = r & 1;
    x0[p] = f2_fma(wt[r * 3 + 0], win[r][A], x0[p]);
    y0[p] = f2_fma(wt[r * 3 + 0], win[r][B], y0[p]);
    x1[p] = f2_fma(wt[r * 3 + 1], win[r][B], x1[p]);
  }
#pragma unroll
  for (int r = 0; r < 9; r++) {
    const int p = r & 1;
    x0[p] = f2_fma(wt[r * 3 + 2], win[r][C], x0[p]);
    y0[p] = f2_fma(wt[r * 3 + 1], win[r][C], y0[p]);
    y2[p] = f2_fma(wt[r * 3 + 2], win[r][D], y2[p]);
  }
  const float2 a0 = f2_unpack(x0[0]), a1 = f2_unpack(x0[1]), a2 = f2_unpack(x1[0]), a3 = f2_unpack(x1[1]);
  const float2 b0 = f2_unpack(y0[0]), b1 = f2_unpack(y0[1]), b2 = f2_unpack(y2[0]), b3 = f2_unpack(y2[1]);
  o0 = f2_pack((a0.x + a1.x) + (a2.x + a3.x), (a0.y + a1.y) + (a2.y + a3.y));
  o1 = f2_pack((b0.x + b1.x) + (b2.x + b3.x), (b0.y + b1.y) + (b2.y + b3.y));
}
template <int R0>
__device__ __forceinline__ void peg_wout2(f2_t (&win)[9][4], const char* smb, const uint32_t (&po)[9], f2_t (&acc)[27],
                                          f2_t d0, f2_t d1) {
  constexpr int A = R0 % 4, B = (R0 + 1) % 4, C = (R0 + 2) % 4, D = (R0 + 3) % 4;
#pragma unroll
  for (int r = 0; r < 9; r++) {
    win[r][C] = PEG_LD2(smb, po, r, R0 + 2);
    win[r][D] = PEG_LD2(smb, po, r, R0 + 3);
  }
#pragma unroll
  for (int r = 0; r < 9; r++) {   // taps on the two columns already in registers first
    acc[r * 3 + 0] = f2_fma(d0, win[r][A], acc[r * 3 + 0]);
    acc[r * 3 + 1] = f2_fma(d0, win[r][B], acc[r * 3 + 1]);
  }
#pragma unroll
  for (int r = 0; r < 9; r++) {
    acc[r * 3 + 0] = f2_fma(d1, win[r][B], acc[r * 3 + 0]);
    acc[r * 3 + 2] = f2_fma(d0, win[r][C], acc[r * 3 + 2]);
    acc[r * 3 + 1] = f2_fma(d1, win[r][C], acc[r * 3 + 1]);
    acc[r * 3 + 2] = f2_fma(d1, win[r][D], acc[r * 3 + 2]);
  }
}

// MODE 0: y = x + conv(x) + bias (forward; taps a0-2..a0)     MODE 1: dx = dy + conv^T(dy) (taps a0..a0+2, mirrored)
template <int MODE>
__global__ void __launch_bounds__(PEG_THREADS, 1) peg_conv4_kernel(ctclip_peg_args a, int steps_per_cta) {
  extern __shared__ __align__(16) float peg_sm[];
  const PegGeom g{a.T, a.H, a.W, a.D, a.temporal, a.canon_table};
  const int a2h = a.W + 2;
  const int slot_elems = (A1T + 2) * a2h * CB;
  const int n_items = (A1T + 2) * a2h * (CB / 4);
  const int nit = (n_items + PEG_THREADS - 1) / PEG_THREADS;
  const int n_a1t = (a.H + A1T - 1) / A1T, n_cb = a.D / CB;
  const int total = n_cb * n_a1t * a.B * a.T;
  const int s_begin = blockIdx.x * steps_per_cta;
  const int s_end = min(total, s_begin + steps_per_cta);
  const long long vol = (long long)a.T * a.H * a.W * a.D;
  const int pair = threadIdx.x & 15, line = (threadIdx.x >> 4) & (A1T - 1), seg = threadIdx.x >> 7;
  const int seg_len = (((a.W + PEG_NSEG - 1) / PEG_NSEG) + 1) & ~1;   // even
  const int b0 = min(a.W, seg * seg_len), b1 = min(a.W, b0 + seg_len);
  int* s_tok = reinterpret_cast<int*>(peg_sm + (size_t)NSLOT * slot_elems);   // [2][A1T][W] canonical token of every output
  const char* smb = reinterpret_cast<const char*>(peg_sm);
  int inpl[PEG_MAXIT];
  int par = 0;
  for (int s = s_begin; s < s_end;) {
    const int col = s / a.T;
    const int p_begin = s - col * a.T;
    const int p_end = min(a.T, p_begin + (s_end - s));
    const PegCol cc = peg_column(col, n_cb, n_a1t);
    const float* xin = a.x + (long long)cc.b * vol + cc.c0;
    float* yout = a.y + (long long)cc.b * vol + cc.c0 + 2 * pair;
    __nv_bfloat16* ybf =
        a.y_bf16 ? reinterpret_cast<__nv_bfloat16*>(a.y_bf16) + (long long)cc.b * vol + cc.c0 + 2 * pair : nullptr;
    const int ch = cc.c0 + 2 * pair;
    f2_t wt[27];
#pragma unroll
    for (int k = 0; k < 27; k++) {
      const int kk = (MODE == 0) ? k : 26 - k;
      wt[k] = f2_pack(__ldg(a.weight + (long long)ch * 27 + kk), __ldg(a.weight + (long long)(ch + 1) * 27 + kk));
    }
    const f2_t bias2 = (MODE == 0 && a.bias != nullptr) ? f2_pack(__ldg(a.bias + ch), __ldg(a.bias + ch + 1)) : 0ull;
    const int a1 = cc.a1_0 + line;
    __syncthreads();   // the previous column's last plane (ring + token table) is fully consumed
    peg_loader_setup(inpl, nit, n_items, a2h, cc.a1_0, a.H, a.W);
    // planes needed for output plane a0: a0-2, a0-1, a0 (MODE 0) / a0, a0+1, a0+2 (MODE 1). Walk a0 in direction `dirn`
    // so that only ONE new plane enters per step.
    const int first = (MODE == 0) ? p_begin : p_end - 1;
    const int dirn = (MODE == 0) ? 1 : -1;
    int ptok[PEG_MAXIT];
    for (int d = 2; d >= 0; d--) {
      const int pl = first - dirn * d;
      peg_plane_tokens(ptok, g, pl, inpl, nit);
      peg_load_plane5(peg_sm + (size_t)peg_slot(pl) * slot_elems, xin, a.D, ptok, inpl, nit);
    }
    peg_commit();
    peg_plane_tokens(ptok, g, first + dirn, inpl, nit);   // tokens of the first prefetched plane
    const int n_steps = p_end - p_begin;
    for (int step = 0; step < n_steps; step++, par ^= 1) {
      const int a0 = first + dirn * step;
      for (int i = threadIdx.x; i < A1T * a.W; i += PEG_THREADS) {
        const int l1 = cc.a1_0 + i / a.W;
        s_tok[par * A1T * a.W + i] = (l1 < a.H) ? (int)peg_canon(g, a0, l1, i % a.W) : 0;
      }
      peg_wait_all();
      __syncthreads();   // plane a0 landed for everyone; everyone finished computing the previous plane
      if (step + 1 < n_steps) {   // prefetch the next plane into the slot released by plane a0 - 3*dirn
        const int pl = a0 + dirn;
        peg_load_plane5(peg_sm + (size_t)peg_slot(pl) * slot_elems, xin, a.D, ptok, inpl, nit);
        peg_commit();
        if (step + 2 < n_steps) peg_plane_tokens(ptok, g, pl + dirn, inpl, nit);   // used one step from now
      }
      if (a1 < a.H && b0 < b1) {
        uint32_t po[9];
#pragma unroll
        for (int r = 0; r < 9; r++) {
          const int k0 = r / 3, k1 = r % 3;
          const int pl = (MODE == 0) ? (a0 + k0 - 2) : (a0 + k0);
          po[r] = (uint32_t)(peg_slot(pl) * slot_elems + ((line + k1) * a2h + b0) * CB + 2 * pair) * 4u;
        }
        f2_t win[9][4];
#pragma unroll
        for (int r = 0; r < 9; r++) {
          win[r][0] = PEG_LD2(smb, po, r, 0);   // a2 = b0 - 1 (zero padding when b0 == 0)
          win[r][1] = PEG_LD2(smb, po, r, 1);   // a2 = b0
        }
        constexpr int CR = (MODE == 0) ? 7 : 1;   // row of the centre tap (k0 = 2 | 0, k1 = 1): its centre element = residual
        const int* tokl = s_tok + par * A1T * a.W + line * a.W;
        auto emit = [&](int a2, f2_t conv, f2_t centre) {
          const float2 cv = f2_unpack(conv), ce = f2_unpack(centre);
          const float2 val = make_float2(cv.x + ce.x, cv.y + ce.y);
          const long long off = (long long)tokl[a2] * a.D;
          *reinterpret_cast<float2*>(yout + off) = val;
          if (ybf != nullptr) *reinterpret_cast<uint32_t*>(ybf + off) = pack_bf16x2(val.x, val.y);
        };
        for (int a2 = b0; a2 < b1; a2 += 4) {
          f2_t o0, o1;
          peg_out2<0>(win, smb, po, wt, bias2, o0, o1);
          emit(a2, o0, win[CR][1]);
          if (a2 + 1 < b1) emit(a2 + 1, o1, win[CR][2]);
          if (a2 + 2 < b1) {
            peg_out2<2>(win, smb, po, wt, bias2, o0, o1);
            emit(a2 + 2, o0, win[CR][3]);
            if (a2 + 3 < b1) emit(a2 + 3, o1, win[CR][0]);
          }
#pragma unroll
          for (int r = 0; r < 9; r++) po[r] += 4 * CB * 4;
        }
      }
    }
    s += n_steps;
  }
}

// dw[c][k] += sum_p dy[p] * x[p + off(k)], db[c] += sum_p dy[p]
// x planes in the 4-slot ring, the upstream-gradient tile of the same plane in a 2-slot ring, both prefetched with cp.async.
// Same persistent (column, plane) ranges and thread layout as the conv kernel; the 27+1 partial sums of a thread are
// reduced across the 8 lines x 2 segments of the CTA through shared memory at the end of every column segment -> one
// atomic per (channel, tap).
__global__ void __launch_bounds__(PEG_THREADS, 1) peg_wgrad4_kernel(ctclip_peg_args a, int steps_per_cta, int dy_double) {
  extern __shared__ __align__(16) float peg_sm[];
  const PegGeom g{a.T, a.H, a.W, a.D, a.temporal, a.canon_table};
  const int a2h = a.W + 2;
  const int slot_elems = (A1T + 2) * a2h * CB;
  const int dy_elems = A1T * a.W * CB;
  const int n_items = (A1T + 2) * a2h * (CB / 4);
  const int nit = (n_items + PEG_THREADS - 1) / PEG_THREADS;
  float* sdy = peg_sm + (size_t)NSLOT * slot_elems;   // [2][A1T][W][CB]
  const int n_a1t = (a.H + A1T - 1) / A1T, n_cb = a.D / CB;
  const int total = n_cb * n_a1t * a.B * a.T;
  const int s_begin = blockIdx.x * steps_per_cta;
  const int s_end = min(total, s_begin + steps_per_cta);
  const long long vol = (long long)a.T * a.H * a.W * a.D;
  const int pair = threadIdx.x & 15, line = (threadIdx.x >> 4) & (A1T - 1), seg = threadIdx.x >> 7;
  const int seg_len = (((a.W + PEG_NSEG - 1) / PEG_NSEG) + 1) & ~1;
  const int b0 = min(a.W, seg * seg_len), b1 = min(a.W, b0 + seg_len);
  const char* smb = reinterpret_cast<const char*>(peg_sm);
  int inpl[PEG_MAXIT];
  for (int s = s_begin; s < s_end;) {
    const int col = s / a.T;
    const int p_begin = s - col * a.T;
    const int p_end = min(a.T, p_begin + (s_end - s));
    const PegCol cc = peg_column(col, n_cb, n_a1t);
    const float* xin = a.x + (long long)cc.b * vol + cc.c0;
    const float* dy = a.dy + (long long)cc.b * vol + cc.c0;
    const int a1 = cc.a1_0 + line;
    f2_t acc[27];
#pragma unroll
    for (int k = 0; k < 27; k++) acc[k] = 0ull;
    float accb0 = 0.f, accb1 = 0.f;
    __syncthreads();   // previous column: reduction scratch / ring fully consumed
    peg_loader_setup(inpl, nit, n_items, a2h, cc.a1_0, a.H, a.W);
    int ptok[PEG_MAXIT];
    for (int d = 2; d >= 0; d--) {
      const int pl = p_begin - d;
      peg_plane_tokens(ptok, g, pl, inpl, nit);
      peg_load_plane5(peg_sm + (size_t)peg_slot(pl) * slot_elems, xin, a.D, ptok, inpl, nit);
    }
    peg_load_dy4(sdy, dy, g, p_begin, cc.a1_0);
    peg_commit();
    peg_plane_tokens(ptok, g, p_begin + 1, inpl, nit);
    for (int a0 = p_begin; a0 < p_end; a0++) {
      const int par = dy_double ? ((a0 - p_begin) & 1) : 0;
      if (!dy_double && a0 > p_begin) {   // wide grids: a single gradient tile fits; fetch it without overlap
        __syncthreads();
        peg_load_dy4(sdy, dy, g, a0, cc.a1_0);
        peg_commit();
      }
      peg_wait_all();
      __syncthreads();
      if (a0 + 1 < p_end) {
        peg_load_plane5(peg_sm + (size_t)peg_slot(a0 + 1) * slot_elems, xin, a.D, ptok, inpl, nit);
        if (dy_double) peg_load_dy4(sdy + (size_t)(par ^ 1) * dy_elems, dy, g, a0 + 1, cc.a1_0);
        peg_commit();
        if (a0 + 2 < p_end) peg_plane_tokens(ptok, g, a0 + 2, inpl, nit);
      }
      if (a1 < a.H && b0 < b1) {
        uint32_t po[9];
#pragma unroll
        for (int r = 0; r < 9; r++) {
          const int k0 = r / 3, k1 = r % 3;
          po[r] = (uint32_t)(peg_slot(a0 + k0 - 2) * slot_elems + ((line + k1) * a2h + b0) * CB + 2 * pair) * 4u;
        }
        const f2_t* dyl = reinterpret_cast<const f2_t*>(sdy + (size_t)par * dy_elems + (size_t)(line * a.W) * CB + 2 * pair);
        f2_t win[9][4];
#pragma unroll
        for (int r = 0; r < 9; r++) {
          win[r][0] = PEG_LD2(smb, po, r, 0);
          win[r][1] = PEG_LD2(smb, po, r, 1);
        }
        for (int a2 = b0; a2 < b1; a2 += 4) {
          const f2_t d0 = dyl[a2 * (CB / 2)];
          const f2_t d1 = (a2 + 1 < b1) ? dyl[(a2 + 1) * (CB / 2)] : 0ull;
          const f2_t d2 = (a2 + 2 < b1) ? dyl[(a2 + 2) * (CB / 2)] : 0ull;
          const f2_t d3 = (a2 + 3 < b1) ? dyl[(a2 + 3) * (CB / 2)] : 0ull;
          const float2 e0 = f2_unpack(d0), e1 = f2_unpack(d1), e2 = f2_unpack(d2), e3 = f2_unpack(d3);
          accb0 += (e0.x + e1.x) + (e2.x + e3.x);
          accb1 += (e0.y + e1.y) + (e2.y + e3.y);
          peg_wout2<0>(win, smb, po, acc, d0, d1);
          if (a2 + 2 < b1) peg_wout2<2>(win, smb, po, acc, d2, d3);
#pragma unroll
          for (int r = 0; r < 9; r++) po[r] += 4 * CB * 4;
        }
      }
    }
    // reduce the (line, segment) partial sums of this CTA through shared memory, one atomic per (channel, tap)
    __syncthreads();
    float* red = peg_sm;  // [A1T * PEG_NSEG][28][CB]
    const int slot = line * PEG_NSEG + seg;
#pragma unroll
    for (int k = 0; k < 27; k++) {
      const float2 v = f2_unpack(acc[k]);
      *reinterpret_cast<float2*>(red + (slot * 28 + k) * CB + 2 * pair) = v;
    }
    *reinterpret_cast<float2*>(red + (slot * 28 + 27) * CB + 2 * pair) = make_float2(accb0, accb1);
    __syncthreads();
    for (int i = threadIdx.x; i < 28 * CB; i += PEG_THREADS) {
      const int k = i / CB, c = i % CB;
      float t = 0.f;
#pragma unroll
      for (int l = 0; l < A1T * PEG_NSEG; l++) t += red[(l * 28 + k) * CB + c];
      if (k < 27) atomicAdd(a.dweight + (long long)(cc.c0 + c) * 27 + k, t);
      else if (a.dbias != nullptr) atomicAdd(a.dbias + cc.c0 + c, t);
    }
    s += p_end - p_begin;
  }
}

}  // namespace ctb

namespace ctb {   // plane-streaming path (peg_stream.cu): the default whenever the token grid allows it
bool peg_stream_supported(const ctclip_peg_args* a, bool wgrad);
int peg_stream_launch_conv(int mode, const ctclip_peg_args* a, cudaStream_t stream);
int peg_stream_launch_wgrad(const ctclip_peg_args* a, cudaStream_t stream);
}

using namespace ctb;

static int peg_check(const ctclip_peg_args* a, const char* who) {
  CTB_CHECK_ARG(a && a->x, "%s: null x", who);
  CTB_CHECK_ARG(a->B > 0 && a->T > 0 && a->H > 0 && a->W > 0, "%s: bad grid", who);
  CTB_CHECK_ARG(a->D % 32 == 0, "%s: D must be a multiple of 32", who);
  CTB_CHECK_ARG(a->W <= 42, "%s: token grid width %d > 42 is not supported by the shared-memory plane ring", who, a->W);
  CTB_CHECK_ARG((long long)a->B * a->T * a->H * a->W < (1ll << 31) / 64, "%s: token count too large", who);
  return CTCLIP_OK;
}

// Persistent launch: all (column, plane) steps are cut into one contiguous range per SM (1 CTA per SM: 133-182 KB of
// shared memory).
static void peg_launch_shape(const ctclip_peg_args* a, int dy_bufs, int* grid, int* steps_per_cta, size_t* smem) {
  const int n_a1t = (a->H + A1T - 1) / A1T;
  const long long total = (long long)(a->D / CB) * n_a1t * a->B * a->T;
  long long ctas = num_sms();
  if (ctas > total) ctas = total;
  *steps_per_cta = (int)((total + ctas - 1) / ctas);
  *grid = (int)((total + *steps_per_cta - 1) / *steps_per_cta);
  *smem = sizeof(float) * NSLOT * (size_t)(A1T + 2) * (a->W + 2) * CB;
  if (dy_bufs == 0) *smem += sizeof(int) * 2 * (size_t)A1T * a->W;   // token-index table of the conv kernels
  *smem += sizeof(float) * dy_bufs * (size_t)A1T * a->W * CB;
  *smem += 1024;   // the rotating window reads up to 6 columns past the end of the last row
  const size_t red_bytes = sizeof(float) * A1T * PEG_NSEG * 28 * CB;  // weight-gradient reduction scratch
  if (*smem < red_bytes) *smem = red_bytes;
}

static int peg_launch_conv(int mode, const ctclip_peg_args* a, cudaStream_t stream) {
  int grid, spc;
  size_t smem;
  peg_launch_shape(a, 0, &grid, &spc, &smem);
  CTB_CHECK_ARG(smem <= 227 * 1024, "peg: token grid width %d needs %zu B of shared memory", a->W, smem);
  if (mode == 0) {
    CTB_CUDA(cudaFuncSetAttribute(peg_conv4_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    peg_conv4_kernel<0><<<grid, PEG_THREADS, smem, stream>>>(*a, spc);
  } else {
    CTB_CUDA(cudaFuncSetAttribute(peg_conv4_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    peg_conv4_kernel<1><<<grid, PEG_THREADS, smem, stream>>>(*a, spc);
  }
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}

static int peg_launch_wgrad(const ctclip_peg_args* a, cudaStream_t stream) {
  int grid, spc;
  size_t smem;
  int dy_bufs = 2;
  peg_launch_shape(a, dy_bufs, &grid, &spc, &smem);
  if (smem > 227 * 1024) {
    dy_bufs = 1;
    peg_launch_shape(a, dy_bufs, &grid, &spc, &smem);
  }
  CTB_CHECK_ARG(smem <= 227 * 1024, "peg: token grid width %d needs %zu B of shared memory", a->W, smem);
  CTB_CUDA(cudaFuncSetAttribute(peg_wgrad4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  peg_wgrad4_kernel<<<grid, PEG_THREADS, smem, stream>>>(*a, spc, dy_bufs == 2);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}

extern "C" int ctclip_peg_fwd(const ctclip_peg_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = peg_check(a, "peg_fwd")) return rc;
  CTB_CHECK_ARG(a->y && a->weight, "peg_fwd: null y/weight");
  if (peg_stream_supported(a, false)) return peg_stream_launch_conv(0, a, stream);
  return peg_launch_conv(0, a, stream);
}

// x = upstream gradient dy (fp32), y = dx out (fp32), y_bf16 = optional bf16 copy of dx
extern "C" int ctclip_peg_bwd_data(const ctclip_peg_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = peg_check(a, "peg_bwd_data")) return rc;
  CTB_CHECK_ARG(a->y && a->weight, "peg_bwd_data: null y/weight");
  if (peg_stream_supported(a, false)) return peg_stream_launch_conv(1, a, stream);
  return peg_launch_conv(1, a, stream);
}

// x = forward input, dy = upstream gradient; dweight [D,27] and dbias [D] are accumulated (atomics)
extern "C" int ctclip_peg_bwd_weight(const ctclip_peg_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = peg_check(a, "peg_bwd_weight")) return rc;
  CTB_CHECK_ARG(a->dy && a->dweight, "peg_bwd_weight: null dy/dweight");
  if (peg_stream_supported(a, true)) return peg_stream_launch_wgrad(a, stream);
  return peg_launch_wgrad(a, stream);
}
