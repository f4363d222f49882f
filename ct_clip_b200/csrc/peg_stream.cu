// PEG, plane-streaming kernels (the default path when the token grid allows it; csrc/peg.cu keeps the general kernels).
//
// Same operation as csrc/peg.cu: the depthwise causal 3x3x3 convolution + residual of attention.py:63-84 / :324 on the fp32
// token stream [b, t, h, w, D] and its two backward passes. HBM-bound by design: 4 B read + 4 B written per element
// (8 B read for the weight gradient) against 27 fp32 FMAs per element (42 us of FMA-pipe time per 226 MB tensor).
//
// Why a second formulation. The v4 kernels of peg.cu are OUTPUT-stationary: an output needs 27 inputs, the sliding register
// window still reloads 9 of them per output from shared memory, holds 250 registers per thread and therefore runs 8 warps
// per SM: 240-290 us per launch, 23-33 % of the HBM roofline (profiles/r1f_ncu_full_top_kernels.txt, r2e_stages.md).
// Here the stencil is INPUT-stationary along the causal axis:
//   * a column = (channel block of 32, volume, tile of 8 lines along a1) is walked along the causal axis a0; the walk
//     brings ONE input plane (10 x (W+2) halo tokens x 128 B) per step into a ring slot with ONE 5-D TMA box copy
//     (out-of-bounds coordinates give the zero padding of F.pad for free; no loader instructions at all);
//   * a thread owns a channel PAIR (packed fp32x2 FFMA2) and a 2 x 4 patch of positions. It keeps the partial sums of THREE
//     output planes in registers (3 x 8 packed accumulators): input plane j adds its k0 = 2 taps to output plane j (which is
//     then complete and stored), its k0 = 1 taps to plane j+1 and its k0 = 0 taps to plane j+2;
//   * per step a thread loads its 4 x 6 input patch ONCE (24 LDS.64) and issues 216 FFMA2: 3 shared loads per output
//     instead of 9, no address arithmetic in the inner loop (all offsets are immediates of one base register), no
//     __syncthreads (a full mbarrier per ring slot; the last warp that releases a slot refills it, see PsSeq);
//   * 12 compute warps per SM with 24 independent accumulation chains each.
// The weight gradient uses the same walk with the roles swapped: the x plane is loaded once per step, the upstream-gradient
// tiles of planes j, j+1, j+2 sit in a second ring, 27 packed accumulators per thread live across the whole channel block
// and are flushed with shared-memory atomics + one global atomic per (channel, tap).
//
// Temporal stack (SURVEY trap T1: `(b h w) t d` memory re-read as a (T,H,W) grid): when T == H == W the reshape is a pure
// axis permutation -- conv-grid (a0,a1,a2) = canonical (ih, iw, it). The tensor map keeps the canonical (monotonic) stride
// order (c, iw = a1, ih = a0, it = a2, b), so the box lands TRANSPOSED in shared memory ([a2][a1][c] instead of [a1][a2][c]):
// template parameter TR swaps the two shared-memory strides, everything else is shared.
// Other temporal geometries, W > 24 and D % 32 != 0 stay on the peg.cu kernels (ctb::peg_stream_supported).
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/ctclip_b200.h"

namespace ctb {

namespace {

typedef unsigned long long f2_t;   // (lo, hi) = (channel c, channel c+1)
__device__ __forceinline__ f2_t ps_pack(float lo, float hi) {
  f2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float2 ps_unpack(f2_t v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ f2_t ps_fma(f2_t a, f2_t b, f2_t c) {
  f2_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f2_t ps_lds(uint32_t addr) {
  f2_t v;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr));
  return v;
}

constexpr int PS_A1T = 8;        // output lines per column tile (4 thread rows of 2 lines)
constexpr int PS_CB = 32;        // channels per column: 16 lane pairs = 128 B per token
constexpr int PS_MAXCG = 6;      // thread columns of 4 outputs: W <= 24
constexpr int PS_THREADS = 64 * PS_MAXCG;        // 12 warps (launch bound: 168 registers, no spills in the plane step)
constexpr int PS_MAXSLOTS = 8;

struct PsParams {
  float* y;
  __nv_bfloat16* ybf;
  const float* weight;
  const float* bias;
  float* dweight;
  float* dbias;
  int B, T, H, W, D;
  int n_cg, n_a1t, n_cb;
  int total, steps_per_cta;
  int s0, s1, s2;        // element strides (x D) of the conv-grid axes a0, a1, a2 in the canonical [b, t, h, w, D] tensor
  long long vol;         // elements per volume
  int ns, ndy;           // ring depths (input planes / upstream-gradient tiles)
  int slot_bytes, dy_bytes;
};

struct PsCol {
  int cb, b, a1_0;
};
__device__ __forceinline__ PsCol ps_column(const PsParams& p, int col) {
  PsCol c;
  c.a1_0 = (col % p.n_a1t) * PS_A1T;
  c.b = (col / p.n_a1t) % p.B;
  c.cb = col / (p.n_a1t * p.B);
  return c;
}

// One plane step of the convolution kernels. R = (walk index - first walk index of the segment) % 3: the accumulator set of
// output plane j is S[(j - j_start) % 3], so all register indices are compile-time constants.
// patch element (li, ci) of a slot: [a1][a2][128 B] (roww bytes per line), or [a2][a1][128 B] when TR
template <bool TR>
__device__ __forceinline__ uint32_t ps_off(int li, int ci, uint32_t roww) {
  return TR ? (uint32_t)(ci * (PS_A1T + 2) + li) * 128u : (uint32_t)li * roww + (uint32_t)ci * 128u;
}
template <bool TR>
__device__ __forceinline__ uint32_t ps_dyoff(int lo, int co, uint32_t dyroww) {
  return TR ? (uint32_t)(co * PS_A1T + lo) * 128u : (uint32_t)lo * dyroww + (uint32_t)co * 128u;
}

template <int R, bool TR>
__device__ __forceinline__ void ps_conv_step(f2_t (&S)[3][2][4], const f2_t (&wt)[27], uint32_t base, uint32_t roww) {
#pragma unroll
  for (int li = 0; li < 4; li++) {
    f2_t xr[6];
#pragma unroll
    for (int ci = 0; ci < 6; ci++) xr[ci] = ps_lds(base + ps_off<TR>(li, ci, roww));
#pragma unroll
    for (int lo = 0; lo < 2; lo++) {
      const int k1 = li - lo;
      if (k1 < 0 || k1 > 2) continue;
#pragma unroll
      for (int k2 = 0; k2 < 3; k2++)
#pragma unroll
        for (int d = 0; d < 3; d++)
#pragma unroll
          for (int co = 0; co < 4; co++)
            S[(R + d) % 3][lo][co] = ps_fma(wt[d * 9 + k1 * 3 + k2], xr[co + k2], S[(R + d) % 3][lo][co]);
    }
  }
}

// Who issues the TMA copies. There is NO producer warp: 12 warps keep the register allocation at 168 per thread, a 13th warp is
// accounted as 16 warps = 128 registers. Two producer-warp variants were measured on B200 (tools/peg_stream_probe.py, fwd /
// bwd_data / bwd_weight per launch): (a) producer warp at 128 registers, ~30 spilled values: 152 / 174 / 141 us; (b) producer
// warpgroup + setmaxnreg 40/152: 113 / 148 / 117 us but INTERMITTENTLY wrong outputs (tools/peg_debug.py: whole-warp errors in
// exactly the outputs whose accumulators / store addresses live in R128+, while other warps of the SM reload their weights),
// so setmaxnreg is not used. Instead the LAST warp that releases a ring slot refills it: every warp bumps a per-slot counter
// in shared memory (atom.acq_rel) when it is done with the slot; the warp that sees the 12th arrival of this use knows all
// reads are over and issues the copy of the plane that uses the slot next (the ring position of every copy is a closed-form
// function of its sequence number, PsSeq). Nobody polls, nobody waits for a release.
struct PsSeq {
  int col0, js0, len0, total;   // input planes: first segment (column col0, planes js0 .. js0+len0-1), then whole columns
  int djb0, dlen0, dtotal;      // weight gradient: upstream-gradient tiles, same shape of sequence
};
template <int KIND>   // 0 / 1: convolutions (two priming planes when the range starts mid-column), 2: weight gradient
__device__ __forceinline__ PsSeq ps_sequence(const PsParams& p, int s_begin, int s_end) {
  PsSeq q;
  const int n = s_end - s_begin;
  q.col0 = s_begin / p.T;
  const int jb0 = s_begin - q.col0 * p.T, je0 = min(p.T, jb0 + n);
  q.js0 = (KIND == 2) ? jb0 : max(0, jb0 - 2);
  q.len0 = je0 - q.js0;
  q.total = q.len0 + (n - (je0 - jb0));
  q.djb0 = jb0;
  q.dlen0 = min(p.T - 1, je0 + 1) - jb0 + 1;
  const int r = n - (je0 - jb0), rem = r % p.T;
  q.dtotal = q.dlen0 + (r / p.T) * p.T + (rem ? min(p.T - 1, rem + 1) + 1 : 0);
  if (n <= 0) q.total = q.dtotal = 0;
  return q;
}
// input plane number G of the CTA's sequence -> ring slot G % ns
template <int KIND, bool TR>
__device__ __forceinline__ void ps_issue_x(const PsParams& p, const PsSeq& q, int G, uint8_t* ring, uint64_t* full, const CUtensorMap* tx) {
  if (G >= q.total) return;
  int col = q.col0, j = q.js0 + G;
  if (G >= q.len0) {
    const int r = G - q.len0, c = r / p.T;
    col = q.col0 + 1 + c;
    j = r - c * p.T;
  }
  const PsCol cc = ps_column(p, col);
  const int slot = G % p.ns;
  const int plane = (KIND == 1) ? p.T - 1 - j : j;
  mbar_arrive_expect_tx(&full[slot], (uint32_t)p.slot_bytes);
  uint8_t* dst = ring + (size_t)slot * p.slot_bytes;
  if (TR) tma_load_5d(dst, tx, &full[slot], cc.cb * PS_CB, cc.a1_0 - 1, plane, -1, cc.b);
  else tma_load_5d(dst, tx, &full[slot], cc.cb * PS_CB, -1, cc.a1_0 - 1, plane, cc.b);
}
template <bool TR>
__device__ __forceinline__ void ps_issue_dy(const PsParams& p, const PsSeq& q, int Hn, uint8_t* dring, uint64_t* dfull, const CUtensorMap* tdy) {
  if (Hn >= q.dtotal) return;
  int col = q.col0, plane = q.djb0 + Hn;
  if (Hn >= q.dlen0) {
    const int r = Hn - q.dlen0, c = r / p.T;
    col = q.col0 + 1 + c;
    plane = r - c * p.T;
  }
  const PsCol cc = ps_column(p, col);
  const int slot = Hn % p.ndy;
  mbar_arrive_expect_tx(&dfull[slot], (uint32_t)p.dy_bytes);
  uint8_t* dst = dring + (size_t)slot * p.dy_bytes;
  if (TR) tma_load_5d(dst, tdy, &dfull[slot], cc.cb * PS_CB, cc.a1_0, plane, 0, cc.b);
  else tma_load_5d(dst, tdy, &dfull[slot], cc.cb * PS_CB, 0, cc.a1_0, plane, cc.b);
}
// one warp is done with use number `use` of a ring slot (lane 0 calls this after a __syncwarp): true for the last of n_warps.
// A RELAXED shared-memory atomic: the acq_rel form costs two MEMBAR.ALL.CTA per step, which wait for the step's outstanding
// global stores (measured: forward 113 -> 177 us). What has to be ordered is only "this warp's shared-memory reads of the
// slot are over before the refill lands": the increment is made data-dependent on accumulators (`dep`, masked with a
// run-time zero ptxas cannot fold) that the LAST loads of the step feed, so the atomic cannot issue before those loads have
// returned, and the refill it may trigger is a DRAM/L2 round trip away.
// The returned count is only LOOKED at after the step's stores have been issued (ps_last), so the atomic's round trip is off the
// warp's critical path.
__device__ __forceinline__ uint32_t ps_release(uint32_t* cnt, uint32_t dep, uint32_t rt_zero) {
  uint32_t old;
  const uint32_t inc = 1u + (dep & rt_zero);
  asm volatile("atom.relaxed.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(cnt)), "r"(inc) : "memory");
  return old;
}
__device__ __forceinline__ bool ps_last(uint32_t old, int use, int n_warps) { return old == (uint32_t)(n_warps * (use + 1) - 1); }
__device__ __forceinline__ uint32_t ps_lo32(f2_t v) { return (uint32_t)v; }

// Store of the completed output plane (accumulator set R) and re-initialisation of that set for output plane j + 3.
template <int R, int MODE>
__device__ __forceinline__ void ps_conv_emit(f2_t (&S)[3][2][4], bool store, unsigned vmask, float* yb, __nv_bfloat16* ybf, int s1, int s2,
                                             f2_t bias2) {
  if (store) {
    if (vmask == 0xffu) {
#pragma unroll
      for (int lo = 0; lo < 2; lo++)
#pragma unroll
        for (int co = 0; co < 4; co++) {
          const float2 v = ps_unpack(S[R][lo][co]);
          const int off = lo * s1 + co * s2;
          *reinterpret_cast<float2*>(yb + off) = v;
          if (MODE == 1 && ybf != nullptr) *reinterpret_cast<uint32_t*>(ybf + off) = pack_bf16x2(v.x, v.y);
        }
    } else {
#pragma unroll
      for (int lo = 0; lo < 2; lo++)
#pragma unroll
        for (int co = 0; co < 4; co++)
          if (vmask & (1u << (lo * 4 + co))) {
            const float2 v = ps_unpack(S[R][lo][co]);
            const int off = lo * s1 + co * s2;
            *reinterpret_cast<float2*>(yb + off) = v;
            if (MODE == 1 && ybf != nullptr) *reinterpret_cast<uint32_t*>(ybf + off) = pack_bf16x2(v.x, v.y);
          }
    }
  }
#pragma unroll
  for (int lo = 0; lo < 2; lo++)
#pragma unroll
    for (int co = 0; co < 4; co++) S[R][lo][co] = bias2;
}

// MODE 0: y = x + conv(x) + bias, walk a0 = 0 .. T-1.   MODE 1: dx = dy + conv^T(dy), walk a0 = T-1 .. 0 (mirrored taps).
template <int MODE, bool TR>
__global__ void __launch_bounds__(PS_THREADS, 1) peg_stream_conv_kernel(const __grid_constant__ CUtensorMap tx, const PsParams p) {
  extern __shared__ uint8_t ps_smem_raw[];
  uint8_t* smem = ps_smem_raw + ((128u - (smem_u32(ps_smem_raw) & 127u)) & 127u);
  uint8_t* ring = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)p.ns * p.slot_bytes);
  uint32_t* cnt = reinterpret_cast<uint32_t*>(full + PS_MAXSLOTS);
  const int tid = threadIdx.x, lane = tid & 31;
  const int n_warps = (int)blockDim.x >> 5;
  const int s_begin = (int)blockIdx.x * p.steps_per_cta;
  const int s_end = min(p.total, s_begin + p.steps_per_cta);
  const PsSeq seq = ps_sequence<MODE>(p, s_begin, s_end);
  const uint32_t rt_zero = blockDim.x >> 16;   // 0, but not a compile-time constant
  if (tid == 0) {
    for (int i = 0; i < p.ns; i++) {
      mbar_init(&full[i], 1);
      cnt[i] = 0;
    }
    fence_barrier_init();
    tma_prefetch_desc(&tx);
    for (int G = 0; G < p.ns; G++) ps_issue_x<MODE, TR>(p, seq, G, ring, full, &tx);   // fill the ring
  }
  __syncthreads();
  const int pair = tid & 15, tile = tid >> 4;
  const int cg = tile % p.n_cg, lp = tile / p.n_cg;
  const int l0 = 2 * lp, q0 = 4 * cg;
  const uint32_t roww = (uint32_t)(4 * p.n_cg + 2) * 128u;
  const uint32_t toff = ps_off<TR>(l0, q0, roww) + (uint32_t)pair * 8u;
  const uint32_t ring_u32 = smem_u32(ring);
  f2_t wt[27];
  f2_t bias2 = 0ull;
  int cur_cb = -1;
  int g = 0, slot = 0, use = 0;   // step number of the CTA's sequence, its ring slot g % ns and use number g / ns
  for (int s = s_begin; s < s_end;) {
    const int col = s / p.T, jb = s - col * p.T, je = min(p.T, jb + (s_end - s));
    const PsCol cc = ps_column(p, col);
    const int js = max(0, jb - 2);
    if (cc.cb != cur_cb) {
      cur_cb = cc.cb;
      const int ch = cc.cb * PS_CB + 2 * pair;
#pragma unroll
      for (int d = 0; d < 3; d++)
#pragma unroll
        for (int k1 = 0; k1 < 3; k1++)
#pragma unroll
          for (int k2 = 0; k2 < 3; k2++) {
            const int src = (MODE == 0) ? ((2 - d) * 9 + k1 * 3 + k2) : ((2 - d) * 9 + (2 - k1) * 3 + (2 - k2));
            float w0 = __ldg(p.weight + (long long)ch * 27 + src), w1 = __ldg(p.weight + (long long)(ch + 1) * 27 + src);
            if (d == 0 && k1 == 1 && k2 == 1) {   // the residual rides on the centre tap of the completing plane
              w0 += 1.0f;
              w1 += 1.0f;
            }
            wt[d * 9 + k1 * 3 + k2] = ps_pack(w0, w1);
          }
      bias2 = (MODE == 0 && p.bias != nullptr) ? ps_pack(__ldg(p.bias + ch), __ldg(p.bias + ch + 1)) : 0ull;
    }
    // outputs of this thread: lines a1_0 + l0 + {0,1}, positions q0 + {0..3}
    unsigned vmask = 0;
#pragma unroll
    for (int lo = 0; lo < 2; lo++)
#pragma unroll
      for (int co = 0; co < 4; co++)
        if (cc.a1_0 + l0 + lo < p.H && q0 + co < p.W) vmask |= 1u << (lo * 4 + co);
    const long long obase = (long long)cc.b * p.vol + cc.cb * PS_CB + 2 * pair + (long long)(cc.a1_0 + l0) * p.s1 + (long long)q0 * p.s2;
    f2_t S[3][2][4];
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
      for (int lo = 0; lo < 2; lo++)
#pragma unroll
        for (int co = 0; co < 4; co++) S[a][lo][co] = bias2;
    int ph = 0;
    for (int j = js; j < je; j++, g++) {
      mbar_wait_tag(&full[slot], (uint32_t)(use & 1), 61);
      const uint32_t base = ring_u32 + (uint32_t)slot * (uint32_t)p.slot_bytes + toff;
      uint32_t dep;   // accumulators fed by the last row of the patch (see ps_release)
      if (ph == 0) {
        ps_conv_step<0, TR>(S, wt, base, roww);
        dep = ps_lo32(S[0][1][0]) ^ ps_lo32(S[0][1][3]);
      } else if (ph == 1) {
        ps_conv_step<1, TR>(S, wt, base, roww);
        dep = ps_lo32(S[1][1][0]) ^ ps_lo32(S[1][1][3]);
      } else {
        ps_conv_step<2, TR>(S, wt, base, roww);
        dep = ps_lo32(S[2][1][0]) ^ ps_lo32(S[2][1][3]);
      }
      // this warp's reads of the slot are over; the last of the warps refills it with the plane of step g + ns
      __syncwarp();
      uint32_t arrivals = 0;
      if (lane == 0) arrivals = ps_release(&cnt[slot], dep, rt_zero);
      // output plane j is complete after this step: store it (not during the priming steps of a range that starts mid-column)
      const long long ob = obase + (long long)(MODE == 0 ? j : p.T - 1 - j) * p.s0;
      float* yb = p.y + ob;
      __nv_bfloat16* ybf = (MODE == 1 && p.ybf != nullptr) ? p.ybf + ob : nullptr;
      const bool store = j >= jb;
      if (ph == 0) {
        ps_conv_emit<0, MODE>(S, store, vmask, yb, ybf, p.s1, p.s2, bias2);
        ph = 1;
      } else if (ph == 1) {
        ps_conv_emit<1, MODE>(S, store, vmask, yb, ybf, p.s1, p.s2, bias2);
        ph = 2;
      } else {
        ps_conv_emit<2, MODE>(S, store, vmask, yb, ybf, p.s1, p.s2, bias2);
        ph = 0;
      }
      if (lane == 0 && ps_last(arrivals, use, n_warps)) ps_issue_x<MODE, TR>(p, seq, g + p.ns, ring, full, &tx);
      if (++slot == p.ns) {
        slot = 0;
        use++;
      }
    }
    s += je - jb;
  }
}

// dw[c][k] += sum_p dy[p] * x[p + off(k)], db[c] += sum_p dy[p]   (attention.py:63-84 backward w.r.t. the conv parameters)
template <bool TR>
__global__ void __launch_bounds__(PS_THREADS, 1)
peg_stream_wgrad_kernel(const __grid_constant__ CUtensorMap tx, const __grid_constant__ CUtensorMap tdy, const PsParams p) {
  extern __shared__ uint8_t ps_smem_raw[];
  uint8_t* smem = ps_smem_raw + ((128u - (smem_u32(ps_smem_raw) & 127u)) & 127u);
  uint8_t* ring = smem;
  uint8_t* dring = ring + (size_t)p.ns * p.slot_bytes;
  float* red = reinterpret_cast<float*>(dring + (size_t)p.ndy * p.dy_bytes);   // [28][32]
  uint64_t* full = reinterpret_cast<uint64_t*>(red + 28 * PS_CB);
  uint64_t* dfull = full + PS_MAXSLOTS;
  uint32_t* cnt = reinterpret_cast<uint32_t*>(dfull + PS_MAXSLOTS);
  uint32_t* dcnt = cnt + PS_MAXSLOTS;
  const int tid = threadIdx.x, lane = tid & 31;
  const int n_thr = (int)blockDim.x, n_warps = n_thr >> 5;
  const int s_begin = (int)blockIdx.x * p.steps_per_cta;
  const int s_end = min(p.total, s_begin + p.steps_per_cta);
  const PsSeq seq = ps_sequence<2>(p, s_begin, s_end);
  const uint32_t rt_zero = blockDim.x >> 16;   // 0, but not a compile-time constant
  if (tid == 0) {
    for (int i = 0; i < p.ns; i++) {
      mbar_init(&full[i], 1);
      cnt[i] = 0;
    }
    for (int i = 0; i < p.ndy; i++) {
      mbar_init(&dfull[i], 1);
      dcnt[i] = 0;
    }
    fence_barrier_init();
    tma_prefetch_desc(&tx);
    tma_prefetch_desc(&tdy);
    for (int Hn = 0; Hn < p.ndy; Hn++) ps_issue_dy<TR>(p, seq, Hn, dring, dfull, &tdy);
    for (int G = 0; G < p.ns; G++) ps_issue_x<2, TR>(p, seq, G, ring, full, &tx);
  }
  for (int i = tid; i < 28 * PS_CB; i += n_thr) red[i] = 0.f;
  __syncthreads();
  const int pair = tid & 15, tile = tid >> 4;
  const int cg = tile % p.n_cg, lp = tile / p.n_cg;
  const int l0 = 2 * lp, q0 = 4 * cg;
  const uint32_t roww = (uint32_t)(4 * p.n_cg + 2) * 128u;
  const uint32_t toff = ps_off<TR>(l0, q0, roww) + (uint32_t)pair * 8u;
  const uint32_t dyroww = (uint32_t)(4 * p.n_cg) * 128u;
  const uint32_t dytoff = ps_dyoff<TR>(l0, q0, dyroww) + (uint32_t)pair * 8u;
  const uint32_t ring_u32 = smem_u32(ring), dring_u32 = smem_u32(dring);
  const f2_t one2 = ps_pack(1.0f, 1.0f);
  f2_t acc[27];
  f2_t accb = 0ull;
#pragma unroll
  for (int k = 0; k < 27; k++) acc[k] = 0ull;
  int cur_cb = -1;
  auto flush = [&](int cb) {
    // the two patches of a warp first (lanes l and l^16 hold the same channel pair), then shared-memory atomics over the
    // warps, then one global atomic per (channel, tap)
#pragma unroll
    for (int k = 0; k < 28; k++) {
      float2 v = ps_unpack(k < 27 ? acc[k < 27 ? k : 0] : accb);
      v.x += __shfl_xor_sync(0xffffffffu, v.x, 16);
      v.y += __shfl_xor_sync(0xffffffffu, v.y, 16);
      if (lane < 16) {
        atomicAdd(red + k * PS_CB + 2 * pair, v.x);
        atomicAdd(red + k * PS_CB + 2 * pair + 1, v.y);
      }
    }
    __syncthreads();
    for (int i = tid; i < 28 * PS_CB; i += n_thr) {
      const int k = i / PS_CB, c = i % PS_CB;
      const float t = red[i];
      red[i] = 0.f;
      if (k < 27) atomicAdd(p.dweight + (long long)(cb * PS_CB + c) * 27 + k, t);
      else if (p.dbias != nullptr) atomicAdd(p.dbias + cb * PS_CB + c, t);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 27; k++) acc[k] = 0ull;
    accb = 0ull;
  };
  // release of gradient tile number hd of the CTA's sequence; the last warp loads the tile that uses the slot next
  auto release_dy = [&](int hd, uint32_t dep) {
    if (lane == 0 && ps_last(ps_release(&dcnt[hd % p.ndy], dep, rt_zero), hd / p.ndy, n_warps))
      ps_issue_dy<TR>(p, seq, hd + p.ndy, dring, dfull, &tdy);
  };
  int g = 0, slot = 0, use = 0, h0 = 0;
  for (int s = s_begin; s < s_end;) {
    const int col = s / p.T, jb = s - col * p.T, je = min(p.T, jb + (s_end - s));
    const PsCol cc = ps_column(p, col);
    const int last_dy = min(p.T - 1, je + 1);
    if (cc.cb != cur_cb) {
      if (cur_cb >= 0) flush(cur_cb);
      cur_cb = cc.cb;
    }
    for (int j = jb; j < je; j++, g++) {
      mbar_wait_tag(&full[slot], (uint32_t)(use & 1), 64);
      const uint32_t base = ring_u32 + (uint32_t)slot * (uint32_t)p.slot_bytes + toff;
      f2_t xr[4][6];
#pragma unroll
      for (int li = 0; li < 4; li++)
#pragma unroll
        for (int ci = 0; ci < 6; ci++) xr[li][ci] = ps_lds(base + ps_off<TR>(li, ci, roww));
      uint32_t arr_x = 0, arr_d = 0;
#pragma unroll
      for (int d = 0; d < 3; d++) {
        if (j + d > p.T - 1) break;
        const int hd = h0 + (j - jb) + d;
        const int dslot = hd % p.ndy;
        if (j == jb || d == 2) mbar_wait_tag(&dfull[dslot], (uint32_t)((hd / p.ndy) & 1), 65);
        const uint32_t dbase = dring_u32 + (uint32_t)dslot * (uint32_t)p.dy_bytes + dytoff;
#pragma unroll
        for (int lo = 0; lo < 2; lo++)
#pragma unroll
          for (int co = 0; co < 4; co++) {
            const f2_t dyv = ps_lds(dbase + ps_dyoff<TR>(lo, co, dyroww));
            if (d == 0) accb = ps_fma(dyv, one2, accb);
#pragma unroll
            for (int k1 = 0; k1 < 3; k1++)
#pragma unroll
              for (int k2 = 0; k2 < 3; k2++)
                acc[(2 - d) * 9 + k1 * 3 + k2] = ps_fma(dyv, xr[lo + k1][co + k2], acc[(2 - d) * 9 + k1 * 3 + k2]);
          }
        if (d == 0) {
          // the x plane is in registers and the gradient tile of plane j has been consumed: release both now (acc[18..26] are
          // fed by the last loads of either), look at the counts after the rest of the step
          const uint32_t dep = ps_lo32(acc[18]) ^ ps_lo32(acc[26]);
          __syncwarp();
          if (lane == 0) {
            arr_x = ps_release(&cnt[slot], dep, rt_zero);
            arr_d = ps_release(&dcnt[hd % p.ndy], dep, rt_zero);
          }
        }
      }
      if (lane == 0) {
        const int hd0 = h0 + (j - jb);
        if (ps_last(arr_x, use, n_warps)) ps_issue_x<2, TR>(p, seq, g + p.ns, ring, full, &tx);
        if (ps_last(arr_d, hd0 / p.ndy, n_warps)) ps_issue_dy<TR>(p, seq, hd0 + p.ndy, dring, dfull, &tdy);
      }
      if (++slot == p.ns) {
        slot = 0;
        use++;
      }
    }
    // tiles loaded ahead for the planes after the range (their last reads fed the accumulators of the last step)
    __syncwarp();
    for (int pl = je; pl <= last_dy; pl++) release_dy(h0 + (pl - jb), ps_lo32(acc[8]) ^ ps_lo32(acc[17]));
    h0 += last_dy - jb + 1;
    s += je - jb;
  }
  if (cur_cb >= 0) flush(cur_cb);
}

int g_peg_variant = 0;   // 0: plane-streaming kernels when supported, 1: always the peg.cu kernels (A/B timing, tests)

struct PsPlan {
  PsParams p;
  int grid, threads;
  size_t smem;
};

bool ps_geometry(const ctclip_peg_args* a, bool wgrad, PsPlan* plan) {
  if (g_peg_variant == 1) return false;
  if (a->D % PS_CB != 0 || a->W > 4 * PS_MAXCG) return false;
  if (a->temporal && !(a->T == a->H && a->H == a->W)) return false;
  PsParams& p = plan->p;
  p.B = a->B; p.T = a->T; p.H = a->H; p.W = a->W; p.D = a->D;
  p.n_cg = (a->W + 3) / 4;
  p.n_a1t = (a->H + PS_A1T - 1) / PS_A1T;
  p.n_cb = a->D / PS_CB;
  const long long vol = (long long)a->T * a->H * a->W * a->D;
  if (vol >= (1ll << 29)) return false;   // per-volume offsets are kept in 32 bits
  p.vol = vol;
  if (!a->temporal) {
    p.s0 = a->H * a->W * a->D; p.s1 = a->W * a->D; p.s2 = a->D;
  } else {   // conv grid (a0, a1, a2) = canonical (ih, iw, it)
    p.s0 = a->W * a->D; p.s1 = a->D; p.s2 = a->H * a->W * a->D;
  }
  p.slot_bytes = (PS_A1T + 2) * (4 * p.n_cg + 2) * 128;
  p.dy_bytes = PS_A1T * (4 * p.n_cg) * 128;
  const size_t budget = 227 * 1024 - 128;
  if (!wgrad) {
    const size_t fixed = PS_MAXSLOTS * 8 + PS_MAXSLOTS * 4;
    int ns = (int)((budget - fixed) / p.slot_bytes);
    if (ns > 6) ns = 6;
    if (ns < 3) return false;
    p.ns = ns; p.ndy = 0;
    plan->smem = (size_t)ns * p.slot_bytes + fixed + 128;
  } else {
    const size_t fixed = 2 * PS_MAXSLOTS * 8 + 2 * PS_MAXSLOTS * 4 + 28 * PS_CB * 4;
    int ns = 3, ndy = 5;
    while (ns * (size_t)p.slot_bytes + ndy * (size_t)p.dy_bytes + fixed > budget) {
      if (ndy > 4) ndy--;
      else if (ns > 2) ns--;
      else return false;
    }
    p.ns = ns; p.ndy = ndy;
    plan->smem = (size_t)ns * p.slot_bytes + (size_t)ndy * p.dy_bytes + fixed + 128;
  }
  const long long total = (long long)p.n_cb * p.n_a1t * a->B * a->T;
  if (total >= (1ll << 31)) return false;
  long long ctas = num_sms();
  if (ctas > total) ctas = total;
  p.total = (int)total;
  p.steps_per_cta = (int)((total + ctas - 1) / ctas);
  plan->grid = (int)((total + p.steps_per_cta - 1) / p.steps_per_cta);
  plan->threads = 64 * p.n_cg;
  return true;
}

int ps_tensor_map(CUtensorMap* m, const ctclip_peg_args* a, const float* base, int box_a2, int box_a1) {
  const uint64_t D = a->D, T = a->T, H = a->H, W = a->W;
  // canonical tensor [b][t][h][w][D]: dims (c, w, h, t, b) with monotonic strides in both stacks
  uint64_t dims[5] = {D, W, H, T, (uint64_t)a->B};
  uint64_t strides[4] = {D * 4, W * D * 4, H * W * D * 4, T * H * W * D * 4};
  uint32_t box[5] = {(uint32_t)PS_CB, (uint32_t)box_a2, (uint32_t)box_a1, 1u, 1u};   // spatial: (a2, a1, a0) = (w, h, t)
  if (a->temporal) {   // (a1, a0, a2) = (w, h, t)  (T == H == W)
    box[1] = (uint32_t)box_a1; box[2] = 1u; box[3] = (uint32_t)box_a2;
  }
  return encode_tmap_nd(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
}

}  // namespace

bool peg_stream_supported(const ctclip_peg_args* a, bool wgrad) {
  PsPlan plan;
  return ps_geometry(a, wgrad, &plan);
}

int peg_stream_launch_conv(int mode, const ctclip_peg_args* a, cudaStream_t stream) {
  PsPlan plan;
  if (!ps_geometry(a, false, &plan)) {
    set_error("peg_stream: unsupported geometry");
    return CTCLIP_ERR_UNSUPPORTED;
  }
  PsParams& p = plan.p;
  p.y = a->y;
  p.ybf = reinterpret_cast<__nv_bfloat16*>(a->y_bf16);
  p.weight = a->weight;
  p.bias = a->bias;
  p.dweight = nullptr;
  p.dbias = nullptr;
  CUtensorMap tx;
  if (int rc = ps_tensor_map(&tx, a, a->x, 4 * p.n_cg + 2, PS_A1T + 2)) return rc;
#define PS_LAUNCH_CONV(M, TRV)                                                                                              \
  do {                                                                                                                      \
    CTB_CUDA(cudaFuncSetAttribute(peg_stream_conv_kernel<M, TRV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem)); \
    peg_stream_conv_kernel<M, TRV><<<plan.grid, plan.threads, plan.smem, stream>>>(tx, p);                                   \
  } while (0)
  if (mode == 0 && !a->temporal) PS_LAUNCH_CONV(0, false);
  else if (mode == 0) PS_LAUNCH_CONV(0, true);
  else if (!a->temporal) PS_LAUNCH_CONV(1, false);
  else PS_LAUNCH_CONV(1, true);
#undef PS_LAUNCH_CONV
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}

int peg_stream_launch_wgrad(const ctclip_peg_args* a, cudaStream_t stream) {
  PsPlan plan;
  if (!ps_geometry(a, true, &plan)) {
    set_error("peg_stream: unsupported geometry");
    return CTCLIP_ERR_UNSUPPORTED;
  }
  PsParams& p = plan.p;
  p.y = nullptr;
  p.ybf = nullptr;
  p.weight = nullptr;
  p.bias = nullptr;
  p.dweight = a->dweight;
  p.dbias = a->dbias;
  CUtensorMap tx, tdy;
  if (int rc = ps_tensor_map(&tx, a, a->x, 4 * p.n_cg + 2, PS_A1T + 2)) return rc;
  if (int rc = ps_tensor_map(&tdy, a, a->dy, 4 * p.n_cg, PS_A1T)) return rc;
  if (!a->temporal) {
    CTB_CUDA(cudaFuncSetAttribute(peg_stream_wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
    peg_stream_wgrad_kernel<false><<<plan.grid, plan.threads, plan.smem, stream>>>(tx, tdy, p);
  } else {
    CTB_CUDA(cudaFuncSetAttribute(peg_stream_wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
    peg_stream_wgrad_kernel<true><<<plan.grid, plan.threads, plan.smem, stream>>>(tx, tdy, p);
  }
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}

}  // namespace ctb

extern "C" int ctclip_debug_set_peg_variant(int variant) {
  if (variant != 0 && variant != 1) {
    ctb::set_error("ctclip_debug_set_peg_variant: 0 (plane-streaming when supported) or 1 (general kernels)");
    return CTCLIP_ERR_ARG;
  }
  ctb::g_peg_variant = variant;
  return CTCLIP_OK;
}
