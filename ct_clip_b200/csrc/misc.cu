// Small / bandwidth-bound kernels around the GEMM and attention cores:
//   * fp32 CUDA-core GEMM for the tiny continuous-position-bias MLP (attention.py:257-276, "fp32 forced")
//   * CPB table -> per-head bias expansion / gradient reduction (2209 distinct offsets instead of S^2 rows)
//   * GEGLU backward (attention.py:39-42), column sums (bias gradients), casts
//   * vector-quantiser pieces (vector_quantize_pytorch 1.1.2 CosineSimCodebook): code-book l2norm,
//     gather + temporal mean pool (ct_clip.py:724), EMA statistics and update
//   * weight preparation: LayerNorm-affine folding, GEGLU column interleave, zero padding, and the
//     matching un-folding of weight gradients
#include "common.cuh"
#include "ptx.cuh"
#include "rng.cuh"
#include "../../include/ctclip_b200.h"

namespace ctb {

// ------------------------------------------------------------------------------------------------
// fp32 GEMM (tiny problems only): C[m,n] = epi( sum_k opA(m,k) * opB(k,n) )
// ------------------------------------------------------------------------------------------------
// 64 x 64 tile per CTA, 4 x 4 outputs per thread, K tiles of 16: two 16-byte shared loads per 16 FMAs (the former 32 x 32 tile
// with 2 x 2 outputs per thread did one 4-byte shared load per FMA: ~90 us for the 2209 x 512 x 512 layers of the CPB MLP).
__global__ void __launch_bounds__(256) sgemm_small_kernel(ctclip_sgemm_args a) {
  __shared__ __align__(16) float sA[16][68];   // [k][m]
  __shared__ __align__(16) float sB[16][68];   // [k][n]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < a.K; k0 += 16) {
    for (int i = threadIdx.x; i < 1024; i += 256) {
      // A tile: element (m, k); consecutive threads walk the contiguous direction of the operand
      const int am = a.trans_a ? (i & 63) : (i >> 4), ak = a.trans_a ? (i >> 6) : (i & 15);
      float va = 0.f;
      if (m0 + am < a.M && k0 + ak < a.K)
        va = a.trans_a ? a.A[(long long)(k0 + ak) * a.lda + m0 + am] : a.A[(long long)(m0 + am) * a.lda + k0 + ak];
      sA[ak][am] = va;
      const int bn = a.trans_b ? (i >> 4) : (i & 63), bk = a.trans_b ? (i & 15) : (i >> 6);
      float vb = 0.f;
      if (k0 + bk < a.K && n0 + bn < a.N)
        vb = a.trans_b ? a.B[(long long)(n0 + bn) * a.ldb + k0 + bk] : a.B[(long long)(k0 + bk) * a.ldb + n0 + bn];
      sB[bk][bn] = vb;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const float4 av = *reinterpret_cast<const float4*>(&sA[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&sB[k][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m >= a.M || n >= a.N) continue;
      float v = acc[i][j];
      if (a.bias != nullptr) v += a.bias[n];
      if (a.act == 1) v = v > 0.f ? v : 0.1f * v;                       // LeakyReLU(0.1), attention.py:17
      if (a.mask_ref != nullptr) v *= (a.mask_ref[(long long)m * a.ld_mask + n] > 0.f) ? 1.f : 0.1f;
      float* dst = a.C + (long long)m * a.ldc + n;
      *dst = a.accumulate ? (*dst + v) : v;
    }
}

// column sums: out[n] (+)= sum_m X[m,n]
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, long long ld, long long M, int N,
                                                    float* __restrict__ out, int rows_per_cta) {
  const int n = blockIdx.x * 256 + threadIdx.x;
  const long long m0 = (long long)blockIdx.y * rows_per_cta;
  const long long m1 = (m0 + rows_per_cta < M) ? m0 + rows_per_cta : M;
  if (n >= N) return;
  float s = 0.f;
  for (long long m = m0; m < m1; m++) s += (float)x[m * ld + n];
  atomicAdd(out + n, s);
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    reinterpret_cast<uint2*>(y)[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

// ------------------------------------------------------------------------------------------------
// continuous position bias (attention.py:229-276), evaluated on the (2h-1)(2w-1) distinct offsets
// ------------------------------------------------------------------------------------------------
__global__ void cpb_inputs_kernel(float* __restrict__ X, int h, int w) {
  const int R = (2 * h - 1) * (2 * w - 1);
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int dy = r / (2 * w - 1) - (h - 1), dx = r % (2 * w - 1) - (w - 1);
  const float fy = (float)dy, fx = (float)dx;
  X[2 * r] = (dy > 0 ? 1.f : (dy < 0 ? -1.f : 0.f)) * logf(fabsf(fy) + 1.f);   // sign(rel) * log(|rel| + 1)
  X[2 * r + 1] = (dx > 0 ? 1.f : (dx < 0 ? -1.f : 0.f)) * logf(fabsf(fx) + 1.f);
}
// bias[hd][i][j] = table[rel(i,j)][hd]; also the (i<->j) transposed copy used by the dK/dV kernel
__global__ void cpb_expand_kernel(const float* __restrict__ table, int heads, int h, int w,
                                  __nv_bfloat16* __restrict__ bias, __nv_bfloat16* __restrict__ bias_t) {
  const int n = h * w;
  const long long total = (long long)heads * n * n;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % n);
    const int i = (int)((idx / n) % n);
    const int hd = (int)(idx / ((long long)n * n));
    const int r = ((i / w) - (j / w) + h - 1) * (2 * w - 1) + ((i % w) - (j % w) + w - 1);
    const __nv_bfloat16 v = __float2bfloat16(table[(long long)r * heads + hd]);
    bias[idx] = v;
    if (bias_t != nullptr) bias_t[((long long)hd * n + j) * n + i] = v;
  }
}
// fragment-ordered bias tables: out[hd][rt][cb][lane][nt][e] = bias[hd][row][col] (transposed table: bias[hd][col][row])
// with row = rt*16 + lane/4 + 8*(e>>1), col = cb*64 + nt*8 + 2*(lane%4) + (e&1); zero outside the n x n grid
__global__ void cpb_expand_frag_kernel(const float* __restrict__ table, int heads, int h, int w,
                                       __nv_bfloat16* __restrict__ frag, __nv_bfloat16* __restrict__ frag_t) {
  const int n = h * w;
  const int n_pad = (n + 15) & ~15;
  const int RT = n_pad / 16, CBk = (n_pad + 63) / 64;
  const long long total = (long long)heads * RT * CBk * 32 * 32;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(idx & 3), nt = (int)((idx >> 2) & 7), lane = (int)((idx >> 5) & 31);
    long long rest = idx >> 10;
    const int cb = (int)(rest % CBk);
    rest /= CBk;
    const int rt = (int)(rest % RT);
    const int hd = (int)(rest / RT);
    const int row = rt * 16 + (lane >> 2) + 8 * (e >> 1);
    const int col = cb * 64 + nt * 8 + 2 * (lane & 3) + (e & 1);
    float v = 0.f, vt = 0.f;
    if (row < n && col < n) {
      const int r1 = ((row / w) - (col / w) + h - 1) * (2 * w - 1) + ((row % w) - (col % w) + w - 1);   // bias[row][col]
      const int r2 = ((col / w) - (row / w) + h - 1) * (2 * w - 1) + ((col % w) - (row % w) + w - 1);   // bias[col][row]
      v = table[(long long)r1 * heads + hd];
      vt = table[(long long)r2 * heads + hd];
    }
    // the attention kernels work in the log2 domain: store bias * log2(e) (one multiply less per score element)
    frag[idx] = __float2bfloat16(v * 1.4426950408889634f);
    frag_t[idx] = __float2bfloat16(vt * 1.4426950408889634f);
  }
}
// dtable[r][hd] = sum over (i,j) with rel(i,j) == r of dbias[hd][i][j]
// mirror = 1: the input is the TRANSPOSED table dbias_t[hd][j][i] (what the tcgen05 backward accumulates) and the result is
// ADDED to dtable: element (row a, col b) of the input is d bias[b][a], whose offset is rel(b, a) = R - 1 - rel(a, b).
__global__ void cpb_reduce_kernel(const float* __restrict__ dbias, int heads, int h, int w, float* __restrict__ dtable, int mirror) {
  const int r = blockIdx.x;
  const int n = h * w;
  const int dy = r / (2 * w - 1) - (h - 1), dx = r % (2 * w - 1) - (w - 1);
  const int iy0 = dy > 0 ? dy : 0, iy1 = dy < 0 ? h - 1 + dy : h - 1;
  const int ix0 = dx > 0 ? dx : 0, ix1 = dx < 0 ? w - 1 + dx : w - 1;
  const int ny = iy1 - iy0 + 1, nx = ix1 - ix0 + 1;
  __shared__ float red[256];
  for (int hd = 0; hd < heads; hd++) {
    float s = 0.f;
    for (int p = threadIdx.x; p < ny * nx; p += blockDim.x) {
      const int iy = iy0 + p / nx, ix = ix0 + p % nx;
      const int i = iy * w + ix, j = (iy - dy) * w + (ix - dx);
      s += dbias[((long long)hd * n + i) * n + j];
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      if (mirror) dtable[(long long)((2 * h - 1) * (2 * w - 1) - 1 - r) * heads + hd] += red[0];
      else dtable[(long long)r * heads + hd] = red[0];
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// GEGLU backward on the interleaved pre-activation h[m, 2j] = value, h[m, 2j+1] = gate:
//   dvalue = dg * gelu(gate),  dgate = dg * value * gelu'(gate);  written in place of h;  colsum -> s
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) geglu_bwd_kernel(const __nv_bfloat16* __restrict__ dg, long long ld_dg,
                                                       __nv_bfloat16* __restrict__ h, long long ld_h, long long M,
                                                       int n_pairs, float* __restrict__ colsum, int rows_per_cta) {
  // thread = (column group of 4 pairs = 8 h-columns, row lane)
  const int cg = blockIdx.x * 64 + (threadIdx.x & 63);
  const int rl = threadIdx.x >> 6;
  const long long m0 = (long long)blockIdx.y * rows_per_cta;
  const long long m1 = (m0 + rows_per_cta < M) ? m0 + rows_per_cta : M;
  float cs[8];
#pragma unroll
  for (int i = 0; i < 8; i++) cs[i] = 0.f;
  const bool ok = cg * 4 < n_pairs;
  if (ok) {
    // HBM-bound at ~70 % of the copy bandwidth whatever the arithmetic: neither the rcp/ex2 form of gelu / gelu' (instead of two
    // erff + expf) nor two rows in flight per thread changed the time (profiles/r2_stages_history.md)
    for (long long m = m0 + rl; m < m1; m += 4) {
      const uint2 ud = *reinterpret_cast<const uint2*>(dg + m * ld_dg + cg * 4);
      uint4 uh = *reinterpret_cast<const uint4*>(h + m * ld_h + cg * 8);
      const float2 d01 = unpack_bf16x2(ud.x), d23 = unpack_bf16x2(ud.y);
      const float dgv[4] = {d01.x, d01.y, d23.x, d23.y};
      uint32_t* ph = reinterpret_cast<uint32_t*>(&uh);
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const float2 xv = unpack_bf16x2(ph[i]);  // (value, gate)
        float ge, dge;
        gelu_erf_fast_both(xv.y, ge, dge);
        const float dval = dgv[i] * ge;
        const float dgate = dgv[i] * xv.x * dge;
        ph[i] = pack_bf16x2(dval, dgate);
        cs[2 * i] += dval;
        cs[2 * i + 1] += dgate;
      }
      *reinterpret_cast<uint4*>(h + m * ld_h + cg * 8) = uh;
    }
  }
  if (colsum == nullptr) return;
  __shared__ float red[4][64][8];
#pragma unroll
  for (int i = 0; i < 8; i++) red[rl][threadIdx.x & 63][i] = cs[i];
  __syncthreads();
  if (rl == 0 && ok) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const float t = red[0][threadIdx.x][i] + red[1][threadIdx.x][i] + red[2][threadIdx.x][i] + red[3][threadIdx.x][i];
      atomicAdd(colsum + cg * 8 + i, t);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// vector quantiser
// ------------------------------------------------------------------------------------------------
// row-wise l2 normalisation fp32 -> bf16 (code-book operand of the distance GEMM)
__global__ void l2norm_rows_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int rows, int D) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* xr = x + (long long)warp * D;
  float ss = 0.f;
  for (int c = lane; c < D; c += 32) ss += xr[c] * xr[c];
  const float inv = 1.f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
  for (int c = lane; c < D; c += 32) y[(long long)warp * D + c] = __float2bfloat16(xr[c] * inv);
}
// tokens[m, :] = embed[idx[m], :]
__global__ void vq_gather_kernel(const int* __restrict__ idx, const float* __restrict__ embed, float* __restrict__ out,
                                 long long M, int D) {
  const int d4 = D / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M * d4; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / d4;
    const int c = (int)(i % d4);
    reinterpret_cast<float4*>(out)[i] = reinterpret_cast<const float4*>(embed + (long long)idx[m] * D)[c];
  }
}
// pooled[b, s, :] = mean_t embed[idx[b, t, s], :]   (ct_clip.py:724 fused with the code-book gather)
__global__ void vq_gather_pool_kernel(const int* __restrict__ idx, const float* __restrict__ embed, int B, int T, int S,
                                      int D, float* __restrict__ pooled_f32, __nv_bfloat16* __restrict__ pooled_bf16) {
  const int d4 = D / 4;
  const long long total = (long long)B * S * d4;
  const float invT = 1.f / T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % d4);
    const long long bs = i / d4;
    const int s = (int)(bs % S);
    const int b = (int)(bs / S);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = 0; t < T; t++) {
      const int code = idx[((long long)b * T + t) * S + s];
      const float4 e = reinterpret_cast<const float4*>(embed + (long long)code * D)[c];
      acc.x += e.x; acc.y += e.y; acc.z += e.z; acc.w += e.w;
    }
    acc.x *= invT; acc.y *= invT; acc.z *= invT; acc.w *= invT;
    if (pooled_f32 != nullptr) reinterpret_cast<float4*>(pooled_f32)[i] = acc;
    if (pooled_bf16 != nullptr)
      reinterpret_cast<uint2*>(pooled_bf16)[i] = make_uint2(pack_bf16x2(acc.x, acc.y), pack_bf16x2(acc.z, acc.w));
  }
}
// dtokens[b, t, s, :] = dpooled[b, s, :] / T    (backward of the mean pool through the straight-through VQ)
__global__ void pool_bwd_kernel(const float* __restrict__ dpooled, int B, int T, int S, int D, float* __restrict__ dtok) {
  const int d4 = D / 4;
  const long long total = (long long)B * T * S * d4;
  const float invT = 1.f / T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % d4);
    const long long bts = i / d4;
    const int s = (int)(bts % S);
    const int b = (int)(bts / ((long long)T * S));
    float4 v = reinterpret_cast<const float4*>(dpooled)[((long long)b * S + s) * d4 + c];
    v.x *= invT; v.y *= invT; v.z *= invT; v.w *= invT;
    reinterpret_cast<float4*>(dtok)[i] = v;
  }
}
// EMA statistics: bins[code] += 1, embed_sum[code, :] += l2norm(x[m, :])
__global__ void vq_ema_accum_kernel(const float* __restrict__ x, const int* __restrict__ idx, long long M, int D,
                                    float* __restrict__ bins, float* __restrict__ embed_sum) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= M) return;
  const float* xr = x + warp * D;
  float ss = 0.f;
  for (int c = lane; c < D; c += 32) ss += xr[c] * xr[c];
  const float inv = 1.f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
  const int code = idx[warp];
  if (lane == 0) atomicAdd(bins + code, 1.f);
  for (int c = lane; c < D; c += 32) atomicAdd(embed_sum + (long long)code * D + c, xr[c] * inv);
}
// embed <- embed*decay + (1-decay) * (bins>0 ? l2norm(embed_sum/bins) : l2norm(embed)); cluster_size likewise
__global__ void vq_ema_update_kernel(float* __restrict__ embed, float* __restrict__ cluster_size,
                                     const float* __restrict__ bins, const float* __restrict__ embed_sum, int C, int D,
                                     float decay) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= C) return;
  const float b = bins[warp];
  float* er = embed + (long long)warp * D;
  const float* sr = embed_sum + (long long)warp * D;
  const bool zero = (b == 0.f);
  float ss = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float v = zero ? er[c] : sr[c] / b;
    ss += v * v;
  }
  const float inv = 1.f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
  for (int c = lane; c < D; c += 32) {
    const float v = (zero ? er[c] : sr[c] / b) * inv;
    er[c] = er[c] * decay + v * (1.f - decay);
  }
  if (lane == 0) cluster_size[warp] = cluster_size[warp] * decay + b * (1.f - decay);
}

// ------------------------------------------------------------------------------------------------
// weight preparation (fp32 master -> bf16 GEMM operand) and gradient un-folding
// ------------------------------------------------------------------------------------------------
// out[r, k] = rowmap[r] >= 0 && k < K ? W[rowmap[r], k] * (gamma ? gamma[k] : 1) : 0      out: bf16 [Np, Kp]
__global__ void prep_weight_kernel(const float* __restrict__ W, long long ldw, int K, const float* __restrict__ gamma,
                                   const int* __restrict__ rowmap, int Np, int Kp, __nv_bfloat16* __restrict__ out) {
  const long long total = (long long)Np * Kp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % Kp);
    const int r = (int)(i / Kp);
    const int src = rowmap ? rowmap[r] : r;
    float v = 0.f;
    if (src >= 0 && k < K) v = W[(long long)src * ldw + k] * (gamma ? gamma[k] : 1.f);
    out[i] = __float2bfloat16(v);
  }
}
// out[r] = sum_k W[rowmap[r], k] * beta[k] + (bias_in ? bias_in[rowmap[r]] : 0)
__global__ void prep_bias_kernel(const float* __restrict__ W, long long ldw, int K, const float* __restrict__ beta,
                                 const float* __restrict__ bias_in, const int* __restrict__ rowmap, int Np,
                                 float* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= Np) return;
  const int src = rowmap ? rowmap[warp] : warp;
  float s = 0.f;
  if (src >= 0 && beta != nullptr)
    for (int k = lane; k < K; k += 32) s += W[(long long)src * ldw + k] * beta[k];
  s = warp_sum(s);
  if (lane == 0) out[warp] = (src >= 0) ? s + (bias_in ? bias_in[src] : 0.f) : 0.f;
}
// Batched weight preparation: one launch walks a DEVICE table of descriptors (blockIdx.y = descriptor). kind 0 = operand
// (prep_weight_kernel), kind 1 = bias (prep_bias_kernel). 193 separate launches of 5-12 us each cost 2.3 ms per optimiser
// step at configs[1]; the data is 150 MB (25 us).
__global__ void __launch_bounds__(256) prep_batched_kernel(const ctclip_prep_desc* __restrict__ descs) {
  const ctclip_prep_desc d = descs[blockIdx.y];
  if (d.kind == 0) {
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(d.out);
    const long long total = (long long)d.Np * d.Kp;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      const int k = (int)(i % d.Kp);
      const int r = (int)(i / d.Kp);
      const int src = d.rowmap ? d.rowmap[r] : r;
      float v = 0.f;
      if (src >= 0 && k < d.K) v = d.W[(long long)src * d.ldw + k] * (d.gamma ? d.gamma[k] : 1.f);
      out[i] = __float2bfloat16(v);
    }
  } else {
    float* out = reinterpret_cast<float*>(d.out);
    const int lane = threadIdx.x & 31;
    for (int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < d.Np; row += gridDim.x * (blockDim.x >> 5)) {
      const int src = d.rowmap ? d.rowmap[row] : row;
      float s = 0.f;
      if (src >= 0 && d.beta != nullptr)
        for (int k = lane; k < d.K; k += 32) s += d.W[(long long)src * d.ldw + k] * d.beta[k];
      s = warp_sum(s);
      if (lane == 0) out[row] = (src >= 0) ? s + (d.bias_in ? d.bias_in[src] : 0.f) : 0.f;
    }
  }
}
// Given G = dL/dW' (W' = prepared weight, fp32 [Np, ldg]) and s[r] = dL/db'[r]:
//   dW[src, k] += G[r, k] * gamma[k];  dgamma[k] += sum_r W[src,k] * G[r,k];  dbeta[k] += sum_r W[src,k] * s[r]
//   dbias[src] += s[r]
__global__ void __launch_bounds__(128) unprep_wgrad_kernel(const float* __restrict__ G, long long ldg,
                                                          const float* __restrict__ W, long long ldw, int K,
                                                          const float* __restrict__ gamma, const int* __restrict__ rowmap,
                                                          int Np, const float* __restrict__ s, float* __restrict__ dW,
                                                          float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                          float* __restrict__ dbias, int rows_per_cta) {
  const int k = blockIdx.x * 128 + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_cta;
  const int r1 = (r0 + rows_per_cta < Np) ? r0 + rows_per_cta : Np;
  if (dbias != nullptr && blockIdx.x == 0 && s != nullptr) {
    for (int r = r0 + threadIdx.x; r < r1; r += 128) {
      const int src = rowmap ? rowmap[r] : r;
      if (src >= 0) atomicAdd(dbias + src, s[r]);
    }
  }
  if (k >= K) return;
  const float gm = gamma ? gamma[k] : 1.f;
  float ag = 0.f, ab = 0.f;
  for (int r = r0; r < r1; r++) {
    const int src = rowmap ? rowmap[r] : r;
    if (src < 0) continue;
    const float g = G[(long long)r * ldg + k];
    const float w = W[(long long)src * ldw + k];
    dW[(long long)src * ldw + k] += g * gm;
    ag += w * g;
    if (s != nullptr) ab += w * s[r];
  }
  if (dgamma != nullptr) atomicAdd(dgamma + k, ag);
  if (dbeta != nullptr && s != nullptr) atomicAdd(dbeta + k, ab);
}


// ------------------------------------------------------------------------------------------------
// BERT text tower helpers (transformers BertEmbeddings / BertIntermediate)
// ------------------------------------------------------------------------------------------------
// out[row, :] = word[ids[row], :] + pos[row % n, :] + type0[:]
__global__ void bert_embed_kernel(const long long* __restrict__ ids, const float* __restrict__ word,
                                  const float* __restrict__ pos, const float* __restrict__ type0, float* __restrict__ out,
                                  long long rows, int n, int H) {
  const int h4 = H / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows * h4; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / h4;
    const int c = (int)(i % h4);
    const float4 w = reinterpret_cast<const float4*>(word + ids[row] * H)[c];
    const float4 p = reinterpret_cast<const float4*>(pos + (row % n) * H)[c];
    const float4 t = reinterpret_cast<const float4*>(type0)[c];
    reinterpret_cast<float4*>(out)[i] = make_float4(w.x + p.x + t.x, w.y + p.y + t.y, w.z + p.z + t.z, w.w + p.w + t.w);
  }
}
// dword[ids[row]] += g[row], dpos[row % n] += g[row]   (dtype0 = column sum of g, done with ctclip_colsum)
__global__ void bert_embed_bwd_kernel(const long long* __restrict__ ids, const float* __restrict__ g, float* __restrict__ dword,
                                      float* __restrict__ dpos, long long rows, int n, int H) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows * H; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / H;
    const int c = (int)(i % H);
    const float v = g[i];
    atomicAdd(dword + ids[row] * H + c, v);
    atomicAdd(dpos + (row % n) * H + c, v);
  }
}
// dpre = dy * gelu'(pre)  (bf16, written over dy); colsum[N] += column sums (bias gradient)
__global__ void __launch_bounds__(256) gelu_bwd_kernel(__nv_bfloat16* __restrict__ dy, long long ld_dy,
                                                      const __nv_bfloat16* __restrict__ pre, long long ld_pre, long long M,
                                                      int N, float* __restrict__ colsum, int rows_per_cta) {
  const int cg = blockIdx.x * 64 + (threadIdx.x & 63);
  const int rl = threadIdx.x >> 6;
  const long long m0 = (long long)blockIdx.y * rows_per_cta;
  const long long m1 = (m0 + rows_per_cta < M) ? m0 + rows_per_cta : M;
  float cs[8];
#pragma unroll
  for (int i = 0; i < 8; i++) cs[i] = 0.f;
  const bool ok = cg * 8 < N;
  if (ok) {
    for (long long m = m0 + rl; m < m1; m += 4) {
      uint4 ud = *reinterpret_cast<const uint4*>(dy + m * ld_dy + cg * 8);
      const uint4 up = *reinterpret_cast<const uint4*>(pre + m * ld_pre + cg * 8);
      uint32_t* pd = reinterpret_cast<uint32_t*>(&ud);
      const uint32_t* pp = reinterpret_cast<const uint32_t*>(&up);
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const float2 d = unpack_bf16x2(pd[i]), x = unpack_bf16x2(pp[i]);
        float g0, g1, dg0, dg1;
        gelu_erf_fast_both(x.x, g0, dg0);
        gelu_erf_fast_both(x.y, g1, dg1);
        const float a = d.x * dg0, b = d.y * dg1;
        pd[i] = pack_bf16x2(a, b);
        cs[2 * i] += a;
        cs[2 * i + 1] += b;
      }
      *reinterpret_cast<uint4*>(dy + m * ld_dy + cg * 8) = ud;
    }
  }
  if (colsum == nullptr) return;
  __shared__ float red[4][64][8];
#pragma unroll
  for (int i = 0; i < 8; i++) red[rl][threadIdx.x & 63][i] = cs[i];
  __syncthreads();
  if (rl == 0 && ok) {
#pragma unroll
    for (int i = 0; i < 8; i++)
      atomicAdd(colsum + cg * 8 + i, red[0][threadIdx.x][i] + red[1][threadIdx.x][i] + red[2][threadIdx.x][i] + red[3][threadIdx.x][i]);
  }
}
// zero-shot head (scripts/zero_shot.py:140-143): probs[v, p] = softmax([s(v,2p), s(v,2p+1)])[0], s = img . text * exp(T)
__global__ void zero_shot_probs_kernel(const float* __restrict__ img, const float* __restrict__ txt, int V, int P2, int L,
                                       const float* __restrict__ temperature, float* __restrict__ probs) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= V * (P2 / 2)) return;
  const int v = warp / (P2 / 2), p = warp % (P2 / 2);
  float s0 = 0.f, s1 = 0.f;
  for (int c = lane; c < L; c += 32) {
    const float x = img[(long long)v * L + c];
    s0 += x * txt[(long long)(2 * p) * L + c];
    s1 += x * txt[(long long)(2 * p + 1) * L + c];
  }
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  if (lane == 0) {
    const float t = expf(temperature[0]);
    s0 *= t;
    s1 *= t;
    const float m = fmaxf(s0, s1);
    const float e0 = expf(s0 - m), e1 = expf(s1 - m);
    probs[warp] = e0 / (e0 + e1);
  }
}

// warp per token: fp32 scores of the two bf16-GEMM candidates, x . e / ||e|| (the token norm is a common positive factor)
__global__ void __launch_bounds__(256) vq_rerank_kernel(const float* __restrict__ x, const float* __restrict__ embed,
                                                        int* __restrict__ idx, const int* __restrict__ idx2, long long M, int D) {
  const long long m = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (m >= M) return;
  const int c1 = idx[m], c2 = idx2[m];
  const float4* xr = reinterpret_cast<const float4*>(x + m * D);
  const float4* e1 = reinterpret_cast<const float4*>(embed + (long long)c1 * D);
  const float4* e2 = reinterpret_cast<const float4*>(embed + (long long)c2 * D);
  float d1 = 0.f, d2 = 0.f, n1 = 0.f, n2 = 0.f;
  for (int i = lane; i < D / 4; i += 32) {
    const float4 a = xr[i], b = __ldg(e1 + i), c = __ldg(e2 + i);
    d1 += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
    d2 += a.x * c.x + a.y * c.y + a.z * c.z + a.w * c.w;
    n1 += b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
    n2 += c.x * c.x + c.y * c.y + c.z * c.z + c.w * c.w;
  }
  d1 = warp_sum(d1); d2 = warp_sum(d2); n1 = warp_sum(n1); n2 = warp_sum(n2);
  if (lane == 0) {
    const float s1 = d1 / fmaxf(sqrtf(n1), 1e-12f), s2 = d2 / fmaxf(sqrtf(n2), 1e-12f);
    if (s2 > s1 || (s2 == s1 && c2 < c1)) idx[m] = c2;
  }
}

// ------------------------------------------------------------------------------------------------
// Dropout (HF BertModel: embeddings, BertSelfOutput, BertOutput -- modeling_bert.py `self.dropout(hidden_states)`): masks are
// regenerated from (seed, offset, element index) with Philox (rng.cuh); 4 elements per thread = one Philox call.
//   forward : y = resid + keep * x / (1 - p)          (resid optional; y_f32 / y_bf16 optional outputs)
//   backward: dx = keep * dy / (1 - p)                (dx_f32 / dx_bf16 optional outputs)   -- the same kernel with resid = NULL
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dropout_kernel(const float* __restrict__ x, const float* __restrict__ resid,
                                                      float* __restrict__ y_f32, __nv_bfloat16* __restrict__ y_bf16, long long n,
                                                      float inv_keep, uint32_t thresh, unsigned long long seed,
                                                      unsigned long long offset) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // group of 4 elements
  const long long i0 = g * 4;
  if (i0 >= n) return;
  uint32_t w[4];
  philox4(seed, offset + (unsigned long long)g, w);
  float v[4];
  if (i0 + 3 < n) {
    const float4 xv = *reinterpret_cast<const float4*>(x + i0);
    v[0] = xv.x; v[1] = xv.y; v[2] = xv.z; v[3] = xv.w;
  } else {
    for (int e = 0; e < 4; e++) v[e] = (i0 + e < n) ? x[i0 + e] : 0.f;
  }
#pragma unroll
  for (int e = 0; e < 4; e++) {
    v[e] = (w[e] >= thresh) ? v[e] * inv_keep : 0.f;
    if (resid != nullptr && i0 + e < n) v[e] += resid[i0 + e];
  }
  if (i0 + 3 < n) {
    if (y_f32 != nullptr) *reinterpret_cast<float4*>(y_f32 + i0) = make_float4(v[0], v[1], v[2], v[3]);
    if (y_bf16 != nullptr) *reinterpret_cast<uint2*>(y_bf16 + i0) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
  } else {
    for (int e = 0; e < 4 && i0 + e < n; e++) {
      if (y_f32 != nullptr) y_f32[i0 + e] = v[e];
      if (y_bf16 != nullptr) y_bf16[i0 + e] = __float2bfloat16(v[e]);
    }
  }
}

}  // namespace ctb

using namespace ctb;

static inline int grid_for(long long n, int block, int per_sm = 16) {
  long long g = (n + block - 1) / block;
  const long long cap = (long long)num_sms() * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

extern "C" int ctclip_sgemm_f32(const ctclip_sgemm_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(a && a->A && a->B && a->C && a->M > 0 && a->N > 0 && a->K > 0, "sgemm: bad args");
  dim3 grid(ceil_div(a->N, 64), ceil_div(a->M, 64));
  sgemm_small_kernel<<<grid, 256, 0, stream>>>(*a);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}

extern "C" int ctclip_colsum(const void* x, int32_t is_bf16, int64_t ld, int64_t M, int32_t N, float* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(x && out && M > 0 && N > 0, "colsum: bad args");
  const int rows_per_cta = M >= 16384 ? 512 : 32;   // few rows: spread them over more CTAs (latency-bound otherwise)
  dim3 grid(ceil_div(N, 256), ceil_div(M, rows_per_cta));
  if (is_bf16) colsum_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), ld, M, N, out, rows_per_cta);
  else colsum_kernel<float><<<grid, 256, 0, stream>>>(reinterpret_cast<const float*>(x), ld, M, N, out, rows_per_cta);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}

extern "C" int ctclip_cast_f32_bf16(const float* x, void* y, int64_t n, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(x && y && n > 0 && n % 4 == 0, "cast: n must be a positive multiple of 4");
  cast_f32_bf16_kernel<<<grid_for(n / 4, 256), 256, 0, stream>>>(x, reinterpret_cast<__nv_bfloat16*>(y), n / 4);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}

extern "C" int ctclip_cpb_inputs(float* X, int32_t h, int32_t w, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(X && h > 0 && w > 0, "cpb_inputs: bad args");
  const int R = (2 * h - 1) * (2 * w - 1);
  cpb_inputs_kernel<<<ceil_div(R, 256), 256, 0, stream>>>(X, h, w);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}
extern "C" int ctclip_cpb_expand(const float* table, int32_t heads, int32_t h, int32_t w, void* bias, void* bias_t,
                                 void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(table && bias && heads > 0 && h > 0 && w > 0, "cpb_expand: bad args");
  const long long total = (long long)heads * h * w * h * w;
  cpb_expand_kernel<<<grid_for(total, 256), 256, 0, stream>>>(table, heads, h, w, reinterpret_cast<__nv_bfloat16*>(bias),
                                                            reinterpret_cast<__nv_bfloat16*>(bias_t));
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}
extern "C" int ctclip_cpb_expand_frag(const float* table, int32_t heads, int32_t h, int32_t w, void* bias_frag, void* bias_t_frag,
                                      void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(table && bias_frag && bias_t_frag && heads > 0 && h > 0 && w > 0, "cpb_expand_frag: bad args");
  const int n_pad = (h * w + 15) & ~15;
  const long long total = (long long)heads * (n_pad / 16) * ((n_pad + 63) / 64) * 1024;
  cpb_expand_frag_kernel<<<grid_for(total, 256), 256, 0, stream>>>(table, heads, h, w, reinterpret_cast<__nv_bfloat16*>(bias_frag),
                                                                  reinterpret_cast<__nv_bfloat16*>(bias_t_frag));
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}
extern "C" int ctclip_cpb_reduce(const float* dbias, int32_t heads, int32_t h, int32_t w, float* dtable, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(dbias && dtable && heads > 0 && h > 0 && w > 0, "cpb_reduce: bad args");
  cpb_reduce_kernel<<<(2 * h - 1) * (2 * w - 1), 256, 0, stream>>>(dbias, heads, h, w, dtable, 0);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}
extern "C" int ctclip_cpb_reduce_t(const float* dbias_t, int32_t heads, int32_t h, int32_t w, float* dtable, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(dbias_t && dtable && heads > 0 && h > 0 && w > 0, "cpb_reduce_t: bad args");
  cpb_reduce_kernel<<<(2 * h - 1) * (2 * w - 1), 256, 0, stream>>>(dbias_t, heads, h, w, dtable, 1);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}

extern "C" int ctclip_geglu_bwd(const void* dg, int64_t ld_dg, void* h, int64_t ld_h, int64_t M, int32_t n_pairs,
                                float* colsum, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(dg && h && M > 0 && n_pairs > 0 && n_pairs % 4 == 0, "geglu_bwd: n_pairs must be a multiple of 4");
  CTB_CHECK_ARG(ld_dg % 4 == 0 && ld_h % 8 == 0, "geglu_bwd: bad leading dimensions");
  const int rows_per_cta = 256;
  dim3 grid(ceil_div(n_pairs / 4, 64), ceil_div(M, rows_per_cta));
  geglu_bwd_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(dg), ld_dg,
                                            reinterpret_cast<__nv_bfloat16*>(h), ld_h, M, n_pairs, colsum, rows_per_cta);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}

extern "C" int ctclip_l2norm_rows_bf16(const float* x, void* y, int32_t rows, int32_t D, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(x && y && rows > 0 && D > 0, "l2norm_rows: bad args");
  l2norm_rows_bf16_kernel<<<ceil_div((long long)rows * 32, 256), 256, 0, stream>>>(x, reinterpret_cast<__nv_bfloat16*>(y), rows, D);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}
extern "C" int ctclip_dropout(const float* x, const float* resid, float* y_f32, void* y_bf16, int64_t n, float p, uint64_t seed,
                              uint64_t offset, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(x && (y_f32 || y_bf16) && n > 0 && p >= 0.f && p < 1.f, "dropout: bad args");
  CTB_CHECK_ARG((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(resid) | reinterpret_cast<uintptr_t>(y_f32)) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(y_bf16) % 8 == 0, "dropout: pointers must be 16-byte aligned");
  const long long groups = (n + 3) / 4;
  dropout_kernel<<<ceil_div(groups, 256), 256, 0, stream>>>(x, resid, y_f32, reinterpret_cast<__nv_bfloat16*>(y_bf16), n, 1.f / (1.f - p),
                                                          dropout_threshold(p), seed, offset);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}
extern "C" int ctclip_vq_rerank(const float* x, const float* embed, int32_t* idx, const int32_t* idx2, int64_t M, int32_t D,
                                void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(x && embed && idx && idx2 && M > 0 && D > 0 && D % 4 == 0, "vq_rerank: bad args");
  vq_rerank_kernel<<<ceil_div(M * 32, 256), 256, 0, stream>>>(x, embed, idx, idx2, M, D);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}
extern "C" int ctclip_vq_gather(const int32_t* idx, const float* embed, float* out, int64_t M, int32_t D, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(idx && embed && out && M > 0 && D % 4 == 0, "vq_gather: bad args");
  vq_gather_kernel<<<grid_for(M * (D / 4), 256), 256, 0, stream>>>(idx, embed, out, M, D);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}
extern "C" int ctclip_vq_gather_pool(const int32_t* idx, const float* embed, int32_t B, int32_t T, int32_t S, int32_t D,
                                     float* pooled_f32, void* pooled_bf16, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(idx && embed && (pooled_f32 || pooled_bf16) && B > 0 && T > 0 && S > 0 && D % 4 == 0, "vq_gather_pool: bad args");
  vq_gather_pool_kernel<<<grid_for((long long)B * S * (D / 4), 256), 256, 0, stream>>>(
      idx, embed, B, T, S, D, pooled_f32, reinterpret_cast<__nv_bfloat16*>(pooled_bf16));
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}
extern "C" int ctclip_pool_bwd(const float* dpooled, int32_t B, int32_t T, int32_t S, int32_t D, float* dtok, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(dpooled && dtok && B > 0 && T > 0 && S > 0 && D % 4 == 0, "pool_bwd: bad args");
  pool_bwd_kernel<<<grid_for((long long)B * T * S * (D / 4), 256), 256, 0, stream>>>(dpooled, B, T, S, D, dtok);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}
extern "C" int ctclip_vq_ema_accum(const float* x, const int32_t* idx, int64_t M, int32_t D, float* bins, float* embed_sum,
                                   void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(x && idx && bins && embed_sum && M > 0 && D > 0, "vq_ema_accum: bad args");
  vq_ema_accum_kernel<<<ceil_div(M * 32, 256), 256, 0, stream>>>(x, idx, M, D, bins, embed_sum);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}
extern "C" int ctclip_vq_ema_update(float* embed, float* cluster_size, const float* bins, const float* embed_sum, int32_t C,
                                    int32_t D, float decay, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(embed && cluster_size && bins && embed_sum && C > 0 && D > 0, "vq_ema_update: bad args");
  vq_ema_update_kernel<<<ceil_div((long long)C * 32, 256), 256, 0, stream>>>(embed, cluster_size, bins, embed_sum, C, D, decay);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}

extern "C" int ctclip_prep_weight(const float* W, int64_t ldw, int32_t K, const float* gamma, const int32_t* rowmap,
                                  int32_t Np, int32_t Kp, void* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(W && out && K > 0 && Np > 0 && Kp >= K, "prep_weight: bad args");
  prep_weight_kernel<<<grid_for((long long)Np * Kp, 256), 256, 0, stream>>>(W, ldw, K, gamma, rowmap, Np, Kp,
                                                                          reinterpret_cast<__nv_bfloat16*>(out));
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}
extern "C" int ctclip_prep_batched(const ctclip_prep_desc* descs_device, int32_t n, int32_t blocks_per_desc, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(descs_device && n > 0 && blocks_per_desc > 0, "prep_batched: bad args");
  prep_batched_kernel<<<dim3((unsigned)blocks_per_desc, (unsigned)n), 256, 0, stream>>>(descs_device);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}
extern "C" int ctclip_prep_bias(const float* W, int64_t ldw, int32_t K, const float* beta, const float* bias_in,
                                const int32_t* rowmap, int32_t Np, float* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(W && out && K > 0 && Np > 0, "prep_bias: bad args");
  prep_bias_kernel<<<ceil_div((long long)Np * 32, 256), 256, 0, stream>>>(W, ldw, K, beta, bias_in, rowmap, Np, out);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}
extern "C" int ctclip_unprep_wgrad(const float* G, int64_t ldg, const float* W, int64_t ldw, int32_t K, const float* gamma,
                                   const int32_t* rowmap, int32_t Np, const float* s, float* dW, float* dgamma,
                                   float* dbeta, float* dbias, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(G && W && dW && K > 0 && Np > 0, "unprep_wgrad: bad args");
  const int rows_per_cta = 16;   // 64 serial rows per thread left this kernel latency-bound (43 us for 17 MB)
  dim3 grid(ceil_div(K, 128), ceil_div(Np, rows_per_cta));
  unprep_wgrad_kernel<<<grid, 128, 0, stream>>>(G, ldg, W, ldw, K, gamma, rowmap, Np, s, dW, dgamma, dbeta, dbias, rows_per_cta);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}

extern "C" int ctclip_bert_embed(const int64_t* ids, const float* word, const float* pos, const float* type0, float* out,
                                 int64_t rows, int32_t n, int32_t H, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(ids && word && pos && type0 && out && rows > 0 && n > 0 && H % 4 == 0, "bert_embed: bad args");
  bert_embed_kernel<<<grid_for(rows * (H / 4), 256), 256, 0, stream>>>(reinterpret_cast<const long long*>(ids), word, pos,
                                                                     type0, out, rows, n, H);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}
extern "C" int ctclip_bert_embed_bwd(const int64_t* ids, const float* g, float* dword, float* dpos, int64_t rows, int32_t n,
                                     int32_t H, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(ids && g && dword && dpos && rows > 0 && n > 0 && H > 0, "bert_embed_bwd: bad args");
  bert_embed_bwd_kernel<<<grid_for(rows * H, 256), 256, 0, stream>>>(reinterpret_cast<const long long*>(ids), g, dword, dpos,
                                                                   rows, n, H);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}
extern "C" int ctclip_gelu_bwd(void* dy, int64_t ld_dy, const void* pre, int64_t ld_pre, int64_t M, int32_t N, float* colsum,
                               void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(dy && pre && M > 0 && N > 0 && N % 8 == 0 && ld_dy % 8 == 0 && ld_pre % 8 == 0, "gelu_bwd: bad args");
  const int rows_per_cta = 128;
  dim3 grid(ceil_div(N / 8, 64), ceil_div(M, rows_per_cta));
  gelu_bwd_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<__nv_bfloat16*>(dy), ld_dy,
                                           reinterpret_cast<const __nv_bfloat16*>(pre), ld_pre, M, N, colsum, rows_per_cta);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}
extern "C" int ctclip_zero_shot_probs(const float* img, const float* txt, int32_t V, int32_t P2, int32_t L,
                                      const float* temperature, float* probs, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(img && txt && temperature && probs && V > 0 && P2 > 0 && P2 % 2 == 0 && L > 0, "zero_shot_probs: bad args");
  zero_shot_probs_kernel<<<ceil_div((long long)V * (P2 / 2) * 32, 256), 256, 0, stream>>>(img, txt, V, P2, L, temperature, probs);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}
