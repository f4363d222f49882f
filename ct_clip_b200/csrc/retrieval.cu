// Retrieval over saved CT-CLIP latents: the evaluation scripts scripts/report_to_volume_new.py:47-63 and
// scripts/volume_to_volume_new.py:80-96 score every (query, gallery) pair in nested Python loops and sort each row with
// `sorted(enumerate(values), reverse=True)`. Here: scores = one fp32 GEMM (ctclip_sgemm_f32), then this kernel keeps the k best
// of every row (descending, the lower index first among equal scores -- what Python's stable sort with reverse=True yields).
// One CTA per query row; the row lives in shared memory (G <= 49152) and the k winners are extracted by k block-wide arg-max
// passes: O(k * G / 256) per row, far below the GEMM for the k <= 100 the scripts use.
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/ctclip_b200.h"

namespace ctb {

__global__ void __launch_bounds__(256) topk_rows_kernel(const float* __restrict__ scores, long long ld, int G, int k,
                                                       int* __restrict__ idx_out, float* __restrict__ val_out) {
  extern __shared__ float s_row[];
  __shared__ float s_v[8];
  __shared__ int s_i[8];
  const float* src = scores + (long long)blockIdx.x * ld;
  for (int i = threadIdx.x; i < G; i += 256) s_row[i] = src[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = 0; r < k; r++) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < G; i += 256) {
      const float v = s_row[i];
      if (v > bv) { bv = v; bi = i; }            // strided scan: the first hit of a thread is its lowest index
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { s_v[warp] = bv; s_i[warp] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float fv = s_v[0];
      int fi = s_i[0];
      for (int w = 1; w < 8; w++)
        if (s_v[w] > fv || (s_v[w] == fv && s_i[w] < fi)) { fv = s_v[w]; fi = s_i[w]; }
      if (fi == 0x7fffffff) fi = 0;              // fewer than k finite scores (all -inf): repeat index 0
      idx_out[(long long)blockIdx.x * k + r] = fi;
      if (val_out != nullptr) val_out[(long long)blockIdx.x * k + r] = fv;
      if (fi < G) s_row[fi] = -INFINITY;
    }
    __syncthreads();
  }
}

// rows <- rows / max(||row||, eps)   (cosine similarity of volume_to_volume_new.py:86-90 = dot of the normalised rows)
__global__ void __launch_bounds__(256) l2norm_rows_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int rows, int D) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  float ss = 0.f;
  for (int i = lane; i < D; i += 32) { const float v = x[(long long)row * D + i]; ss += v * v; }
  ss = warp_sum(ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  for (int i = lane; i < D; i += 32) y[(long long)row * D + i] = x[(long long)row * D + i] * inv;
}

}  // namespace ctb

using namespace ctb;

extern "C" int ctclip_topk_rows(const float* scores, int64_t ld, int32_t Q, int32_t G, int32_t k, int32_t* idx_out, float* val_out,
                                void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(scores && idx_out && Q > 0 && G > 0 && k > 0 && k <= G && ld >= G, "topk_rows: bad args");
  CTB_CHECK_ARG(G <= 49152, "topk_rows: gallery of %d rows exceeds the shared-memory row buffer (49152)", G);
  const size_t smem = (size_t)G * sizeof(float);
  CTB_CUDA(cudaFuncSetAttribute(topk_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  topk_rows_kernel<<<Q, 256, smem, stream>>>(scores, ld, G, k, idx_out, val_out);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}

extern "C" int ctclip_l2norm_rows_f32(const float* x, float* y, int32_t rows, int32_t D, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(x && y && rows > 0 && D > 0, "l2norm_rows_f32: bad args");
  l2norm_rows_f32_kernel<<<ceil_div(rows, 8), 256, 0, stream>>>(x, y, rows, D);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}
