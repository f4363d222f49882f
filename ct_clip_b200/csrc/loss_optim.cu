// (1) Fused contrastive head: l2norm of both latents, logits = t_hat . i_hat * exp(T), symmetric InfoNCE
//     in the reference's exp/log form (ct_clip.py:771, :796, :845-878, log(x) := log(x + 1e-20)), and the
//     complete backward (d latents, d temperature) in the same call. Latency-bound: B <= 256, L <= 1024.
// (2) Global-norm gradient clipping + Adam over one flat fp32 parameter arena
//     (CTCLIPTrainer.py:259-263, optimizer.py:23-24: Adam(betas=(0.9,0.99), eps=1e-8), clip 0.5).
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/ctclip_b200.h"

namespace ctb {

// normalised rows + inverse norms. grid = 2*B warps.  which = 0 text, 1 image
__global__ void clip_normalize_kernel(const float* __restrict__ t_raw, const float* __restrict__ i_raw, int B, int L,
                                      float* __restrict__ t_hat, float* __restrict__ i_hat, float* __restrict__ inv_norm) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= 2 * B) return;
  const float* src = (warp < B) ? t_raw + (long long)warp * L : i_raw + (long long)(warp - B) * L;
  float* dst = (warp < B) ? t_hat + (long long)warp * L : i_hat + (long long)(warp - B) * L;
  float ss = 0.f;
  for (int c = lane; c < L; c += 32) ss += src[c] * src[c];
  const float inv = 1.f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);  // F.normalize eps
  for (int c = lane; c < L; c += 32) dst[c] = src[c] * inv;
  if (lane == 0) inv_norm[warp] = inv;
}

// single CTA: logits, loss, dlogits (w.r.t. cosine sims), dtemperature
__global__ void __launch_bounds__(1024) clip_loss_kernel(const float* __restrict__ t_hat, const float* __restrict__ i_hat,
                                                        int B, int L, const float* __restrict__ temperature,
                                                        float* __restrict__ sim /*[B,B] scratch: logits then dC*/,
                                                        float* __restrict__ loss_out, float* __restrict__ dtemp_out,
                                                        float loss_scale) {
  extern __shared__ float sh[];  // rowsum[B], colsum[B], red[32]
  float* rowsum = sh;
  float* colsum = sh + B;
  float* red = sh + 2 * B;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const float temp = expf(temperature[0]);
  for (int p = warp; p < B * B; p += nwarps) {
    const int i = p / B, j = p % B;
    float s = 0.f;
    for (int c = lane; c < L; c += 32) s += t_hat[(long long)i * L + c] * i_hat[(long long)j * L + c];
    s = warp_sum(s);
    if (lane == 0) sim[p] = s * temp;  // text_to_image[i][j]
  }
  __syncthreads();
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    float rs = 0.f, cs = 0.f;
    for (int j = 0; j < B; j++) {
      rs += expf(sim[i * B + j]);
      cs += expf(sim[j * B + i]);
    }
    rowsum[i] = rs;  // text->image denominator for text i
    colsum[i] = cs;  // image->text denominator for image i
  }
  __syncthreads();
  float part = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const float pos = expf(sim[i * B + i]);
    part += (-logf(pos + 1e-20f) + logf(rowsum[i] + 1e-20f)) + (-logf(pos + 1e-20f) + logf(colsum[i] + 1e-20f));
  }
  part = warp_sum(part);
  if (lane == 0) red[warp] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < nwarps; w++) t += red[w];
    loss_out[0] = t / (2.f * B);
  }
  __syncthreads();
  // dL/dS_ij = E_ij/(2B) * (1/(rowsum_i+eps) + 1/(colsum_j+eps) - [i==j] * 2/(E_ii+eps));  dT = sum dS*S;  dC = dS*temp
  float dt = 0.f;
  for (int p = threadIdx.x; p < B * B; p += blockDim.x) {
    const int i = p / B, j = p % B;
    const float S = sim[p];
    const float E = expf(S);
    float d = 1.f / (rowsum[i] + 1e-20f) + 1.f / (colsum[j] + 1e-20f);
    if (i == j) d -= 2.f / (E + 1e-20f);
    const float dS = loss_scale * E * d / (2.f * B);
    dt += dS * S;
    sim[p] = dS * temp;
  }
  dt = warp_sum(dt);
  __syncthreads();
  if (lane == 0) red[warp] = dt;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < nwarps; w++) t += red[w];
    dtemp_out[0] = t;
  }
}

// d raw latents for rows [row0, row0 + nrows): dn = dC @ other_hat; draw = (dn - n (n.dn)) * inv_norm
__global__ void clip_grad_kernel(const float* __restrict__ dC, const float* __restrict__ t_hat,
                                 const float* __restrict__ i_hat, const float* __restrict__ inv_norm, int B, int L,
                                 int row0, int nrows, float* __restrict__ d_t_raw, float* __restrict__ d_i_raw) {
  extern __shared__ float dn[];  // [L]
  __shared__ float red[32];
  const int which = blockIdx.x / nrows;  // 0 text, 1 image
  const int r = row0 + blockIdx.x % nrows;
  const float* other = which == 0 ? i_hat : t_hat;
  const float* self = (which == 0 ? t_hat : i_hat) + (long long)r * L;
  float dot = 0.f;
  for (int c = threadIdx.x; c < L; c += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < B; j++) {
      const float w = which == 0 ? dC[r * B + j] : dC[j * B + r];
      s += w * other[(long long)j * L + c];
    }
    dn[c] = s;
    dot += s * self[c];
  }
  dot = warp_sum(dot);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); w++) tot += red[w];
  const float inv = inv_norm[which * B + r];
  float* dst = (which == 0 ? d_t_raw : d_i_raw) + (long long)(r - row0) * L;
  for (int c = threadIdx.x; c < L; c += blockDim.x) dst[c] = (dn[c] - self[c] * tot) * inv;
}

// inference similarity (ct_clip.py:805-807): out[k] = sum_d t_hat[k % Bt][d] * i_hat[k % Bi][d] * exp(T) with broadcasting
__global__ void clip_sims_kernel(const float* __restrict__ t_hat, int Bt, const float* __restrict__ i_hat, int Bi, int L,
                                 const float* __restrict__ temperature, float* __restrict__ out, int n_out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_out) return;
  const float* tr = t_hat + (long long)(Bt == 1 ? 0 : warp) * L;
  const float* ir = i_hat + (long long)(Bi == 1 ? 0 : warp) * L;
  float s = 0.f;
  for (int c = lane; c < L; c += 32) s += tr[c] * ir[c];
  s = warp_sum(s);
  if (lane == 0) out[warp] = s * expf(temperature[0]);
}

// ------------------------------------------------------------------------------------------------
// optimiser
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
  float s = 0.f;
  const long long n4 = n / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0)
    for (long long i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) s += g[i] * g[i];
  __shared__ float red[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; w++) t += red[w];
    atomicAdd(out, t);
  }
}

// torch.nn.utils.clip_grad_norm_ (coef = min(1, max_norm/(norm+1e-6))) followed by torch.optim.Adam (no amsgrad), or -- with
// decay_mul < 1 -- torch.optim.AdamW: the first n_decay elements of the arena (the tensors with ndim >= 2, optimizer.py:3-8,
// 26-34) are multiplied by decay_mul = 1 - lr * weight_decay before the Adam update (decoupled weight decay).
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                  float* __restrict__ v, long long n, float lr, float beta1, float beta2,
                                                  float eps, float bc1, float bc2_sqrt, float max_norm,
                                                  const float* __restrict__ sumsq, float grad_scale, float decay_mul,
                                                  long long n_decay) {
  float coef = grad_scale;
  if (max_norm > 0.f && sumsq != nullptr) {
    const float norm = sqrtf(sumsq[0]) * grad_scale;
    coef *= fminf(1.f, max_norm / (norm + 1e-6f));
  }
  const float step = lr / bc1;
  auto upd = [&](float gi_, float& mi_, float& vi_, float& pi_) {
    const float gi = gi_ * coef;
    mi_ = beta1 * mi_ + (1.f - beta1) * gi;
    vi_ = beta2 * vi_ + (1.f - beta2) * gi * gi;
    pi_ -= step * mi_ / (sqrtf(vi_) / bc2_sqrt + eps);
  };
  // 16-byte accesses (the arena keeps every tensor 16-byte aligned): 28 B/param of traffic, four streams per thread
  const long long n4 = (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                          reinterpret_cast<uintptr_t>(v)) & 15) == 0) ? n / 4 : 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 g4 = reinterpret_cast<const float4*>(g)[i];
    float4 m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i], p4 = reinterpret_cast<float4*>(p)[i];
    if (4 * i < n_decay) { p4.x *= decay_mul; p4.y *= decay_mul; p4.z *= decay_mul; p4.w *= decay_mul; }   // n_decay % 4 == 0
    upd(g4.x, m4.x, v4.x, p4.x);
    upd(g4.y, m4.y, v4.y, p4.y);
    upd(g4.z, m4.z, v4.z, p4.z);
    upd(g4.w, m4.w, v4.w, p4.w);
    reinterpret_cast<float4*>(m)[i] = m4;
    reinterpret_cast<float4*>(v)[i] = v4;
    reinterpret_cast<float4*>(p)[i] = p4;
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float mi = m[i], vi = v[i], pi = p[i];
    if (i < n_decay) pi *= decay_mul;
    upd(g[i], mi, vi, pi);
    m[i] = mi;
    v[i] = vi;
    p[i] = pi;
  }
}

}  // namespace ctb

using namespace ctb;

// ------------------------------------------------------------------------------------------------
// Data-parallel latent exchange over NVLink peer memory (replaces cat + NCCL all-gather + two copies in front of the loss:
// the north_star's "all-gather of image/text embeddings before the similarity matmul"). Every rank owns a gather buffer
// [2 parities][text | image][world*b][L] fp32 and a flag block [2][world] in symmetric memory that all peers have mapped.
// CTA r of rank k PUSHES rank k's b rows into peer r's buffer (16-byte P2P stores), fences system-wide, releases
// flag[parity][k] = seq in peer r's flag block, then waits until peer r's rows have landed in its OWN buffer
// (flag[parity][r] == seq, ld.acquire.sys). When the kernel retires, the local buffer holds the global batch.
// Two parities: a rank can run at most one step ahead of a peer (it cannot pass the wait of step s+1 before the peer has
// published s+1, which the peer does after its own step-s kernels), so buffers of step s are never overwritten while read.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) latent_exchange_kernel(const float* __restrict__ t_raw, const float* __restrict__ i_raw, int b,
                                                             int L, int rank, int world, const unsigned long long* __restrict__ peer_bufs,
                                                             const unsigned long long* __restrict__ peer_flags, unsigned seq) {
  const int r = blockIdx.x;   // peer served by this CTA
  const int par = (int)(seq & 1u);
  const size_t half = (size_t)world * b * L;
  float* base = reinterpret_cast<float*>(peer_bufs[r]) + (size_t)par * 2 * half + (size_t)rank * b * L;
  float4* dt = reinterpret_cast<float4*>(base);
  float4* di = reinterpret_cast<float4*>(base + half);
  const float4* st = reinterpret_cast<const float4*>(t_raw);
  const float4* si = reinterpret_cast<const float4*>(i_raw);
  const int n4 = b * L / 4;
  for (int k = threadIdx.x; k < n4; k += blockDim.x) {
    dt[k] = st[k];
    di[k] = si[k];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned* pf = reinterpret_cast<unsigned*>(peer_flags[r]) + par * world + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pf), "r"(seq) : "memory");
    const unsigned* lf = reinterpret_cast<const unsigned*>(peer_flags[rank]) + par * world + r;
    unsigned v = 0;
    unsigned long long spins = 0;
    while (true) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(lf) : "memory");
      if (v == seq) break;
      if (++spins > (1ull << 28)) {   // a peer that never arrives must not hang the box
        printf("ctclip: latent exchange watchdog: rank %d waited for rank %d, step %u (flag %u)\n", rank, r, seq, v);
        __trap();
      }
    }
  }
  __syncthreads();
}

extern "C" int ctclip_clip_loss(const ctclip_loss_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(a && a->t_raw && a->i_raw && a->temperature && a->t_hat && a->i_hat && a->inv_norm && a->sim,
                "clip_loss: null pointer");
  CTB_CHECK_ARG(a->B > 0 && a->B <= 256 && a->L > 0 && a->L <= 4096, "clip_loss: B in [1,256], L in [1,4096]");
  clip_normalize_kernel<<<ceil_div((long long)2 * a->B * 32, 256), 256, 0, stream>>>(a->t_raw, a->i_raw, a->B, a->L,
                                                                                    a->t_hat, a->i_hat, a->inv_norm);
  CTB_LAUNCH_CHECK();
  if (a->loss == nullptr) return CTCLIP_OK;  // normalise only (inference path)
  CTB_CHECK_ARG(a->dtemperature != nullptr, "clip_loss: dtemperature required with loss");
  const size_t smem = sizeof(float) * (2 * a->B + 32);
  clip_loss_kernel<<<1, 1024, smem, stream>>>(a->t_hat, a->i_hat, a->B, a->L, a->temperature, a->sim, a->loss,
                                             a->dtemperature, a->loss_scale);
  CTB_LAUNCH_CHECK();
  if (a->d_t_raw != nullptr && a->nrows > 0) {
    CTB_CHECK_ARG(a->d_i_raw != nullptr && a->row0 >= 0 && a->row0 + a->nrows <= a->B, "clip_loss: bad grad row range");
    clip_grad_kernel<<<2 * a->nrows, 256, sizeof(float) * a->L, stream>>>(a->sim, a->t_hat, a->i_hat, a->inv_norm, a->B,
                                                                         a->L, a->row0, a->nrows, a->d_t_raw, a->d_i_raw);
    CTB_LAUNCH_CHECK();
  }
  return CTCLIP_OK;
}

extern "C" int ctclip_clip_sims(const float* t_hat, int32_t Bt, const float* i_hat, int32_t Bi, int32_t L,
                                const float* temperature, float* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(t_hat && i_hat && temperature && out && Bt > 0 && Bi > 0 && L > 0, "clip_sims: bad args");
  CTB_CHECK_ARG(Bt == Bi || Bt == 1 || Bi == 1, "clip_sims: batch sizes %d and %d do not broadcast", Bt, Bi);
  const int n_out = Bt > Bi ? Bt : Bi;
  clip_sims_kernel<<<ceil_div((long long)n_out * 32, 256), 256, 0, stream>>>(t_hat, Bt, i_hat, Bi, L, temperature, out, n_out);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}

extern "C" int ctclip_grad_sumsq(const float* g, int64_t n, float* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(g && out && n > 0, "grad_sumsq: bad args");
  CTB_CHECK_ARG(((uintptr_t)g % 16) == 0, "grad_sumsq: g must be 16B aligned");
  long long ctas = (n / 4 + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (ctas > cap) ctas = cap;
  if (ctas < 1) ctas = 1;
  sumsq_kernel<<<(int)ctas, 256, 0, stream>>>(g, n, out);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}

extern "C" int ctclip_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                                float eps, int32_t step, float max_norm, const float* sumsq, float grad_scale,
                                float weight_decay, int64_t n_decay, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(p && g && m && v && n > 0 && step >= 1, "adam_step: bad args");
  CTB_CHECK_ARG(weight_decay >= 0.f && n_decay >= 0 && n_decay <= n && n_decay % 4 == 0, "adam_step: bad weight-decay range");
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2 = 1.f - powf(beta2, (float)step);
  long long ctas = (n + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (ctas > cap) ctas = cap;
  adam_kernel<<<(int)ctas, 256, 0, stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, bc1, sqrtf(bc2), max_norm, sumsq,
                                            grad_scale, 1.f - lr * weight_decay, weight_decay > 0.f ? (long long)n_decay : 0LL);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}

extern "C" int ctclip_latent_exchange(const float* t_raw, const float* i_raw, int32_t b, int32_t L, int32_t rank, int32_t world,
                                      const uint64_t* peer_bufs, const uint64_t* peer_flags, uint32_t seq, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTB_CHECK_ARG(t_raw && i_raw && peer_bufs && peer_flags, "latent_exchange: null pointer");
  CTB_CHECK_ARG(b > 0 && L > 0 && (b * (long long)L) % 4 == 0, "latent_exchange: b*L must be a multiple of 4");
  CTB_CHECK_ARG(world >= 1 && world <= 64 && rank >= 0 && rank < world && seq > 0, "latent_exchange: bad rank / world / step");
  latent_exchange_kernel<<<world, 256, 0, stream>>>(t_raw, i_raw, b, L, rank, world,
                                                   reinterpret_cast<const unsigned long long*>(peer_bufs),
                                                   reinterpret_cast<const unsigned long long*>(peer_flags), seq);
  CTB_LAUNCH_CHECK();
  return CTCLIP_OK;
}
