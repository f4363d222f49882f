/* ctclip_b200.h -- C ABI of libctclip_b200.so (hand-written sm_100a kernels for the CT-CLIP
 * contrastive forward/backward hot path).
 *
 * The reference (ibrahimethemhamamci/CT-CLIP) has no FFI layer: its hot path is PyTorch eager
 * code. Each entry point below replaces the group of eager ops named in its comment
 * (reference file:line). The Python classes in ct_clip_b200/ (CTViT, CTCLIP, CTClipTrainer,
 * CTClipInference) bind these symbols with ctypes -- see INTEGRATION.md.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise;
 *   - the caller owns all memory (inputs, outputs, workspaces, saved-for-backward buffers);
 *     the library never allocates, frees or synchronises;
 *   - work is enqueued on the cudaStream_t passed as the last argument (as void*);
 *   - return value 0 = success; non-zero = error, message via ctclip_last_error()
 *     (thread-local);
 *   - "bf16" = __nv_bfloat16 storage, "f32" = float; row-major; ld* = leading dimension in
 *     ELEMENTS.
 */
#ifndef CTCLIP_B200_H
#define CTCLIP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTCLIP_B200_VERSION 100

int ctclip_version(void);
const char* ctclip_last_error(void);

/* ------------------------------------------------------------------------------------------
 * GEMM family (tcgen05.mma + TMA + TMEM), bf16 operands, fp32 accumulation.
 *   C[M,N] = sum_k A(m,k) * B(n,k)
 * replaces every nn.Linear / einsum on the path: attention.py:145 (to_q/to_kv), :181 (to_out),
 * :48,:51 (FeedForward linears), ctvit.py:173 (patch Linear), the VQ distance einsum
 * (vector_quantize_pytorch CosineSimCodebook.forward), and their autograd backward GEMMs.
 *
 * a_major / b_major: 0 = operand stored [rows, K] with K contiguous ("K-major");
 *                    1 = operand stored [K, rows] with rows contiguous ("MN-major",
 *                        used by the weight-gradient GEMMs, which reduce over tokens).
 * epilogue:
 *   0 BF16       C(bf16)[m,n]  = acc (+bias[n])
 *   1 F32        C(f32)[m,n]   = acc (+bias[n])
 *   2 RESID_F32  C(f32)[m,n]   = acc (+bias[n]) + resid(f32)[m,n]      (C may alias resid)
 *   3 GEGLU      columns are interleaved (value_j at 2j, gate_j at 2j+1):
 *                C(bf16)[m,n]  = acc (+bias[n])           (pre-activation, saved for backward;
 *                                                           skipped if C == NULL)
 *                C2(bf16)[m,j] = gelu_erf(gate_j) * value_j
 *   4 ATOMIC_F32 C(f32)[m,n]  += acc      (red.global.add; with splits > 1 = split-K)
 *   5 ARGMAX     arg_out[m] = argmax_n acc (first max wins), argval_out[m] = max (optional), arg2_out[m] = index of the
 *                runner-up (optional; ctclip_vq_rerank re-ranks the pair in fp32)
 *   6 L2NORM     per 32-column group (= one attention head, dim_head 32):
 *                C(bf16)[m,n]  = acc                      (raw q/k/v, optional)
 *                C2(bf16)[m,n] = acc / max(||acc_group||, 1e-12) * norm_scale[n % 32]   for n < norm_cols
 *                (attention.py:152-154 l2norm(q)*q_scale fused into the projection)
 *   7 BIAS_GELU  C(bf16) = gelu_erf(acc + bias), C2(bf16, optional) = acc + bias  (BERT intermediate)
 *   8 GEGLU_BWD  acc = dL/d(gelu(gate_j) value_j); C = h(bf16)[m, 2N] interleaved pre-activation, updated IN PLACE:
 *                h[m,2j] <- acc * gelu(gate_j), h[m,2j+1] <- acc * value_j * gelu'(gate_j); colsum[2N] += column sums
 *                of the result (optional). N <= 1536. Fuses ctclip_geglu_bwd into the preceding GEMM.
 * splits: split-K factor (only with ATOMIC_F32), >= 1; 0 = chosen by the library (minimises waves x k-blocks per unit).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t M, N, K;
  int32_t a_major, b_major;
  const void* A;
  int64_t lda;
  const void* B;
  int64_t ldb;
  int32_t epilogue;
  int32_t splits;
  void* C;
  int64_t ldc;
  const float* bias;
  const float* resid;
  int64_t ldr;
  void* C2;
  int64_t ldc2;
  int32_t* arg_out;
  float* argval_out;
  int32_t norm_cols;
  const float* norm_scale;
  float* colsum;          /* GEGLU_BWD: [2N] column sums, accumulated */
  int32_t* arg2_out;      /* ARGMAX: optional runner-up index */
} ctclip_gemm_args;

int ctclip_gemm_bf16(const ctclip_gemm_args* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * LayerNorm over rows of an fp32 [M, D] tensor (D multiple of 128, <= 1024).
 * Replaces F.layer_norm at attention.py:35, attention.py:47, ctvit.py:174, attention.py:333 and
 * BERT's LayerNorms. Outputs are optional (NULL = skip):
 *   xhat_bf16 = (x-mean)*rstd (standardised row: the GEMM operand when gamma/beta are folded into
 *   the next Linear), raw_bf16 = bf16(x) (K/V projections read the raw stream, attention.py:139-145),
 *   y_f32 / y_bf16 = xhat*gamma + beta, rstd_out[M].
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const float* x;
  int64_t M;
  int32_t D;
  float eps;
  const float* gamma;
  const float* beta;
  uint16_t* xhat_bf16;
  uint16_t* raw_bf16;
  float* y_f32;
  uint16_t* y_bf16;
  float* rstd_out;
} ctclip_ln_fwd_args;
int ctclip_ln_fwd(const ctclip_ln_fwd_args* args, void* stream);

/* Backward of the above. g_* = gradient w.r.t. the LN output (gamma != NULL; dgamma/dbeta are
 * accumulated with atomics) or w.r.t. xhat (gamma == NULL).
 *   dx = rstd*(dxhat - mean(dxhat) - xhat*mean(dxhat*xhat)) + dres_in + add_bf16   (optional terms)
 * written as fp32 (dx_f32) and/or bf16 (dx_bf16). dx_f32 may alias dres_in. */
typedef struct {
  int64_t M;
  int32_t D;
  const float* g_f32;
  const uint16_t* g_bf16;
  const float* gamma;
  const uint16_t* xhat;
  const float* rstd;
  const float* dres_in;
  const uint16_t* add_bf16;
  float* dx_f32;
  uint16_t* dx_bf16;
  float* dgamma;
  float* dbeta;
} ctclip_ln_bwd_args;
int ctclip_ln_bwd(const ctclip_ln_bwd_args* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * Tubelet im2col + per-patch standardisation: ctvit.py:171-172 (Rearrange + LayerNorm(P) without its
 * affine, which is folded into the patch Linear). video: [B,C,F,H,W] fp32 (dtype 0) or int16 HU
 * (dtype 1, value = int16 * scale; scripts/data.py:122-125 uses 1/1000). xhat: bf16 [B*T*Ht*Wt, ld_out],
 * feature order (c, pt, p1, p2).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const void* video;
  int32_t dtype;
  float scale;
  int32_t B, C, F, H, W;
  int32_t pt, p1, p2;
  float eps;
  uint16_t* xhat;
  int64_t ld_out;
} ctclip_patchify_args;
int ctclip_patchify(const ctclip_patchify_args* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * PEG: causal depthwise 3x3x3 conv + residual on the canonical fp32 token stream [B,T,H,W,D].
 * attention.py:63-84 + the residual at :324. temporal = 1 reproduces the reference's reshape of
 * the (b,h,w,t)-ordered tokens as (b,T,H,W) (SURVEY trap T1).
 * Exact fp32 arithmetic (packed FFMA2): the plane-streaming TMA kernels of csrc/peg_stream.cu whenever the token grid allows
 * them (W <= 24, D % 32 == 0, temporal only for T == H == W), the general stencil of csrc/peg.cu otherwise.
 *   ctclip_peg_fwd        : y = x + conv(x) + bias          (y_bf16 optional bf16 copy)
 *   ctclip_peg_bwd_data   : y = x + conv^T(x)  with x = upstream gradient
 *   ctclip_peg_bwd_weight : dweight[D,27] += ..., dbias[D] += ...  (x = forward input, dy = upstream)
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const float* x;
  const float* dy;
  float* y;
  uint16_t* y_bf16;
  const float* weight; /* [D, 27] = dsconv.weight (D,1,3,3,3) */
  const float* bias;   /* [D] */
  float* dweight;
  float* dbias;
  int32_t B, T, H, W, D;
  int32_t temporal;
  int32_t lines;       /* reserved (was: selector of a bf16 tensor-core formulation that measured slower and was removed) */
  const int32_t* canon_table; /* optional, temporal only: canon_table[f] = canonical token of conv-grid index
                                 f = (a0*H + a1)*W + a2, i.e. ((f % T)*H + f / (T*W))*W + (f / T) % W */
} ctclip_peg_args;
int ctclip_peg_fwd(const ctclip_peg_args* args, void* stream);
int ctclip_peg_bwd_data(const ctclip_peg_args* args, void* stream);
int ctclip_peg_bwd_weight(const ctclip_peg_args* args, void* stream);
/* measurement / test knob (tools/peg_probe.py, tests): 0 = plane-streaming kernels (csrc/peg_stream.cu) whenever the token grid
 * allows them (default), 1 = always the general kernels of csrc/peg.cu */
int ctclip_debug_set_peg_variant(int32_t variant);

/* ------------------------------------------------------------------------------------------
 * Attention core (attention.py:156-178; also BERT self-attention), dim_head 32 or 64. q/k are the l2-normalised, scaled projections
 * (GEMM epilogue 6), v raw; all bf16 [rows, heads*32] with leading dimensions ldq/ldk/ldv.
 * Token rows: row(seq,i) = (seq / seq_inner)*seq_outer_stride + seq % seq_inner + i*tok_stride.
 * bias (optional): bf16 [heads, n, n]; bias_t its transpose over the last two dims (backward only).
 * lse: fp32 [rows, heads] log2-domain log-sum-exp (forward output, backward input).
 * Backward: o, d_o (ld = ldo), delta (fp32 [rows, heads] scratch) -> dq, dk (w.r.t. the normalised
 * q/k), dv; dbias (fp32 [heads,n,n], accumulated; NULL to skip). total_rows = rows of the token matrix.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const uint16_t* q; int64_t ldq;
  const uint16_t* k; int64_t ldk;
  const uint16_t* v; int64_t ldv;
  uint16_t* o; int64_t ldo;
  float* lse;
  const uint16_t* bias;
  const uint16_t* bias_t;
  int32_t n, heads, dim_head, num_seqs, seq_inner;
  int64_t seq_outer_stride, tok_stride;
  float scale;
  /* backward only */
  const uint16_t* d_o;
  float* delta;
  uint16_t* dq; int64_t ld_dq;
  uint16_t* dk; int64_t ld_dk;
  uint16_t* dv; int64_t ld_dv;
  float* dbias;
  int64_t total_rows;
  const int32_t* key_mask; /* optional [num_seqs, n]: non-zero = key may be attended (BERT padding mask) */
  /* optional: the same bias (and its transpose) re-ordered per MMA fragment by ctclip_cpb_expand_frag:
   * bf16 [heads, ceil16(n)/16, ceil64(n)/64, 32 lanes, 8 n-tiles, 4], values PRE-MULTIPLIED by log2(e) (the kernels
   * work in the log2 domain); when given they replace bias / bias_t in the forward, dQ and dK/dV kernels (fully
   * coalesced 16-byte loads). */
  const uint16_t* bias_frag;
  const uint16_t* bias_t_frag;
  /* optional scratch, bf16 [num_seqs*heads, n, n] (n even): with dbias, the dQ kernel spills its d logits there and a
   * streaming reduction over the sequences replaces the third (recomputing) backward pass */
  uint16_t* ds_scratch;
  /* tcgen05 / TMEM path (csrc/attention_tc.cu), selected by cpb_table != NULL; the spatial stack of CTViT, where the
   * sequence is the grid_h x grid_w token grid of one frame (n == grid_h*grid_w, contiguous sequences, dim_head 32, no key
   * mask) and the bias is the continuous position bias of attention.py:245-282:
   *   cpb_table  fp32 [(2*grid_h-1)*(2*grid_w-1), heads]: bias[h,i,j] = cpb_table[rel(i,j), h] with
   *              rel(i,j) = (yi-yj+grid_h-1)*(2*grid_w-1) + (xi-xj+grid_w-1)  (what ctclip_cpb_expand expands); bias,
   *              bias_t, bias_frag, bias_t_frag are ignored on this path;
   *   qk_bound   device scalar >= max_d |q_scale_d * k_scale_d| (NULL: 1): with unit-norm q_hat/q_scale, k_hat/k_scale it bounds
   *              |q_hat . k_hat|, which lets the kernels use a fixed softmax reference instead of an online maximum;
   *   dcpb_table backward: fp32 [(2*grid_h-1)*(2*grid_w-1), heads], ACCUMULATED gradient w.r.t. cpb_table (replaces dbias;
   *              needs ds_scratch, bf16 [num_seqs*heads, n, n]: d logits are spilled and reduced over the sequences);
   *   OR dbias   (with cpb_table set) fp32 [heads, n, n] holding the TRANSPOSED gradient d bias[h, i, j] at [h][j][i], accumulated
   *              with red.global.add.v4.f32 (stays in L2); ctclip_cpb_reduce_t turns it into the table gradient.
   * ctclip_attn_tc_supported() tells whether a geometry can take this path. */
  const float* cpb_table;
  int32_t grid_h, grid_w;
  const float* qk_bound;
  float* dcpb_table;
  /* attention-probability dropout (HF BertSelfAttention: `attention_probs = self.dropout(attention_probs)`), dim_head 64 path
   * only: element ((seq*heads + head)*n + i)*n + j of the probability tensor is kept per the rule of ctclip_dropout. */
  float dropout_p;
  uint64_t dropout_seed, dropout_offset;
} ctclip_attn_args;
int ctclip_attn_fwd(const ctclip_attn_args* args, void* stream);
int ctclip_attn_bwd(const ctclip_attn_args* args, void* stream);
/* bit 0: the tcgen05 forward kernel takes this geometry; bit 1: the tcgen05 backward kernel does (NOT an error code) */
int ctclip_attn_tc_supported(int32_t n, int32_t grid_h, int32_t grid_w, int32_t dim_head);
/* out[0] = max_d |q_scale[d] * k_scale[d]|  (attention.py:131-132 parameters): the qk_bound of ctclip_attn_args */
int ctclip_qk_bound(const float* q_scale, const float* k_scale, int32_t dim_head, float* out, void* stream);
/* measurement knob (tools/attn_tc_probe.py, tests; not used by the product path): variant of the tcgen05 backward kernel.
 * 8 / 16 = element-wise warps with per-query records in shared memory; 108 / 116 = the same with lse and delta folded into the
 * MMAs as two extra k-steps (108 is the default) */
int ctclip_debug_set_attn_bwd_warps(int32_t warps);

/* Backward of x_hat = x/max(||x||,1e-12)*scale per (row, head) (attention.py:152-154):
 * dxh, xraw, dx: bf16 [rows, heads*32]; dscale[32] accumulated. */
int ctclip_l2norm_bwd(const void* dxh, int64_t ld_dxh, const void* xraw, int64_t ld_x, const float* scale, void* dx,
                      int64_t ld_dx, float* dscale, int64_t rows, int32_t heads, int32_t dim_head, void* stream);

/* ------------------------------------------------------------------------------------------
 * Small fp32 GEMM on CUDA cores for the continuous-position-bias MLP (attention.py:257-276 forces
 * fp32): C[M,N] = epi(sum_k opA(m,k) opB(k,n)); opA = A^T if trans_a (A stored [K,M]); opB = B^T if
 * trans_b (B stored [N,K], i.e. an nn.Linear weight). epi: +bias[n], act 1 = LeakyReLU(0.1),
 * mask_ref: multiply by LeakyReLU'(ref) where ref is the forward activation output; accumulate: C += .
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t M, N, K;
  const float* A; int64_t lda; int32_t trans_a;
  const float* B; int64_t ldb; int32_t trans_b;
  float* C; int64_t ldc;
  const float* bias;
  int32_t act;
  const float* mask_ref; int64_t ld_mask;
  int32_t accumulate;
} ctclip_sgemm_args;
int ctclip_sgemm_f32(const ctclip_sgemm_args* args, void* stream);

/* out[n] += sum_m x[m,n]   (x fp32 or bf16 [M, ld]) -- bias gradients */
int ctclip_colsum(const void* x, int32_t is_bf16, int64_t ld, int64_t M, int32_t N, float* out, void* stream);
int ctclip_cast_f32_bf16(const float* x, void* y, int64_t n, void* stream);

/* Continuous position bias on the (2h-1)(2w-1) distinct relative offsets (attention.py:263-267):
 * inputs X[R,2] = sign(rel)*log(|rel|+1); expand table[R,heads] -> bias bf16 [heads,n,n] (+ transposed
 * copy); reduce dbias fp32 [heads,n,n] -> dtable[R,heads]. */
int ctclip_cpb_inputs(float* X, int32_t h, int32_t w, void* stream);
int ctclip_cpb_expand(const float* table, int32_t heads, int32_t h, int32_t w, void* bias, void* bias_t, void* stream);
int ctclip_cpb_reduce(const float* dbias, int32_t heads, int32_t h, int32_t w, float* dtable, void* stream);
/* the same reduction for the TRANSPOSED table dbias_t fp32 [heads, n(j), n(i)] = d bias[h, i, j] that the tcgen05 backward accumulates
 * when ctclip_attn_args.dbias is given together with cpb_table; the result is ADDED to dtable. */
int ctclip_cpb_reduce_t(const float* dbias_t, int32_t heads, int32_t h, int32_t w, float* dtable, void* stream);
/* fragment-ordered copies of bias * log2(e) and of its transpose (see ctclip_attn_args.bias_frag) */
int ctclip_cpb_expand_frag(const float* table, int32_t heads, int32_t h, int32_t w, void* bias_frag, void* bias_t_frag,
                           void* stream);

/* GEGLU backward (attention.py:39-42) on the interleaved pre-activation h bf16 [M, 2*n_pairs] (in place):
 * h[:,2j] <- dg[:,j]*gelu(gate_j), h[:,2j+1] <- dg[:,j]*value_j*gelu'(gate_j); colsum[2*n_pairs] += column sums. */
int ctclip_geglu_bwd(const void* dg, int64_t ld_dg, void* h, int64_t ld_h, int64_t M, int32_t n_pairs, float* colsum,
                     void* stream);

/* ------------------------------------------------------------------------------------------
 * GPU input pipeline (scripts/data.py:92-162 nii_img_to_tensor, scripts/data_inference_nii.py:96-176): raw NIfTI voxels
 * (x, y, z)-contiguous as nibabel returns them -> slope/intercept -> trilinear resample to (target_xy, target_xy, target_z)
 * spacing (F.interpolate, align_corners=False, size = int(dim * current / target)) -> clip [-1000, 1000] -> /1000 ->
 * centre crop / pad (pad_value, reference: -1) -> out [out_d, out_h, out_w] = (z, x, y) order, the (1, 240, 480, 480) volume
 * of the dataset contract. out_dtype 0: fp32 in [-1, 1]; 1: int16 HU (rounded), read as x/1000 by ctclip_patchify.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const void* raw;
  int32_t raw_dtype;        /* 0 = float32, 1 = int16 */
  int32_t X, Y, Z;
  float slope, intercept;   /* RescaleSlope / RescaleIntercept of the metadata table */
  double xy_spacing, z_spacing;   /* doubles: the resized shape int(dim * (current / target)) must round like the reference's */
  double target_xy, target_z;     /* Python floats (0.6 / 0.75 * 40 = 31.999999999999996 -> 31); reference targets: 0.75, 1.5 */
  int32_t out_d, out_h, out_w; /* reference: 240, 480, 480 */
  void* out;
  int32_t out_dtype;
  float pad_value;          /* reference: -1 */
} ctclip_preprocess_args;
int ctclip_ct_preprocess(const ctclip_preprocess_args* args, void* stream);

/* Retrieval over saved latents (scripts/report_to_volume_new.py:47-63, scripts/volume_to_volume_new.py:80-96): the k best
 * columns of every row of scores fp32 [Q, ld >= G], descending, lower index first among equal scores (Python's stable
 * sorted(..., reverse=True)); G <= 49152. ctclip_l2norm_rows_f32: y = x / max(||x||, 1e-12) per row (cosine similarity). */
int ctclip_topk_rows(const float* scores, int64_t ld, int32_t Q, int32_t G, int32_t k, int32_t* idx_out, float* val_out, void* stream);
int ctclip_l2norm_rows_f32(const float* x, float* y, int32_t rows, int32_t D, void* stream);

/* Dropout with counter-based masks (csrc/rng.cuh: Philox4x32-10; element idx of a site is kept iff
 * philox(seed, offset + idx/4).word[idx%4] >= floor(p*2^32)), so backward regenerates the mask from the same (seed, offset):
 *   y = resid + keep * x / (1-p)   (resid optional; y_f32, y_bf16 optional outputs; n elements, 16-byte aligned pointers).
 * The backward of the same site is the same call on the upstream gradient with resid = NULL.
 * HF BertModel sites (modeling_bert.py): embeddings, BertSelfOutput, BertOutput; attention-probability dropout lives inside
 * ctclip_attn_fwd / ctclip_attn_bwd (dropout_p / dropout_seed / dropout_offset of ctclip_attn_args). */
int ctclip_dropout(const float* x, const float* resid, float* y_f32, void* y_bf16, int64_t n, float p, uint64_t seed,
                   uint64_t offset, void* stream);

/* Vector quantiser pieces (vector_quantize_pytorch==1.1.2 CosineSimCodebook, called at ctvit.py:403):
 * the argmax itself is GEMM epilogue 5 on (tokens, l2norm(embed)). */
int ctclip_l2norm_rows_bf16(const float* x, void* y, int32_t rows, int32_t D, void* stream);
/* fp32 re-ranking of the bf16 GEMM's top-2 candidates: idx[m] <- argmax over {idx[m], idx2[m]} of x[m,:] . embed[c,:] / ||embed[c,:]||
 * (the cosine code-book's ranking, vector_quantize_pytorch 1.1.2 CosineSimCodebook.forward; first index wins ties).
 * x fp32 [M, D] (the quantiser input), embed fp32 [C, D] (un-normalised master copy). */
int ctclip_vq_rerank(const float* x, const float* embed, int32_t* idx, const int32_t* idx2, int64_t M, int32_t D, void* stream);
int ctclip_vq_gather(const int32_t* idx, const float* embed, float* out, int64_t M, int32_t D, void* stream);
int ctclip_vq_gather_pool(const int32_t* idx, const float* embed, int32_t B, int32_t T, int32_t S, int32_t D,
                          float* pooled_f32, void* pooled_bf16, void* stream);   /* + ct_clip.py:724 mean over t */
int ctclip_pool_bwd(const float* dpooled, int32_t B, int32_t T, int32_t S, int32_t D, float* dtok, void* stream);
int ctclip_vq_ema_accum(const float* x, const int32_t* idx, int64_t M, int32_t D, float* bins, float* embed_sum,
                        void* stream);
int ctclip_vq_ema_update(float* embed, float* cluster_size, const float* bins, const float* embed_sum, int32_t C, int32_t D,
                         float decay, void* stream);

/* Weight preparation: out bf16 [Np,Kp] = W[rowmap[r], k] * gamma[k] (zero where rowmap[r] < 0 or k >= K);
 * bias' [Np] = W[rowmap[r], :] . beta + bias_in[rowmap[r]]; and the reverse mapping of gradients. */
int ctclip_prep_weight(const float* W, int64_t ldw, int32_t K, const float* gamma, const int32_t* rowmap, int32_t Np,
                       int32_t Kp, void* out, void* stream);
/* The same two operations for a whole model in ONE launch: `descs` is a DEVICE array of n descriptors (kind 0: operand as
 * ctclip_prep_weight, kind 1: bias as ctclip_prep_bias), blocks_per_desc CTAs of 256 threads work on each descriptor. */
typedef struct {
  const float* W;
  int64_t ldw;
  const float* gamma;       /* kind 0 */
  const float* beta;        /* kind 1 */
  const float* bias_in;     /* kind 1 */
  const int32_t* rowmap;
  void* out;                /* kind 0: bf16 [Np, Kp]; kind 1: f32 [Np] */
  int32_t K, Np, Kp, kind;
} ctclip_prep_desc;
int ctclip_prep_batched(const ctclip_prep_desc* descs, int32_t n, int32_t blocks_per_desc, void* stream);
int ctclip_prep_bias(const float* W, int64_t ldw, int32_t K, const float* beta, const float* bias_in, const int32_t* rowmap,
                     int32_t Np, float* out, void* stream);
int ctclip_unprep_wgrad(const float* G, int64_t ldg, const float* W, int64_t ldw, int32_t K, const float* gamma,
                        const int32_t* rowmap, int32_t Np, const float* s, float* dW, float* dgamma, float* dbeta,
                        float* dbias, void* stream);

/* ------------------------------------------------------------------------------------------
 * Contrastive head, forward + backward in one call (ct_clip.py:771, :796, :845-878).
 * t_raw/i_raw: fp32 [B, L] un-normalised latents of the GLOBAL batch. Scratch (caller-allocated):
 * t_hat/i_hat [B,L], inv_norm [2B], sim [B,B]. Outputs: loss[1], dtemperature[1], and for the rows
 * [row0, row0+nrows) owned by this rank d_t_raw/d_i_raw [nrows, L]. loss == NULL: only normalise
 * (inference). loss_scale multiplies all gradients.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const float* t_raw;
  const float* i_raw;
  int32_t B, L;
  const float* temperature;
  float* t_hat;
  float* i_hat;
  float* inv_norm;
  float* sim;
  float* loss;
  float* dtemperature;
  float* d_t_raw;
  float* d_i_raw;
  int32_t row0, nrows;
  float loss_scale;
} ctclip_loss_args;
int ctclip_clip_loss(const ctclip_loss_args* args, void* stream);
/* Data-parallel exchange of the raw latents over NVLink peer memory, in front of ctclip_clip_loss (replaces the embedding
 * all-gather the north_star names; the reference's own gather, CT_CLIP/ct_clip/distributed.py:9-34, is dead code).
 * peer_bufs[r] / peer_flags[r]: device addresses, valid on THIS device, of rank r's gather buffer
 * ([2 parities][text | image][world*b][L] fp32) and flag block ([2][world] uint32, zero before the first call); the caller
 * maps them (e.g. torch symmetric memory). step >= 1 and identical on all ranks, +1 per call. On return (stream order) the
 * local buffer of parity step & 1 holds the global batch ordered by rank. One kernel: push + release + acquire. */
int ctclip_latent_exchange(const float* t_raw, const float* i_raw, int32_t b, int32_t L, int32_t rank, int32_t world,
                           const uint64_t* peer_bufs, const uint64_t* peer_flags, uint32_t step, void* stream);
/* inference similarity (ct_clip.py:805-807), broadcasting a batch of 1 */
int ctclip_clip_sims(const float* t_hat, int32_t Bt, const float* i_hat, int32_t Bi, int32_t L, const float* temperature,
                     float* out, void* stream);

/* Optimiser over a flat fp32 arena: out[0] += sum g^2; then clip_grad_norm_(max_norm) + Adam
 * (CTCLIPTrainer.py:259-263, optimizer.py:23-24). grad_scale multiplies g before everything else.
 * weight_decay > 0 = torch.optim.AdamW (optimizer.py:26-34): the first n_decay elements (a multiple of 4; the caller lays
 * the ndim >= 2 tensors out first, optimizer.py:3-8) are scaled by 1 - lr*weight_decay before the Adam update. */
int ctclip_grad_sumsq(const float* g, int64_t n, float* out, void* stream);
int ctclip_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                     int32_t step, float max_norm, const float* sumsq, float grad_scale, float weight_decay, int64_t n_decay,
                     void* stream);

/* BERT text tower helpers (transformers.BertModel, called at ct_clip.py:685): embeddings gather
 * (word[ids] + position + token_type 0) and its scatter-add backward; GELU backward of BertIntermediate
 * (dy <- dy * gelu'(pre), colsum += bias gradient). The Linear layers use ctclip_gemm_bf16 (epilogues 0/2/7),
 * the LayerNorms ctclip_ln_*, self-attention ctclip_attn_* with dim_head 64 and key_mask. */
int ctclip_bert_embed(const int64_t* ids, const float* word, const float* pos, const float* type0, float* out, int64_t rows,
                      int32_t n, int32_t H, void* stream);
int ctclip_bert_embed_bwd(const int64_t* ids, const float* g, float* dword, float* dpos, int64_t rows, int32_t n, int32_t H,
                          void* stream);
int ctclip_gelu_bwd(void* dy, int64_t ld_dy, const void* pre, int64_t ld_pre, int64_t M, int32_t N, float* colsum,
                    void* stream);
/* zero-shot head (scripts/zero_shot.py:133-143): probs[v, p] = softmax over the prompt pair (2p, 2p+1) of
 * img_hat[v] . txt_hat[.] * exp(T); img [V,L], txt [P2,L] l2-normalised fp32, probs [V, P2/2]. */
int ctclip_zero_shot_probs(const float* img, const float* txt, int32_t V, int32_t P2, int32_t L, const float* temperature,
                           float* probs, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CTCLIP_B200_H */
