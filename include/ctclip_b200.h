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
 *   5 ARGMAX     arg_out[m] = argmax_n acc (first max wins), argval_out[m] = max (optional)
 * splits: split-K factor (only with ATOMIC_F32), >= 1.
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
} ctclip_gemm_args;

int ctclip_gemm_bf16(const ctclip_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CTCLIP_B200_H */
