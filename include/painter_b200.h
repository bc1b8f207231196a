/* painter_b200 — C ABI of the B200-native Painter/SegGPT hot path.
 *
 * The reference (baaivision/Painter) has no FFI layer: its hot path is the Python module
 * Painter/models_painter.py (+ SegGPT/SegGPT_inference/models_seggpt.py, util/vitdet_utils.py), every
 * arithmetic step being a torch op.  Each entry point below replaces one group of those torch call
 * sites (cited per function) with a hand-written sm_100a kernel.  The host-side mirror of the
 * reference nn.Module API (painter_b200/models_painter.py, models_seggpt.py) calls these through
 * ctypes; see INTEGRATION.md for the binding.
 *
 * Conventions
 *  - all pointers are DEVICE pointers owned by the caller (PyTorch's caching allocator); the library
 *    never allocates device memory and keeps no pointer past return;
 *  - `stream` is a cudaStream_t passed as void*; kernels are enqueued asynchronously on it;
 *  - return value 0 = ok; otherwise pk_last_error() describes the failure (thread-local);
 *  - bf16 = raw 16-bit bfloat16, f32 = IEEE float. "tokens" are rows of [B*h*w, C] matrices.
 */
#ifndef PAINTER_B200_H
#define PAINTER_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

int pk_version(void);
const char* pk_last_error(void);
/* number of kernel launches issued by this library since load (bench.py's gpu_launches claim) */
long long pk_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Dense contraction on tcgen05 tensor cores:  C[M,N] = A[M,K] . B[N,K]^T   (bf16 x bf16 -> fp32)
 * Replaces every nn.Linear / F.linear on the path (models_painter.py:60-61,76,87 qkv/proj; timm Mlp
 * fc1/fc2 used at :201; decoder_embed :327,423) and their autograd backward (dgrad / wgrad GEMMs).
 *   transA = 0: A stored [M,K] row-major (lda = row stride, elements)   transA = 1: stored [K,M]
 *   transB = 0: B stored [N,K] row-major (ldb)                          transB = 1: stored [K,N]
 * Epilogues (PkEpilogue.kind):
 */
enum {
  PK_EPI_BF16 = 0,    /* out(bf16)[m,n] = alpha*acc + bias[n]                                     */
  PK_EPI_F32 = 1,     /* out(f32)[m,n]  = alpha*acc + bias[n] (+ out[m,n] when accumulate != 0)    */
  PK_EPI_GELU = 2,    /* out(bf16) = z = acc + bias ; out2(bf16) = gelu_erf(bf16(z))  (Mlp fc1+act) */
  PK_EPI_RESID = 3,   /* out(f32) = aux_f32[m,n] + rowscale[m / rows_per_group] * (acc + bias[n])
                         (residual add + DropPath scale, models_painter.py:229-230)               */
  PK_EPI_DGELU = 4,   /* out(bf16) = acc * gelu'(aux_bf16[m,n])        (backward through GELU)     */
  PK_EPI_PIXSHUF = 5, /* out(bf16) NHWC [B, h*p, w*p, c]: decoder_embed + 'nhwpqc->nchpwq' pixel
                         shuffle (models_painter.py:423-428), rows m=(b,i,j), cols n=(r,s,c)       */
};
typedef struct {
  int kind;
  void* out;
  void* out2;
  const float* bias;
  const void* aux;
  const float* rowscale;
  int ldc;            /* row stride of out/out2 (elements) */
  int ld_aux;         /* row stride of aux (elements) */
  int rows_per_group; /* rows sharing one rowscale entry */
  int accumulate;
  float alpha;
  int ps_h, ps_w, ps_p, ps_c; /* pixel shuffle geometry: token grid h x w, patch p, channels c */
} PkEpilogue;

int pk_gemm_bf16(const void* A, const void* B, int M, int N, int K, int lda, int ldb, int transA,
                 int transB, const PkEpilogue* epi, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PAINTER_B200_H */
