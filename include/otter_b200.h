/* otter_b200 — C ABI of the B200-native vision-fusion hot path.
 *
 * The reference (Luodian/Otter) ships no native code and therefore has no FFI; the operations
 * below are the torch calls its hot-path modules make, restated as a C ABI so that a Python host
 * (ctypes, see otter_b200/_lib.py and INTEGRATION.md) can bind them.  Each entry point cites the
 * reference lines it replaces (paths relative to /root/reference).
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless stated otherwise; the caller owns every buffer
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream)
 *   - return value: 0 = ok, non-zero = error; otb_last_error() returns a thread-local message
 *   - no entry point allocates device memory, synchronises the device, or throws
 *   - bf16 = IEEE bfloat16 (2 bytes); "K-major" = reduction dimension contiguous in memory
 */
#ifndef OTTER_B200_H
#define OTTER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OTB_OK 0
#define OTB_ERR_INVALID 1
#define OTB_ERR_CUDA 2
#define OTB_ERR_UNSUPPORTED 3

const char* otb_last_error(void);
/* Library/ABI version and the SM architecture the kernels were compiled for (100 = sm_100a). */
int otb_version(void);
int otb_compiled_arch(void);
/* Number of kernels this library has launched in the calling process (bench.py "gpu_launches"). */
long long otb_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * GEMM:  D[M,N] = epilogue( A[M,K] . B[N,K]^T )      bf16 operands, fp32 accumulate in TMEM
 * Replaces every nn.Linear / F.linear on the path (modeling_otter.py:139-148,164-167,180-184,
 * 253-256,284-288,340,363-370; xformers_model/clip.py:106-134,145-149) and their autograd
 * dgrad/wgrad.  tcgen05.mma + TMA, persistent, warp-specialised.
 *
 * Operand layouts (ld* in elements):
 *   a_mn_major = 0 : A is stored [M][K] row-major (lda >= K)         (activations, dY for dgrad)
 *   a_mn_major = 1 : A is stored [K][M] row-major (lda >= M)         (dY^T for wgrad, no transpose copy)
 *   b_mn_major = 0 : B is stored [N][K] row-major (ldb >= K)         (nn.Linear weight for forward)
 *   b_mn_major = 1 : B is stored [K][N] row-major (ldb >= N)         (weight for dgrad, X for wgrad)
 * Requirements: lda, ldb, ld_out, ld_aux, ld_res multiples of 8; pointers 16-byte aligned;
 *   N multiple of 8; for an MN-major operand its MN extent must be a multiple of 8.
 *
 * Epilogue (all optional, applied in this order on the fp32 accumulator v):
 *   v += bias[n]                                   (fp32 [N])
 *   aux_out[m,n] = bf16(v)                         (pre-activation, kept for backward)
 *   v = act(v)             act: 0 none, 1 GELU(erf) (modeling_otter.py:146,367), 2 quick-GELU (CLIP)
 *   v *= gelu'(aux_in[m,n])                        (backward of GELU, aux_in = saved pre-activation)
 *   v *= alpha * (scale_ptr ? (scale_tanh ? tanh(*scale_ptr) : *scale_ptr) : 1)
 *                                                  (tanh gate, modeling_otter.py:387-388,393)
 *   v += residual[m,n]                             (bf16)
 *   out[m,n] = (accumulate ? out[m,n] : 0) + v     (bf16, or fp32 when out_fp32; accumulate needs fp32)
 * ------------------------------------------------------------------------------------------- */
typedef struct otb_gemm_epilogue {
  const float* bias;
  const void* aux_in;
  void* aux_out;
  const float* scale_ptr;
  const void* residual;
  void* out;
  int64_t ld_out, ld_aux_in, ld_aux_out, ld_res;
  int32_t act;
  int32_t scale_tanh;
  int32_t out_fp32;
  int32_t accumulate;
  float alpha;
  int32_t _pad;
} otb_gemm_epilogue;

int otb_gemm_bf16(const void* A, int a_mn_major, int64_t lda, const void* B, int b_mn_major, int64_t ldb, int M,
                  int N, int K, const otb_gemm_epilogue* epi, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OTTER_B200_H */
