/* cvflow C ABI — the drop-in boundary of the B200-native flow-LoRA hot path.
 *
 * Plain C: raw device pointers, sizes, an explicit cudaStream_t (passed as void*), int status
 * codes (0 = ok, <0 = error; text via cvflow_last_error()). No torch types cross this boundary.
 *
 * The reference (leeoisaboy/cosyvoice-lora-finetune-framework) has no FFI of its own for this
 * path; its only precedent for a non-PyTorch estimator is the TensorRT hook of the vendored
 * upstream, which binds raw data_ptr()s by tensor name and runs on the caller's stream
 * (cosyvoice/flow/flow_matching.py:125-152, cosyvoice/utils/common.py:171-186). The entry points
 * below follow that calling convention; each one cites the reference code it replaces.
 * INTEGRATION.md shows the ctypes stub a maintainer of the reference would add.
 */
#ifndef CVFLOW_H_
#define CVFLOW_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CVFLOW_API __attribute__((visibility("default")))
#else
#define CVFLOW_API
#endif

#define CVFLOW_OK 0
#define CVFLOW_ERR_ARG (-1)
#define CVFLOW_ERR_CUDA (-2)
#define CVFLOW_ERR_UNSUPPORTED (-3)

#define CVFLOW_DTYPE_F16 0
#define CVFLOW_DTYPE_BF16 1

/* Thread-local text of the last error returned by any cvflow_* call on this thread. */
CVFLOW_API const char* cvflow_last_error(void);
/* ABI version of this library (bumped on any signature change). */
CVFLOW_API int cvflow_abi_version(void);

/* ---------------------------------------------------------------------------------------------
 * Dense contraction engine (tcgen05 / TMEM / TMA). Exposed so each fused linear / conv of the
 * estimator can be parity-tested on its own. Replaces F.linear / nn.Conv1d / nn.ConvTranspose1d
 * as called from modules.py:65,87,101,112,138,218,266-268,291,943,981 and lora.py:66-74.
 * ------------------------------------------------------------------------------------------- */
typedef struct cvflow_gemm_seg {
  int32_t a_map;     /* A source 0/1 */
  int32_t row_shift; /* row offset of this tap (rows outside the source read as zero) */
  int32_t a_col0;    /* first source column */
  int32_t nkb;       /* number of 64-column blocks */
} cvflow_gemm_seg;

typedef struct cvflow_gemm_desc {
  const void* A[2];     /* 16-bit [nbatch][a_rows][a_cols] */
  int32_t a_rows[2];
  int32_t a_cols[2];
  int64_t a_ld[2];      /* row stride, elements */
  int64_t a_bstride[2]; /* batch stride, elements */
  int32_t nbatch;
  int32_t dtype;        /* CVFLOW_DTYPE_* of A, W and 16-bit outputs */
  const void* W;        /* [N][Ktot] row-major 16-bit */
  int32_t N;
  int32_t Ktot;
  cvflow_gemm_seg seg[8];
  int32_t nseg;
  int32_t R;            /* tile rows per batch */
  int32_t rmul, roff;   /* output row = i*rmul + roff */
  int32_t out_rows;     /* rows per batch of the output */
  void* out;
  int32_t out_f32;
  int32_t transposed_out; /* fp32 out[(b*n_valid+n)*out_rows + row] */
  int64_t ldc;
  int32_t col_off;
  int32_t n_valid;
  float alpha;
  int32_t act;          /* 0 none, 1 gelu-tanh, 2 gelu-erf, 3 *gelu-tanh'(mul_src), 4 *gelu-erf'(mul_src) */
  const float* bias;
  void* aux_out;
  const void* mul_src;
  int64_t ld_aux;
  const float* rowmask;
  const float* resid;
  int64_t ldr;
} cvflow_gemm_desc;

CVFLOW_API int cvflow_gemm(const cvflow_gemm_desc* desc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CVFLOW_H_ */
