// extern "C" entry points of libcvflow.so (declared in include/cvflow.h).
#include "../../include/cvflow.h"
#include "gemm.h"
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

namespace cvflow {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
char* error_buf() { return g_err; }
int error_buf_len() { return (int)sizeof(g_err); }
}  // namespace cvflow

using namespace cvflow;

extern "C" CVFLOW_API const char* cvflow_last_error(void) { return g_err; }
extern "C" CVFLOW_API int cvflow_abi_version(void) { return 1; }

extern "C" CVFLOW_API int cvflow_gemm(const cvflow_gemm_desc* d, void* stream) {
  if (!d) { set_error("cvflow_gemm: null desc"); return CVFLOW_ERR_ARG; }
  GemmArgs a;
  for (int s = 0; s < 2; ++s) {
    a.A[s] = d->A[s]; a.a_rows[s] = d->a_rows[s]; a.a_cols[s] = d->a_cols[s];
    a.a_ld[s] = d->a_ld[s]; a.a_bstride[s] = d->a_bstride[s];
  }
  a.nbatch = d->nbatch; a.bf16 = d->dtype == CVFLOW_DTYPE_BF16;
  a.W = d->W; a.N = d->N; a.Ktot = d->Ktot; a.nseg = d->nseg;
  for (int s = 0; s < 8; ++s) {
    a.seg[s].a_map = d->seg[s].a_map; a.seg[s].row_shift = d->seg[s].row_shift;
    a.seg[s].a_col0 = d->seg[s].a_col0; a.seg[s].nkb = d->seg[s].nkb;
  }
  a.R = d->R; a.rmul = d->rmul; a.roff = d->roff; a.out_rows = d->out_rows;
  a.out = d->out; a.out_f32 = d->out_f32; a.transposed_out = d->transposed_out; a.ldc = d->ldc;
  a.col_off = d->col_off; a.n_valid = d->n_valid; a.alpha = d->alpha; a.act = d->act;
  a.bias = d->bias; a.aux_out = d->aux_out; a.mul_src = d->mul_src; a.ld_aux = d->ld_aux;
  a.rowmask = d->rowmask; a.resid = d->resid; a.ldr = d->ldr;
  GemmParams p;
  if (gemm_prepare(a, &p, error_buf(), error_buf_len())) return CVFLOW_ERR_ARG;
  int r = gemm_launch(p, (cudaStream_t)stream);
  if (r) { set_error("cvflow_gemm: launch failed: %s", cudaGetErrorString((cudaError_t)(-r))); return CVFLOW_ERR_CUDA; }
  return CVFLOW_OK;
}
