// attention backward (placeholder until the tcgen05 dK/dV and dQ kernels land)
#include "kernels.h"
#include <stdio.h>
namespace cvflow {
int attn_bwd_prepare(void*, const void*, const void*, int, int, int, char* err, int errlen) {
  if (err) snprintf(err, errlen, "attention backward not built yet");
  return -1;
}
int attn_bwd_launch(const void*, const float*, int, const void*, const float*, float*, void*, cudaStream_t) { return -1; }
}  // namespace cvflow
