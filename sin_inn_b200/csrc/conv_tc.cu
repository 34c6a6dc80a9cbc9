// tcgen05 / TMEM / TMA implicit-GEMM convolution path (bf16 operands, fp32 accumulate).
#include "common.cuh"

namespace sininn {
int wgrad_simt_splits(const sininn_wgrad_desc* d);
}
using namespace sininn;

extern "C" {

size_t sininn_wgrad_workspace_bytes(const sininn_wgrad_desc* d, int tensor_core) {
  if (!d || d->Cin <= 0 || d->Cout <= 0 || d->taps <= 0) return 0;
  (void)tensor_core;
  return (size_t)wgrad_simt_splits(d) * d->taps * d->Cout * d->Cin * sizeof(float);
}

int sininn_conv_tc(const sininn_conv_desc* d, sininn_stream_t stream) {
  (void)d; (void)stream;
  set_error("conv_tc: not built yet");
  return SININN_EUNSUPPORTED;
}

int sininn_wgrad_tc(const sininn_wgrad_desc* d, sininn_stream_t stream) {
  (void)d; (void)stream;
  set_error("wgrad_tc: not built yet");
  return SININN_EUNSUPPORTED;
}

}  // extern "C"
