// conv_mma.cu — kernel B (tcgen05 INT8 MMA path).  Placeholder until the tensor-core kernel lands:
// reports every layer as unsupported so the plan uses the shift kernel.
#include <string>

#include "../../include/tf2b200.h"
#include "common.cuh"

namespace tf2b {
bool mma_layer_supported(const tf2b_layer_desc&, int, int) { return false; }
cudaError_t launch_conv_mma(const ConvParams&, const int8_t*, int, const int*, void*, cudaStream_t) {
  return cudaErrorNotSupported;
}
size_t mma_tmap_bytes() { return 0; }
int mma_build_tmaps(void*, const ConvParams&, const int8_t*, int, std::string* err) {
  if (err) *err = "tensor-core path not built";
  return -1;
}
int mma_bn() { return 64; }
int mma_bk() { return 64; }
}  // namespace tf2b
