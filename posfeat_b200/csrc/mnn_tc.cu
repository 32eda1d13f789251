// placeholder until the tcgen05 kernel lands
#include "common.cuh"
#include "mnn_common.cuh"
namespace posfeat {
bool tc_supported(int, int, int) { return false; }
size_t tc_workspace_bytes(int, int) { return 0; }
int mnn_tc(const float*, int, int64_t, const float*, int, int64_t, int, int32_t*, int32_t*, void*, size_t, cudaStream_t) {
  return set_error(POSFEAT_EUNSUPPORTED, "tensor-core matcher not built");
}
}  // namespace posfeat
