// placeholder until the tcgen05 engine lands: AUTO resolves to the exact SIMT engine.
#include "common.cuh"
#include "l2_internal.cuh"
namespace hpusm {
bool l2_tc_available() { return false; }
size_t l2_tc_workspace_bytes(int64_t, int64_t, int, int) { return 256; }
int l2_tc_knn2(const float*, int64_t, const float*, int64_t, int, int, int32_t*, float*, void*, size_t, cudaStream_t,
               int*, int64_t*) {
  set_error("tcgen05 L2 engine not built");
  return HPUSM_ERR_UNSUPPORTED;
}
}  // namespace hpusm
