// flat_gemm.cu — K2 placeholder (tensor-core batched Flat); see DESIGN.md
#include "dataset.cuh"
namespace vdb {
bool flat_gemm_supported(const vdb_dataset*, uint32_t, uint32_t) { return false; }
void flat_gemm_keys(const vdb_dataset*, const void*, uint32_t, uint32_t, uint64_t*, cudaStream_t) {
    fail(VDB_EUNSUPPORTED, "tensor-core Flat path not built");
}
}  // namespace vdb
