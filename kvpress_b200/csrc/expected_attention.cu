// expected_attention.cu — placeholder until the tcgen05 kernel lands.
#include "common.cuh"
namespace kvp {
size_t ea_scratch_bytes(const Dims&) { return 0; }
cudaError_t launch_ea_score(const Dims&, int, const void*, const void*, const void*, const void*,
                            float, int, int, const Workspace&, void*, bool, cudaStream_t) {
    return cudaErrorNotSupported;
}
}  // namespace kvp
