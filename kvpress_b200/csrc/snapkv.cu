// snapkv.cu — placeholder until the tcgen05 kernel lands (returns "unsupported", never a fallback).
#include "common.cuh"
namespace kvp {
size_t snapkv_scratch_bytes(const Dims&, int) { return 0; }
cudaError_t launch_snapkv_score(const Dims&, int, const void*, const void*, int, int,
                                const Workspace&, void*, bool, cudaStream_t) {
    return cudaErrorNotSupported;
}
}  // namespace kvp
