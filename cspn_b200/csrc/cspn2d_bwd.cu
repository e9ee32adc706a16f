// 2D backward (adjoint) -- placeholder.
#include "common.cuh"
namespace cspn {
size_t bwd2d_workspace_bytes(int, int, int, int, int) { return 0; }
int bwd2d(const Problem2D&, const float*, float*, float*, void*, size_t, cudaStream_t, int*) {
    set_error("native backward not built yet");
    return CSPN_ERR_UNSUPPORTED;
}
}  // namespace cspn
