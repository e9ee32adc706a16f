// Cluster kernel (single launch, register-resident state, DSMEM halo exchange) -- placeholder
// until the kernel lands; AUTO falls back to the generic path.
#include "common.cuh"
namespace cspn {
bool cluster2d_supported(const Problem2D&, char* why, int why_len) {
    snprintf(why, why_len, "cluster kernel not built yet");
    return false;
}
int cluster2d_forward(const Problem2D&, cudaStream_t, int*) {
    set_error("cluster kernel not built yet");
    return CSPN_ERR_UNSUPPORTED;
}
int cluster2d_describe(int, int, int, int, int, char* buf, int len) { return snprintf(buf, len, "cluster: n/a"); }
}  // namespace cspn
