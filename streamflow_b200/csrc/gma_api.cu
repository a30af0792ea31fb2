// TEMPORARY stub of the GMA entry points (replaced by the real implementation).
#include "sf_internal.h"
using namespace sf;
extern "C" {
int64_t sf_gma_npad(int64_t N) { return (N + 63) / 64 * 64; }
int64_t sf_gma_workspace_bytes(int64_t, int64_t, int64_t, int64_t) { return 0; }
int sf_gma_attention(const void*, int, const float*, int64_t, int64_t, int64_t, int64_t, float, void*, float*, void*,
                     int64_t, void*) {
    set_error("sf_gma_attention: not implemented yet");
    return SF_ERR_INVALID;
}
int sf_gma_aggregate(const void*, const float*, const void*, int, const float*, const float*, float*, int64_t, int64_t,
                     int64_t, int64_t, void*, int64_t, void*) {
    set_error("sf_gma_aggregate: not implemented yet");
    return SF_ERR_INVALID;
}
}
