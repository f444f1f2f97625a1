// proj_f32.cu -- fp32 twins (an extension: the reference is fp64 only; tolerance 1e-4).
#include "kernels.h"
#include "proj_ragged.cuh"

namespace bsls {
int proj_uniform_f32(float *y, long long first, int nb, int K, int mode, int32_t *slow, cudaStream_t stream) {
    return mode == kBall ? launch_proj_uniform<float, kBall>(y, first, nb, K, slow, stream)
                         : launch_proj_uniform<float, kSimplex>(y, first, nb, K, slow, stream);
}

int proj_ragged_f32(float *y, const int32_t *starts, const int32_t *tile_first, int ntiles, const int32_t *mid_ids, int nmid,
                    const int32_t *large_ids, int nlarge, int max_large, int mode, int32_t *slow, int nb, const RaggedStreams &rs,
                    void *huge_buf, int huge_cap, int *huge_lock, cudaStream_t stream) {
    static_assert(kTileElems == kPlanTileElems && kTileMaxBlock == kPlanTileMaxBlock && kLargeMaxBlock == kPlanLargeMaxBlock &&
                      kTileThreadMax == kPlanMidMin, "plan constants");
    return mode == kBall ? launch_proj_ragged<float, kBall>(y, starts, tile_first, ntiles, mid_ids, nmid, large_ids, nlarge, max_large, slow, nb, rs, huge_buf, huge_cap, huge_lock, stream)
                         : launch_proj_ragged<float, kSimplex>(y, starts, tile_first, ntiles, mid_ids, nmid, large_ids, nlarge, max_large, slow, nb, rs, huge_buf, huge_cap, huge_lock, stream);
}
}  // namespace bsls
