// proj_f64.cu -- fp64 instantiations of the projection kernels (the reference's dtype).
#include "kernels.h"
#include "proj_ragged.cuh"

namespace bsls {
int proj_uniform_f64(double *y, long long first, int nb, int K, int mode, int32_t *slow, cudaStream_t stream) {
    return mode == kBall ? launch_proj_uniform<double, kBall>(y, first, nb, K, slow, stream)
                         : launch_proj_uniform<double, kSimplex>(y, first, nb, K, slow, stream);
}

bool proj_step_fuses(int K) { return proj_uniform_fuses(K); }

int proj_step_uniform_f64(const double *x, const double *g, double t, double *x_new, long long first, int nb, int K, int mode,
                          cudaStream_t stream, const StepCtl *ctl) {
    double *xin = const_cast<double *>(x);
    return mode == kBall ? launch_proj_uniform<double, kBall>(xin, first, nb, K, nullptr, stream, g, t, x_new, ctl)
                         : launch_proj_uniform<double, kSimplex>(xin, first, nb, K, nullptr, stream, g, t, x_new, ctl);
}

int proj_ragged_f64(double *y, const int32_t *starts, const int32_t *tile_first, int ntiles, const int32_t *mid_ids, int nmid,
                    const int32_t *large_ids, int nlarge, int max_large, int mode, int32_t *slow, int nb, const RaggedStreams &rs,
                    void *huge_buf, int huge_cap, int *huge_lock, cudaStream_t stream) {
    static_assert(kTileElems == kPlanTileElems && kTileMaxBlock == kPlanTileMaxBlock && kLargeMaxBlock == kPlanLargeMaxBlock &&
                      kTileThreadMax == kPlanMidMin, "plan constants");
    return mode == kBall ? launch_proj_ragged<double, kBall>(y, starts, tile_first, ntiles, mid_ids, nmid, large_ids, nlarge, max_large, slow, nb, rs, huge_buf, huge_cap, huge_lock, stream)
                         : launch_proj_ragged<double, kSimplex>(y, starts, tile_first, ntiles, mid_ids, nmid, large_ids, nlarge, max_large, slow, nb, rs, huge_buf, huge_cap, huge_lock, stream);
}
}  // namespace bsls
