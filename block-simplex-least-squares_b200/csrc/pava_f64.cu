// pava_f64.cu -- double instantiation of the segmented isotonic regression kernels.
#include "kernels.h"
#include "pava.cuh"
#include "pava_words.cuh"

namespace bsls {
int pava_f64(double *y, int32_t *w, const int32_t *starts, const int32_t *win_first, int nwin, const int32_t *large_ids,
             int nlarge, int max_large, int update, int clip01, cudaStream_t stream) {
    static_assert(kPavaPitch == kPlanPavaPitch && kPavaWarpMaxBlock == kPlanPavaWarpMax && kPavaLargeMaxBlock == kPlanPavaLargeMax, "plan constants");
    PavaFlags fl;
    fl.update = update;
    fl.clip01 = clip01;
    fl.has_weight = w != nullptr;
    return launch_pava<double>(y, w, starts, win_first, nwin, large_ids, nlarge, max_large, fl, stream);
}

int pava_small_f64(double *y, int32_t *w, long long first, int nb, int K, int update, int clip01, cudaStream_t stream) {
    PavaFlags fl;
    fl.update = update;
    fl.clip01 = clip01;
    fl.has_weight = w != nullptr;
    return launch_pava_small<double>(y, w, first, nb, K, fl, stream);
}

int pava_tile_f64(double *y, int32_t *w, const int32_t *starts, const int32_t *tile_first, int ntiles, int update, int clip01,
                  int cap_per_sm, cudaStream_t stream) {
    static_assert(kPavaTileElems == kPlanTileElems && kPavaTileMaxBlock == kPlanTileMaxBlock && kPavaThreadMax == kPlanMidMin, "plan constants");
    PavaFlags fl;
    fl.update = update;
    fl.clip01 = clip01;
    fl.has_weight = w != nullptr;
    if (!getenv("BSLS_PAVA_NO_ROWS")) return launch_pava_tile_rows<double>(y, w, starts, tile_first, ntiles, update, clip01, cap_per_sm, stream);
    return launch_pava_tile<double>(y, w, starts, tile_first, ntiles, fl, stream);
}

int pava_mid_f64(double *y, int32_t *w, const int32_t *starts, const int32_t *mid_ids, int nmid, int update, int clip01, cudaStream_t stream) {
    PavaFlags fl;
    fl.update = update;
    fl.clip01 = clip01;
    fl.has_weight = w != nullptr;
    return launch_pava_mid<double>(y, w, starts, mid_ids, nmid, fl, stream);
}

int pava_words_f64(double *y, int32_t *w, const int32_t *starts, const int32_t *ids, const int32_t *pack_first, int npacks, long long first, int nb,
                   int Kuni, int update, int clip01, int cap_per_sm, cudaStream_t stream) {
    static_assert(kWordsMaxBlock == kPlanWordsMax, "plan constants");
    return launch_pava_words<double>(y, w, starts, ids, pack_first, npacks, first, nb, Kuni, update, clip01, cap_per_sm, stream);
}

int pava_words_cta_f64(double *y, int32_t *w, const int32_t *starts, const int32_t *ids, int count, int max_block, int update, int clip01, int cap_per_sm,
                       cudaStream_t stream) {
    static_assert(32 * kWordsCtaThreads == kPlanPavaLargeMax, "plan constants");
    return launch_pava_words_cta<double>(y, w, starts, ids, count, max_block, update, clip01, cap_per_sm, stream);
}
}  // namespace bsls
