// pava_f64.cu -- double instantiation of the segmented isotonic regression kernels.
#include "kernels.h"
#include "pava.cuh"
#include "pava_words.cuh"
#include "pava_seq.cuh"

namespace bsls {
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
    return launch_pava_tile_rows<double>(y, w, starts, tile_first, ntiles, update, clip01, cap_per_sm, stream);
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

int pava_seq_f64(int variant, double *y, int32_t *w, const int32_t *starts, const int32_t *ids, int count, int min_size, int update, int cold,
                 int clip, cudaStream_t stream) {
    return launch_pava_seq<double>(variant, y, w, starts, ids, count, min_size, update, cold, clip, stream);
}
}  // namespace bsls
