// kernels.h -- internal launch interface between the translation units of libbsls_b200.
#pragma once
#include "common.cuh"

namespace bsls {

// proj_f64.cu / proj_f32.cu: all blocks have K entries, first block starts at `first`.
// `slow`: nb + 1 int32 of scratch for the blocks the selection kernel hands to the sorter (may be null: sorting kernels only)
int proj_uniform_f64(double *y, long long first, int nb, int K, int mode, int32_t *slow, cudaStream_t stream);
int proj_uniform_f32(float *y, long long first, int nb, int K, int mode, int32_t *slow, cudaStream_t stream);
// fused projected-gradient step x_new = proj(x - t g), uniform layouts whose K the sorting kernels take
bool proj_step_fuses(int K);
// ctl (device memory, may be null): the step is read from *ctl->t instead of `t`, and the kernel exits when *ctl->done
int proj_step_uniform_f64(const double *x, const double *g, double t, double *x_new, long long first, int nb, int K, int mode,
                          cudaStream_t stream, const StepCtl *ctl = nullptr);


// fork/join inside one call: the tile, mid and large kernels own disjoint blocks and run side by side
struct RaggedStreams {
    cudaStream_t aux[2];
    cudaEvent_t fork, join[2];
};
// ragged layouts: tile kernel (thread per block) + warp-per-block kernel + one-CTA-per-large-block kernel (proj_ragged.cuh)
int proj_ragged_f64(double *y, const int32_t *starts, const int32_t *tile_first, int ntiles, const int32_t *mid_ids, int nmid,
                    const int32_t *large_ids, int nlarge, int max_large, int mode, int32_t *slow, int nb, const RaggedStreams &rs,
                    void *huge_buf, int huge_cap, int *huge_lock, cudaStream_t stream);
int proj_ragged_f32(float *y, const int32_t *starts, const int32_t *tile_first, int ntiles, const int32_t *mid_ids, int nmid,
                    const int32_t *large_ids, int nlarge, int max_large, int mode, int32_t *slow, int nb, const RaggedStreams &rs,
                    void *huge_buf, int huge_cap, int *huge_lock, cudaStream_t stream);

// plan.cu: layout analysis on the device
struct LayoutStats {
    int min_size, max_size, bad;  // bad != 0: not strictly increasing / out of range
};
int plan_layout_stats(const int32_t *starts /* nb+1 */, int nb, int n, LayoutStats *d_out, cudaStream_t stream);
int plan_tile_first(const int32_t *starts, int nb, int first, int pitch, int32_t *tile_first, int ntiles, cudaStream_t stream);
int plan_large_list(const int32_t *starts, int nb, int threshold, int32_t *ids, int *d_count, cudaStream_t stream,
                    int upper = 0x7fffffff);
constexpr int kPlanMidMin = 32;           // ragged layouts: blocks of kPlanMidMin < size <= kPlanTileMaxBlock get a warp each
constexpr int kPlanTileElems = 2048;      // == kTileElems (proj_ragged.cuh)
constexpr int kPlanTileMaxBlock = 512;    // == kTileMaxBlock
constexpr int kPlanLargeMaxBlock = 8192;  // == kLargeMaxBlock
constexpr int kPlanPavaLargeMax = 8192;   // == kPavaLargeMaxBlock: longest block the isotonic regression takes

// pava_f64.cu / pava_f32.cu: segmented isotonic regression (pava.cuh, pava_words.cuh).  w: weight array (in / out) or nullptr.
// uniform layouts with K <= kPlanPavaSmallMax: one thread per row of blocks
constexpr int kPlanPavaSmallMax = 64;
int pava_small_f64(double *y, int32_t *w, long long first, int nb, int K, int update, int clip01, cudaStream_t stream);
int pava_small_f32(float *y, int32_t *w, long long first, int nb, int K, int update, int clip01, cudaStream_t stream);
// ragged layouts: tiles of whole blocks of at most kPlanMidMin entries; longer ones go to the two kernels below
int pava_tile_f64(double *y, int32_t *w, const int32_t *starts, const int32_t *tile_first, int ntiles, int update, int clip01,
                  int cap_per_sm, cudaStream_t stream);
int pava_tile_f32(float *y, int32_t *w, const int32_t *starts, const int32_t *tile_first, int ntiles, int update, int clip01,
                  int cap_per_sm, cudaStream_t stream);
// blocks of 33 .. kPlanWordsMax entries, cold start: one lane per 32-entry word, packs of blocks per warp (pava_words.cuh).
// ragged: ids / pack_first from plan_pack_words, first = 0; uniform: starts = ids = pack_first = nullptr.
constexpr int kPlanWordsMax = 1024;
// cap_per_sm > 0 limits the grid to that many CTAs per SM (kernels meant to be co-resident with others)
// w: weight array (in / out, warm start) or nullptr (cold start)
int pava_words_f64(double *y, int32_t *w, const int32_t *starts, const int32_t *ids, const int32_t *pack_first, int npacks, long long first, int nb,
                   int Kuni, int update, int clip01, int cap_per_sm, cudaStream_t stream);
int pava_words_f32(float *y, int32_t *w, const int32_t *starts, const int32_t *ids, const int32_t *pack_first, int npacks, long long first, int nb,
                   int Kuni, int update, int clip01, int cap_per_sm, cudaStream_t stream);
// one CTA per block of up to kPlanPavaLargeMax entries (cold start), same engine
int pava_words_cta_f64(double *y, int32_t *w, const int32_t *starts, const int32_t *ids, int count, int max_block, int update, int clip01,
                       int cap_per_sm, cudaStream_t stream);
int pava_words_cta_f32(float *y, int32_t *w, const int32_t *starts, const int32_t *ids, int count, int max_block, int update, int clip01,
                       int cap_per_sm, cudaStream_t stream);
// pava_seq.cuh: the reference's routines (variant 1, 2 or 3) as written, one thread per block; only blocks longer than
// min_size entries; cold != 0: w is scratch, set to ones first
int pava_seq_f64(int variant, double *y, int32_t *w, const int32_t *starts, const int32_t *ids, int count, int min_size, int update, int cold,
                 int clip, cudaStream_t stream);
int pava_seq_f32(int variant, float *y, int32_t *w, const int32_t *starts, const int32_t *ids, int count, int min_size, int update, int cold,
                 int clip, cudaStream_t stream);
// greedy packing of ids[0..count) into packs of at most 32 words (32 entries each); pack_first has count + 1 slots
int plan_pack_words(const int32_t *starts, const int32_t *ids, int count, int32_t *pack_first, int *d_npacks, cudaStream_t stream);

int device_ok();  // BSLS_OK when an sm_100 device is current (capi.cu)
}  // namespace bsls

// The analysed block layout behind the opaque handle of the C ABI (built in capi.cu).
struct bsls_plan {
    int nb = 0, n = 0, first = 0;
    int uniform = 0, min_size = 0, max_size = 0;
    int32_t *d_starts = nullptr;      // nb + 1 entries, last = n
    // ragged layouts only
    int tiles = 0, large = 0;
    int32_t *d_tile_first = nullptr;  // tiles + 1 entries
    int32_t *d_large_ids = nullptr;   // `large` block indices (size > kPlanTileMaxBlock)
    bool ragged = false;
    int mid = 0;
    int32_t *d_mid_ids = nullptr;     // ragged: blocks with kPlanMidMin < size <= kPlanTileMaxBlock
    int mid_packs = 0;                // packs of the mid list for pava_words
    int32_t *d_mid_pack = nullptr;    // mid_packs + 1 entries
    // fork/join inside one call: the tile, mid and large kernels own disjoint blocks and run side by side
    cudaStream_t aux[2] = {nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
    void *d_huge = nullptr;           // blocks longer than kPlanLargeMaxBlock: global scratch for their candidates (2 * huge_cap values)
    int huge_cap = 0;
    int *d_huge_lock = nullptr;
    int32_t *d_slow = nullptr;        // nb + 1: queue of dense blocks between the selection kernel and the sorter
    int32_t *d_seq_w = nullptr;       // n: weight scratch of the sequential isotonic regression (layouts with blocks > kPlanPavaLargeMax)
};

namespace bsls {
// capi.cu: the device entry points, for the solver translation units
int project_f64(const bsls_plan *plan, double *y, int mode, cudaStream_t stream);
int pava_clip_f64(const bsls_plan *plan, double *y, int32_t *weight, int update, int clip01, cudaStream_t stream);
// x_new = proj(x - t g): one fused kernel where the layout allows it, else returns 1 ("not fused") and does nothing
int project_step_f64(const bsls_plan *plan, const double *x, const double *g, double t, double *x_new, int mode, cudaStream_t stream, bool *fused,
                     const StepCtl *ctl = nullptr);
}  // namespace bsls

