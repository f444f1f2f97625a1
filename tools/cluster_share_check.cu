// cluster_share_check.cu -- host check of the share arithmetic of the cluster solver (csrc/solver_cluster.cuh): for random
// block layouts, the columns of the blocks a CTA owns (first block at or after column 32 * share(q), up to the first block
// of the next share) must lie inside the groups [share(q), share_end(q)) the host sizes the CTA's shared memory for, the
// shares must tile all blocks, and row shares must tile all rows.  Test infrastructure (tests/test_pava_host.py).
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I block-simplex-least-squares_b200/csrc -o cluster_share_check tools/cluster_share_check.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "solver_cluster.cuh"

static uint64_t s_rng = 0x9E3779B97F4A7C15ull;
static inline uint64_t rnd() {
    s_rng ^= s_rng << 13;
    s_rng ^= s_rng >> 7;
    s_rng ^= s_rng << 17;
    return s_rng;
}

int main(int argc, char **argv) {
    const long reps = argc > 1 ? atol(argv[1]) : 20000;
    using namespace bsls;
    for (long it = 0; it < reps; ++it) {
        const int nb = 1 + (int)(rnd() % 400);
        const int kmax = 1 + (int)(rnd() % 64);
        std::vector<int> starts(nb + 1);
        int max_k = 0;
        starts[0] = 0;
        for (int b = 0; b < nb; ++b) {
            int k = 1 + (int)(rnd() % kmax);
            if (rnd() % 16 == 0) k = kmax;
            max_k = k > max_k ? k : max_k;
            starts[b + 1] = starts[b] + k;
        }
        const int n = starts[nb], groups = (n + 31) / 32;
        int prev_hi = 0;
        for (int q = 0; q < kClusterCtas; ++q) {
            int lohi[2];
            for (int t = 0; t < 2; ++t) {  // the kernel's search: first block with start >= 32 * share(q + t), else nb
                const int target = 32 * cluster_share(groups, q + t);
                int lo = 0, hi = nb;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (starts[mid] >= target)
                        hi = mid;
                    else
                        lo = mid + 1;
                }
                lohi[t] = lo;
            }
            if (lohi[0] != prev_hi) {
                fprintf(stderr, "shares do not tile the blocks: q=%d b_lo=%d expected %d (nb=%d n=%d)\n", q, lohi[0], prev_hi, nb, n);
                return 1;
            }
            prev_hi = lohi[1];
            const int c_lo = starts[lohi[0]], c_hi = starts[lohi[1]];
            if (c_hi > c_lo) {
                const int g0 = cluster_share(groups, q), g1 = cluster_share_end(groups, q, max_k);
                if (c_lo < 32 * g0 || c_hi > 32 * g1 || g1 > groups) {
                    fprintf(stderr, "share %d: columns [%d, %d) outside groups [%d, %d) (nb=%d n=%d max_k=%d)\n", q, c_lo, c_hi, g0, g1, nb, n, max_k);
                    return 1;
                }
            }
        }
        if (prev_hi != nb) {
            fprintf(stderr, "last share ends at block %d of %d\n", prev_hi, nb);
            return 1;
        }
        const int m = 1 + (int)(rnd() % 3000), ga = (m + 31) / 32;
        int rows = 0;
        for (int q = 0; q < kClusterCtas; ++q) {
            const int r0 = 32 * cluster_share(ga, q), r1 = 32 * cluster_share(ga, q + 1);
            const int hi = r1 < m ? r1 : m;
            rows += hi > r0 ? hi - r0 : 0;
        }
        if (rows != m) {
            fprintf(stderr, "row shares cover %d of %d rows\n", rows, m);
            return 1;
        }
    }
    printf("ok %ld layouts\n", reps);
    return 0;
}
