// pava_host_check.cu -- runs the thread-per-block PAVA routine of csrc/pava.cuh ON THE CPU against the
// oracle (oracle/liboracle.so: orc_pava, the restatement of isotonic_regression.h:13-58), bit for bit:
// values, pool sizes and the stale interior entries.  Test infrastructure (tests/test_pava_host.py).
//   nvcc -O2 -std=c++17 -fmad=false -Xcompiler -ffp-contract=off -o pava_host_check tools/pava_host_check.cu -Loracle -loracle
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../block-simplex-least-squares_b200/csrc/pava.cuh"

extern "C" void orc_pava(double *y, int64_t lo, int64_t hi, int32_t *sz, int spread);

static uint64_t s_rng = 88172645463325252ull;
static inline uint64_t rnd() {
    s_rng ^= s_rng << 13;
    s_rng ^= s_rng >> 7;
    s_rng ^= s_rng << 17;
    return s_rng;
}
static inline double uni() { return (double)(rnd() >> 11) * (1.0 / 9007199254740992.0); }

template <typename M> static long run_case(int K, int kind, int warm, int update) {
    std::vector<double> y(K), y0(K), yr(K);
    std::vector<int32_t> wr(K, 1);
    std::vector<uint16_t> w(K, 1);
    for (int i = 0; i < K; ++i) {
        double v;
        switch (kind) {
            case 0: v = uni() * 2 - 1; break;                               // continuous
            case 1: v = (double)(int)(rnd() % 5); break;                    // heavy ties
            case 2: v = (double)((int)(rnd() % 100) - 50) + 50.0 * log(1.0 + i); break;  // the reference's test generator
            case 3: v = (double)i; break;                                   // isotonic, last = -1e12 (worst case)
            case 4: v = -(double)i * 0.1; break;                            // strictly decreasing
            case 5: v = (rnd() % 3 == 0) ? 0.0 : (double)((int)(rnd() % 3) - 1);  // zeros and signs
            default: v = uni() * 1e-3 + (double)(rnd() % 4) / 3.0; break;   // thirds
        }
        y[i] = v;
    }
    if (kind == 3) y[K - 1] = -1e12;
    y0 = y;
    yr = y;
    static double rcp[65];
    for (int i = 1; i <= 64; ++i) rcp[i] = 1.0 / i;
    const M full = K == (int)(8 * sizeof(M)) ? ~M(0) : (((M)1 << K) - 1);
    M heads;
    if (warm) {
        // first a cold pass on BOTH sides with update = 0 leaves a consistent warm state; perturb the values and run again
        orc_pava(yr.data(), 0, K, wr.data(), 0);
        std::vector<uint16_t> ww(K, 1);
        heads = bsls::pava_block_runs<double, uint16_t, M, false>(y.data(), ww.data(), K, full, (M)1, true, rcp, 65);
        for (int i = 0; i < K; ++i)
            if (memcmp(&y[i], &yr[i], 8) || ww[i] != (uint16_t)wr[i]) return 1 + i;
        for (int i = 0; i < K; ++i) {
            const double d = (kind == 1 || kind == 5) ? (double)((int)(rnd() % 3) - 1) : (uni() - 0.5);
            y[i] += d;
            yr[i] = y[i];
        }
        w = ww;
        orc_pava(yr.data(), 0, K, wr.data(), update);
        heads = bsls::pava_block_runs<double, uint16_t, M, true>(y.data(), w.data(), K, bsls::pava_heads_from_weights<uint16_t, M>(w.data(), K), (M)1, true, rcp, 65);
    } else {
        orc_pava(yr.data(), 0, K, wr.data(), update);
        heads = bsls::pava_block_runs<double, uint16_t, M, false>(y.data(), w.data(), K, full, (M)1, true, rcp, 65);
    }
    if (update) bsls::pava_spread(y.data(), K, heads);
    for (int i = 0; i < K; ++i)
        if (memcmp(&y[i], &yr[i], 8) || w[i] != (uint16_t)wr[i]) {
            fprintf(stderr, "mismatch K=%d kind=%d warm=%d update=%d at %d: %a/%d vs %a/%d\n", K, kind, warm, update, i, y[i], (int)w[i], yr[i], wr[i]);
            return 1 + i;
        }
    // the head mask must name exactly the heads the reference's weights chain through
    M chain = 0;
    for (int i = 0; i < K; i += wr[i] > 0 ? wr[i] : 1) chain |= (M)1 << i;
    if (chain != heads) {
        fprintf(stderr, "head mask mismatch K=%d kind=%d warm=%d\n", K, kind, warm);
        return -1;
    }
    return 0;
}

// several short blocks regressed as one row with forced run starts at the block boundaries (cold start)
template <typename M> static long run_group(int K, int G, int kind) {
    const int KR = K * G;
    std::vector<double> y(KR), yr(KR);
    std::vector<int32_t> wr(KR, 1);
    std::vector<uint16_t> w(KR, 1);
    for (int i = 0; i < KR; ++i) {
        double v;
        switch (kind) {
            case 0: v = uni() * 2 - 1; break;
            case 1: v = (double)(int)(rnd() % 5); break;
            case 2: v = (double)((int)(rnd() % 100) - 50) + 50.0 * log(1.0 + (i % K)); break;
            case 4: v = -(double)i * 0.1; break;   // decreasing ACROSS block boundaries: runs must stop there
            case 5: v = (rnd() % 3 == 0) ? 0.0 : (double)((int)(rnd() % 3) - 1); break;
            default: v = uni() * 1e-3 + (double)(rnd() % 4) / 3.0; break;
        }
        y[i] = v;
    }
    yr = y;
    static double rcp[65];
    for (int i = 1; i <= 64; ++i) rcp[i] = 1.0 / i;
    M bst = 0;
    for (int g = 0; g < G; ++g) {
        bst |= (M)1 << (g * K);
        orc_pava(yr.data(), g * K, (g + 1) * K, wr.data(), 1);
    }
    const M full = KR == (int)(8 * sizeof(M)) ? ~M(0) : (((M)1 << KR) - 1);
    const M heads = bsls::pava_block_runs<double, uint16_t, M, false>(y.data(), w.data(), KR, full, bst, true, rcp, 65);
    bsls::pava_spread(y.data(), KR, heads);
    for (int i = 0; i < KR; ++i)
        if (memcmp(&y[i], &yr[i], 8) || w[i] != (uint16_t)wr[i]) {
            fprintf(stderr, "group mismatch K=%d G=%d kind=%d at %d\n", K, G, kind, i);
            return 1 + i;
        }
    return 0;
}

int main(int argc, char **argv) {
    const long reps = argc > 1 ? atol(argv[1]) : 20000;
    long cases = 0;
    for (long it = 0; it < reps; ++it) {
        for (int kind = 0; kind < 7; ++kind)
            for (int warm = 0; warm < 2; ++warm)
                for (int update = 0; update < 2; ++update) {
                    const int K32 = 1 + (int)(rnd() % 32), K64 = 1 + (int)(rnd() % 64);
                    if (run_case<uint32_t>(K32, kind, warm, update)) return 1;
                    if (run_case<uint64_t>(K64, kind, warm, update)) return 1;
                    cases += 2;
                }
        for (int kind = 0; kind < 7; ++kind) {
            const int K = 1 + (int)(rnd() % 16), G32 = 1 + (int)(rnd() % (32 / K)), G64 = 1 + (int)(rnd() % (64 / K));
            if (run_group<uint32_t>(K, G32, kind) || run_group<uint64_t>(K, G64, kind)) return 1;
            cases += 2;
        }
    }
    for (int once = 0; once < 1; ++once) {
    // full-width blocks
    for (int kind = 0; kind < 7; ++kind) {
        if (run_case<uint32_t>(32, kind, 0, 1) || run_case<uint64_t>(64, kind, 0, 1) || run_case<uint32_t>(32, kind, 1, 1) || run_case<uint64_t>(64, kind, 1, 0)) return 1;
        cases += 4;
    }
    }
    printf("ok %ld cases\n", cases);
    return 0;
}
