// rrng.cu — host-only: the index vectors R's default RNG stream gives clustexp (reference R/RaceID_utils.R:111-112,
// 132-133: set.seed(rseed); fpc::clusterboot(seed = rseed) -> sample(n, n, replace = TRUE) + unique(), or
// sample(n, subtuning) for bootmethod = "subset").  No GPU work here — the kernels draw no random numbers — but at
// 100k cells x 50 resamples the draws are 13 M Mersenne-Twister words, which the Python mirror of R's generator
// (raceid3_stemid2_package_b200/rrng.py, kept as the checked reference) needs a second for.
// Restates R's RNG.c (RNG_Init / FixupSeeds / MT_genrand), R_unif_index / rbits (R >= 3.6 "Rejection" sampling) and
// do_sample; pinned by rrng.py's known answers (tests/test_host_and_lib.py).
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace {
struct RMT {
    uint32_t mt[624];
    int mti;
    void seed(uint32_t s)
    {
        for (int j = 0; j < 50; ++j) s = 69069u * s + 1u; // RNG_Init: initial scrambling
        uint32_t key[625];
        for (int j = 0; j < 625; ++j) { s = 69069u * s + 1u; key[j] = s; }
        memcpy(mt, key + 1, sizeof(mt)); // i_seed[0] is mti (forced to 624 by FixupSeeds), i_seed[1..624] the state
        mti = 624;
    }
    uint32_t out[624]; // tempered outputs of the current state block
    void refill()
    {
        int kk = 0;
        for (; kk < 624 - 397; ++kk) { const uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu); mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u); }
        for (; kk < 623; ++kk) { const uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu); mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u); }
        const uint32_t y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
        mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        for (int i = 0; i < 624; ++i) { // tempering, vectorisable
            uint32_t t = mt[i];
            t ^= (t >> 11);
            t ^= (t << 7) & 0x9d2c5680u;
            t ^= (t << 15) & 0xefc60000u;
            t ^= (t >> 18);
            out[i] = t;
        }
        mti = 0;
    }
    inline uint32_t next()
    {
        if (mti >= 624) refill();
        return out[mti++];
    }
    // rbits(): floor(unif_rand() * 65536) per 16 bits; unif_rand() = fixup(y * 2^-32), whose floor(. * 65536) is y >> 16
    inline uint64_t unif_index(uint64_t dn, int bits, int words, uint64_t mask)
    {
        for (;;) {
            uint64_t v = 0;
            for (int w = 0; w < words; ++w) v = 65536u * v + (uint64_t)(next() >> 16);
            v &= mask;
            if (v < dn) return v;
        }
    }
};
inline int ceil_log2(uint64_t dn)
{
    int bits = (int)ceil(log2((double)dn)); // as R_unif_index computes it
    return bits;
}
} // namespace

// out_cat: B * n ints at most (the resamples concatenated, 0-based); out_len[B].  subset_size <= 0: "boot" resamples
// (with replacement, unique() in first-occurrence order); > 0: "subset" resamples of min(n, subset_size) cells.
// pre_draw > 0: sample(n, pre_draw) is drawn first and discarded into pre_out (nullable) — clustfun draws the `samp`
// subsample from the same stream before clusterboot re-seeds, so this is only used by callers that share a stream.
extern "C" int rid_r_bootstrap_samples(int n, int B, unsigned int seed, int subset_size, int *out_cat, int *out_len)
{
    if (n < 1 || B < 0 || !out_cat || !out_len) { rid_set_error("rid_r_bootstrap_samples: bad arguments"); return RID_ERR_ARG; }
    RMT g;
    g.seed(seed);
    std::vector<unsigned char> seen;
    std::vector<int> pop;
    size_t off = 0;
    for (int b = 0; b < B; ++b) {
        int len = 0;
        if (subset_size <= 0) {
            const int bits = ceil_log2((uint64_t)n), words = bits / 16 + 1;
            const uint64_t mask = bits >= 64 ? ~0ull : ((1ull << bits) - 1);
            seen.assign((size_t)n, 0);
            for (int i = 0; i < n; ++i) {
                const int v = n > 1 ? (int)g.unif_index((uint64_t)n, bits, words, mask) : 0;
                if (n == 1) { // bits = 0: rbits still consumes one word
                    (void)g.next();
                }
                if (!seen[v]) { seen[v] = 1; out_cat[off + len++] = v; }
            }
        } else {
            const int size = subset_size < n ? subset_size : n;
            if (size < 2) {
                // sample.int(n, size) with size < 2 draws like the replace = TRUE arm
                const int bits = ceil_log2((uint64_t)n), words = bits / 16 + 1;
                const uint64_t mask = bits >= 64 ? ~0ull : ((1ull << bits) - 1);
                for (int i = 0; i < size; ++i) out_cat[off + len++] = (int)g.unif_index((uint64_t)n, bits, words, mask);
            } else {
                pop.resize((size_t)n);
                for (int i = 0; i < n; ++i) pop[i] = i;
                int nn = n;
                for (int i = 0; i < size; ++i) {
                    const int bits = ceil_log2((uint64_t)nn), words = bits / 16 + 1;
                    const uint64_t mask = bits >= 64 ? ~0ull : ((1ull << bits) - 1);
                    int j;
                    if (nn > 1) j = (int)g.unif_index((uint64_t)nn, bits, words, mask);
                    else { (void)g.next(); j = 0; }
                    out_cat[off + len++] = pop[j];
                    pop[j] = pop[--nn];
                }
            }
        }
        out_len[b] = len;
        off += (size_t)len;
    }
    return RID_OK;
}
