// Host side of the data path: one graph epoch of the guidance-graph negative sampler
// (scglue/models/data.py:794-822, GraphDataset.propose_shuffle), draw for draw on the legacy MT19937 stream of
// numpy.random.RandomState(seed), as one native call.
//
// Why native: a graph epoch of the PBMC shape (1.15 M sampled edges) feeds 32 training steps = 21 ms of GPU
// time.  The numpy restatement (glue_b200/models/data.py) takes 0.11 s per graph epoch alone, and its look-ahead
// threads do not scale: between its ~150 array calls a thread needs the interpreter lock, which the training loop
// holds.  This function runs the whole epoch without the interpreter (ctypes drops the lock for the call), so
// look-ahead threads scale with the host cores.
//
// No device code here; it lives in the library because the sampler is part of the path's data side and the
// boundary is this C ABI.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <new>
#include <vector>

#include "glue_b200.h"

namespace {

// MT19937 as numpy's legacy RandomState drives it (init_genrand seeding for an integer seed, 32-bit outputs;
// doubles from two outputs: 27 + 26 bits).
struct Mt19937 {
    static constexpr int N = 624, M = 397;
    uint32_t key[N];
    int pos;

    explicit Mt19937(uint32_t seed) {
        for (int i = 0; i < N; ++i) {
            key[i] = seed;
            seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)(i + 1);
        }
        pos = N;
    }

    void refill() {
        auto twist = [](uint32_t u, uint32_t v) {
            uint32_t y = (u & 0x80000000u) | (v & 0x7fffffffu);
            return (y >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
        };
        int i = 0;
        for (; i < N - M; ++i) key[i] = key[i + M] ^ twist(key[i], key[i + 1]);
        for (; i < N - 1; ++i) key[i] = key[i + (M - N)] ^ twist(key[i], key[i + 1]);
        key[N - 1] = key[M - 1] ^ twist(key[N - 1], key[0]);
        pos = 0;
    }

    inline uint32_t next() {
        if (pos == N) refill();
        uint32_t y = key[pos++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }

    inline double next_double() {  // RandomState.random_sample
        const uint32_t a = next() >> 5, b = next() >> 6;
        return (a * 67108864.0 + b) / 9007199254740992.0;
    }

    inline uint64_t interval(uint64_t max) {  // legacy shuffle's bounded draw (mask + rejection), max < 2^32
        uint64_t mask = max;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4;
        mask |= mask >> 8; mask |= mask >> 16; mask |= mask >> 32;
        uint64_t value;
        while ((value = (next() & mask)) > max) {}
        return value;
    }
};

constexpr int BLOCK = 64;  // draws resolved together: their table lines are prefetched before the first is read

// RandomState.choice(n, size, replace=True, p=p): per draw a uniform, then searchsorted(cdf, u, "right").
// lut[b] = number of cdf entries <= b / 2^bits, so the answer for u in bucket b lies in [lut[b], lut[b + 1]].
struct CdfTable {
    const double *cdf;
    const int32_t *lut;
    double scale;

    template <class Sink>
    void draw(Mt19937 &rng, int64_t n, Sink &&sink) const {  // sink(k, answer) for k = 0 .. n - 1, in order
        double u[BLOCK];
        int64_t bucket[BLOCK];
        for (int64_t base = 0; base < n; base += BLOCK) {
            const int m = (int)std::min<int64_t>(BLOCK, n - base);
            for (int k = 0; k < m; ++k) {
                u[k] = rng.next_double();
                bucket[k] = (int64_t)(u[k] * scale);
                __builtin_prefetch(lut + bucket[k]);
            }
            for (int k = 0; k < m; ++k) {
                const int64_t lo = lut[bucket[k]], hi = lut[bucket[k] + 1];
                sink(base + k, lo == hi ? lo : (int64_t)(std::upper_bound(cdf + lo, cdf + hi, u[k]) - cdf));
            }
        }
    }
};

// The real edges as keys (src * V + dst) * 2 + (sign > 0): a bit filter without false negatives in front of the
// sorted keys.
struct EdgeSet {
    const int64_t *keys;
    int64_t n_keys;
    const uint8_t *filter_bits;
    int shift;
    int64_t n_vertices;

    inline int64_t key(int64_t src, int64_t dst, float sgn) const {
        return (src * n_vertices + dst) * 2 + (sgn > 0.0f ? 1 : 0);
    }
    inline uint64_t slot(int64_t key) const { return ((uint64_t)key * 0x9E3779B97F4A7C15ull) >> shift; }
    inline void prefetch(uint64_t slot) const { __builtin_prefetch(filter_bits + (slot >> 3)); }
    inline bool contains(int64_t key, uint64_t slot) const {
        if (!((filter_bits[slot >> 3] >> (slot & 7)) & 1)) return false;
        return std::binary_search(keys, keys + n_keys, key);
    }
};

}  // namespace

extern "C" int glue_sample_graph_epoch(uint32_t seed, const int64_t *esrc, const int64_t *edst, const float *esgn,
                                       int64_t n_edges, const double *ecdf, const int32_t *elut, int32_t elut_bits,
                                       const double *vcdf, const int32_t *vlut, int32_t vlut_bits,
                                       int64_t n_vertices, const int64_t *edge_keys, int64_t n_keys,
                                       const uint8_t *key_filter, int32_t filter_bits, int64_t n_pos,
                                       int32_t neg_samples, int64_t *out_src, int64_t *out_dst, float *out_wt,
                                       float *out_sgn) {
    if (!esrc || !edst || !esgn || !ecdf || !elut || !vcdf || !vlut || !edge_keys || !key_filter || !out_src ||
        !out_dst || !out_wt || !out_sgn)
        return GLUE_E_BADARG;
    if (n_edges <= 0 || n_vertices <= 0 || n_pos < 0 || neg_samples < 0 || n_keys < 0 || elut_bits < 1 ||
        elut_bits > 30 || vlut_bits < 1 || vlut_bits > 30 || filter_bits < 3 || filter_bits > 40)
        return GLUE_E_BADARG;
    if (n_edges > 0x7fffffffll || n_vertices > 0x7fffffffll) return GLUE_E_UNSUPPORTED;  // (32-bit tables)
    const int64_t n_neg = n_pos * neg_samples, total = n_pos + n_neg;
    if (total > 0xffffffffll) return GLUE_E_UNSUPPORTED;  // (the 64-bit branch of the bounded draw is not restated)
    try {
        Mt19937 rng(seed);
        const CdfTable edges{ecdf, elut, (double)(1ll << elut_bits)};
        const CdfTable vertices{vcdf, vlut, (double)(1ll << vlut_bits)};
        const EdgeSet real{edge_keys, n_keys, key_filter, 64 - filter_bits, n_vertices};

        // positives: round(sum(ewt)) edges with replacement; their triplets are copied out once (the ten
        // corrupted copies and the final gather then read 12 contiguous bytes per positive instead of chasing
        // the edge index into three arrays)
        struct Positive {
            int32_t src, dst;
            float sgn;
        };
        std::vector<Positive> positive((size_t)n_pos);
        std::vector<int64_t> key_base((size_t)n_pos);  // key of (src, 0, sign); a target adds 2 * dst
        std::vector<int32_t> n_dst((size_t)n_neg);
        std::vector<int64_t> clash;
        edges.draw(rng, n_pos, [&](int64_t i, int64_t e) {
            positive[i] = Positive{(int32_t)esrc[e], (int32_t)edst[e], esgn[e]};
            key_base[i] = real.key(esrc[e], 0, esgn[e]);
        });
        // negatives: every positive neg_samples times (tiled), target redrawn from the vertex distribution;
        // all first draws come before any redraw, redraws go over the clashing slots in increasing order
        vertices.draw(rng, n_neg, [&](int64_t t, int64_t v) { n_dst[t] = (int32_t)v; });
        {
            int64_t key[BLOCK];
            uint64_t slot[BLOCK];
            for (int32_t rep = 0; rep < neg_samples; ++rep) {  // slot t = rep * n_pos + i corrupts positive i
                const int32_t *dst = n_dst.data() + (int64_t)rep * n_pos;
                for (int64_t base = 0; base < n_pos; base += BLOCK) {
                    const int m = (int)std::min<int64_t>(BLOCK, n_pos - base);
                    for (int k = 0; k < m; ++k) {
                        key[k] = key_base[base + k] + 2 * (int64_t)dst[base + k];
                        slot[k] = real.slot(key[k]);
                        real.prefetch(slot[k]);
                    }
                    for (int k = 0; k < m; ++k)
                        if (real.contains(key[k], slot[k])) clash.push_back((int64_t)rep * n_pos + base + k);
                }
            }
        }
        while (!clash.empty()) {  // NOTE: does not terminate on (near-)complete graphs, as in the reference
            vertices.draw(rng, (int64_t)clash.size(), [&](int64_t k, int64_t v) { n_dst[clash[k]] = (int32_t)v; });
            size_t kept = 0;
            for (int64_t t : clash) {
                const int64_t key = key_base[t % n_pos] + 2 * (int64_t)n_dst[t];
                if (real.contains(key, real.slot(key))) clash[kept++] = t;
            }
            clash.resize(kept);
        }
        // RandomState.permutation(total): arange, then the legacy in-place shuffle from the back
        std::vector<uint32_t> perm((size_t)total);
        for (int64_t i = 0; i < total; ++i) perm[i] = (uint32_t)i;
        {
            uint32_t partner[BLOCK];
            for (int64_t top = total - 1; top >= 1; top -= BLOCK) {
                const int m = (int)std::min<int64_t>(BLOCK, top);  // positions top, top - 1, ..., top - m + 1 >= 1
                for (int k = 0; k < m; ++k) {
                    partner[k] = (uint32_t)rng.interval((uint64_t)(top - k));
                    __builtin_prefetch(perm.data() + partner[k], 1);
                }
                for (int k = 0; k < m; ++k) std::swap(perm[top - k], perm[partner[k]]);
            }
        }
        constexpr int AHEAD = 16;
        const uint32_t n_pos32 = (uint32_t)n_pos;  // (total < 2^32: 32-bit remainders)
        for (int64_t i = 0; i < total; ++i) {
            if (i + AHEAD < total) {
                const uint32_t ja = perm[i + AHEAD];
                if (ja >= n_pos32) __builtin_prefetch(n_dst.data() + (ja - n_pos32));
                __builtin_prefetch(positive.data() + (ja < n_pos32 ? ja : (ja - n_pos32) % n_pos32));
            }
            const uint32_t j = perm[i];
            const bool pos = j < n_pos32;
            const Positive &p = positive[pos ? j : (j - n_pos32) % n_pos32];
            out_src[i] = p.src;
            out_dst[i] = pos ? p.dst : n_dst[j - n_pos32];
            out_wt[i] = pos ? 1.0f : 0.0f;
            out_sgn[i] = p.sgn;
        }
    } catch (const std::bad_alloc &) {
        return GLUE_E_BADARG;
    }
    return 0;
}
