// generators.cpp (librxgen.so, host only) -- benchmark-input restatements of the reference's seeded random generators.
//
// Generators are not part of the accelerated path (OUT OF SCOPE as a product);
// BASELINE.json's configs name three of them, so their ALGORITHMS are restated here:
//   gnp : Batagelj-Brandes geometric skipping, rustworkx-core/src/generators/random_graph.rs:297-342
//   gnm : rejection sampling with an edge set,  random_graph.rs:446-465
//   BA  : repeated-nodes preferential attachment from a star of m nodes, random_graph.rs:785-822
// The reference draws from rand_pcg::Pcg64 seeded by rand's seed_from_u64 (un-vendored crates), so
// the exact edge sets cannot be reproduced here: graph identity is UNPINNED, distributions match.
// Parity is always measured GPU-vs-oracle on the same explicit edge list.
#include <math.h>
#include <stdint.h>

#include <vector>

#include "../../include/rxgen.h"

namespace {

// PCG XSL RR 128/64 (the generator family behind rand_pcg::Pcg64), seeded by SplitMix64 expansion.
struct Pcg64 {
    __uint128_t state, inc;
    static uint64_t splitmix(uint64_t &x) {
        uint64_t z = (x += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    explicit Pcg64(uint64_t seed) {
        uint64_t x = seed;
        const uint64_t a = splitmix(x), b = splitmix(x), c = splitmix(x), d = splitmix(x);
        const __uint128_t init = ((__uint128_t)a << 64) | b;
        inc = ((((__uint128_t)c << 64) | d) << 1) | 1;
        state = 0;
        step();
        state += init;
        step();
    }
    void step() {
        const __uint128_t mult =
            ((__uint128_t)0x2360ED051FC65DA4ull << 64) | 0x4385DF649FCCF645ull;
        state = state * mult + inc;
    }
    uint64_t next() {
        step();
        const uint64_t hi = (uint64_t)(state >> 64), lo = (uint64_t)state;
        const unsigned rot = (unsigned)(hi >> 58);
        const uint64_t x = hi ^ lo;
        return (x >> rot) | (x << ((64 - rot) & 63));
    }
    double uniform01() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    // unbiased integer in [0, n)
    uint64_t below(uint64_t n) {
        __uint128_t m = (__uint128_t)next() * n;
        uint64_t l = (uint64_t)m;
        if (l < n) {
            const uint64_t t = (0 - n) % n;
            while (l < t) {
                m = (__uint128_t)next() * n;
                l = (uint64_t)m;
            }
        }
        return (uint64_t)(m >> 64);
    }
};

struct EdgeSink {
    uint32_t *src, *dst;
    uint64_t cap, n = 0;
    bool overflow = false;
    void push(uint32_t a, uint32_t b) {
        if (n < cap) {
            src[n] = a;
            dst[n] = b;
        } else {
            overflow = true;
        }
        n++;
    }
};

// open-addressing set of 64-bit keys (key + 1 stored; 0 = empty)
struct U64Set {
    std::vector<uint64_t> tab;
    uint64_t mask;
    explicit U64Set(uint64_t expected) {
        uint64_t c = 16;
        while (c < expected * 2) c <<= 1;
        tab.assign(c, 0);
        mask = c - 1;
    }
    bool insert(uint64_t key) {
        uint64_t h = key * 0x9E3779B97F4A7C15ull;
        h ^= h >> 29;
        for (uint64_t i = h & mask;; i = (i + 1) & mask) {
            if (tab[i] == 0) {
                tab[i] = key + 1;
                return true;
            }
            if (tab[i] == key + 1) return false;
        }
    }
};

}  // namespace

extern "C" int64_t rxc_gen_gnp(uint32_t n, double p, uint64_t seed, int directed, uint32_t *src,
                               uint32_t *dst, uint64_t cap) {
    if (n == 0 || !(p >= 0.0 && p <= 1.0)) return RXG_ERR_INVALID;
    EdgeSink out{src, dst, cap};
    Pcg64 rng(seed);
    if (p > 0.0) {
        if (fabs(p - 1.0) < 2.220446049250313e-16) {
            for (uint32_t u = 0; u < n; u++)
                for (uint32_t v = directed ? 0 : u + 1; v < n; v++)
                    if (!directed || u != v) out.push(u, v);
        } else {
            const int64_t nn = (int64_t)n;
            const double lp = log(1.0 - p);
            if (directed) {  // inward edges to v
                int64_t v = 0, w = -1;
                while (v < nn) {
                    const double lr = log(1.0 - rng.uniform01());
                    w = w + 1 + (int64_t)(lr / lp);
                    while (w >= v && v < nn) {
                        w -= v;
                        v += 1;
                    }
                    if (v == w) {  // skip self-loops
                        w -= v;
                        v += 1;
                    }
                    if (v < nn) out.push((uint32_t)w, (uint32_t)v);
                }
            }
            int64_t v = 1, w = -1;  // outward edges from v
            while (v < nn) {
                const double lr = log(1.0 - rng.uniform01());
                w = w + 1 + (int64_t)(lr / lp);
                while (w >= v && v < nn) {
                    w -= v;
                    v += 1;
                }
                if (v < nn) out.push((uint32_t)v, (uint32_t)w);
            }
        }
    }
    return out.overflow ? (int64_t)RXG_ERR_INVALID : (int64_t)out.n;
}

extern "C" int64_t rxc_gen_gnm(uint32_t n, uint64_t m, uint64_t seed, int directed, uint32_t *src,
                               uint32_t *dst, uint64_t cap) {
    if (n == 0) return RXG_ERR_INVALID;
    EdgeSink out{src, dst, cap};
    Pcg64 rng(seed);
    const uint64_t div_by = directed ? 1 : 2;
    const uint64_t max_edges = (uint64_t)n * (n - 1) / div_by;
    if (m >= max_edges) {
        for (uint32_t u = 0; u < n; u++)
            for (uint32_t v = directed ? 0 : u + 1; v < n; v++)
                if (!directed || u != v) out.push(u, v);
    } else {
        U64Set existing(m);
        uint64_t created = 0;
        while (created < m) {
            const uint32_t u = (uint32_t)rng.below(n), v = (uint32_t)rng.below(n);
            if (u == v) continue;
            uint32_t a = u, b = v;
            if (!directed && b < a) {
                a = v;
                b = u;
            }
            if (existing.insert(((uint64_t)a << 32) | b)) {
                out.push(u, v);
                created++;
            }
        }
    }
    return out.overflow ? (int64_t)RXG_ERR_INVALID : (int64_t)out.n;
}

extern "C" int64_t rxc_gen_barabasi_albert(uint32_t n, uint32_t m, uint64_t seed, uint32_t *src,
                                           uint32_t *dst, uint64_t cap) {
    if (m < 1 || m >= n) return RXG_ERR_INVALID;
    EdgeSink out{src, dst, cap};
    Pcg64 rng(seed);
    // star_graph(m): hub 0, leaves 1..m-1 (star_graph.rs:98-108)
    for (uint32_t a = 1; a < m; a++) out.push(0, a);
    // repeated_nodes: the reference counts edges_directed(Outgoing) chained with
    // edges_directed(Incoming), which on an undirected graph lists every incident edge twice
    std::vector<uint32_t> repeated;
    repeated.reserve((size_t)4 * (m - 1) + (size_t)2 * m * (n - m));
    for (uint32_t x = 0; x < m; x++) {
        const uint32_t degree = (x == 0) ? (m - 1) : 1;
        for (uint32_t k = 0; k < 2 * degree; k++) repeated.push_back(x);
    }
    if (repeated.empty()) repeated.push_back(0);  // m == 1: a single isolated seed node
    std::vector<uint32_t> targets;
    std::vector<uint32_t> stamp(n, 0xFFFFFFFFu);
    for (uint32_t source = m; source < n; source++) {
        targets.clear();
        while (targets.size() < m) {
            const uint32_t t = repeated[rng.below(repeated.size())];
            if (stamp[t] != source) {  // IndexSet: distinct, insertion ordered
                stamp[t] = source;
                targets.push_back(t);
            }
        }
        for (uint32_t t : targets) out.push(source, t);
        repeated.insert(repeated.end(), targets.begin(), targets.end());
        repeated.insert(repeated.end(), m, source);
    }
    return out.overflow ? (int64_t)RXG_ERR_INVALID : (int64_t)out.n;
}
