// BENCH / TEST TOOLING, not part of the product library: builds into agatha_b200/lib/libagatha_synth.so (agatha_b200/build.py).
// Deterministic synthetic read/reference pairs for the workloads of BASELINE.md section 2.3 (C1..C4).
// Every pair is generated from hash(seed, pair index) alone, so any shard of any size reproduces the same pairs
// (the multi-GPU bench generates each rank's shard independently). No GPU needed.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "agatha_synth.h"

namespace {

struct Rng {   // xoshiro256**
    uint64_t s[4];
    static uint64_t splitmix(uint64_t& x) { uint64_t z = (x += 0x9e3779b97f4a7c15ull); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); }
    Rng(uint64_t seed, uint64_t stream) { uint64_t x = seed * 0x2545f4914f6cdd1dull + stream * 0x9e3779b97f4a7c15ull + 0x1234567ull; for (auto& v : s) v = splitmix(x); }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() { const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17; s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45); return r; }
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    double normal() { double u1 = uniform(), u2 = uniform(); if (u1 < 1e-300) u1 = 1e-300; return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2); }
};

struct Profile { double sub, ins, del; double tail_prob; bool cut_tail; };

const char ACGT[4] = {'A', 'C', 'G', 'T'};

struct PairSpec { uint32_t tlen; double sub, ins, del; uint32_t tail_len; int64_t cut; };

// Draws the per-pair shape. All randomness for the shape comes first so that both passes agree.
bool draw_spec(int profile, Rng& r, PairSpec& sp)
{
    sp.tail_len = 0; sp.cut = -1;
    switch (profile) {
        case 1: {   // C1: stand-in for the bundled dataset: 1-8 kb, 5-15 % error, 1/3 with a random tail
            sp.tlen = 1000 + (uint32_t)(r.uniform() * 7001.0);
            const double e = 0.05 + 0.10 * r.uniform();
            sp.sub = 0.4 * e; sp.ins = 0.3 * e; sp.del = 0.3 * e;
            if (r.uniform() < 1.0 / 3.0) sp.tail_len = 200 + (uint32_t)(r.uniform() * 1800.0);
            return true;
        }
        case 2: {   // C2: ONT-like, lognormal mean 10 kb sigma 0.35, clip [1,30] kb, 4/3/3 %
            const double sigma = 0.35, mu = std::log(10000.0) - 0.5 * sigma * sigma;
            double l = std::exp(mu + sigma * r.normal());
            sp.tlen = (uint32_t)std::min(30000.0, std::max(1000.0, l));
            sp.sub = 0.04; sp.ins = 0.03; sp.del = 0.03;
            return true;
        }
        case 3: {   // C3: HiFi-like, N(15 kb, 2 kb) clip [5,25] kb, 0.4/0.3/0.3 %
            double l = 15000.0 + 2000.0 * r.normal();
            sp.tlen = (uint32_t)std::min(25000.0, std::max(5000.0, l));
            sp.sub = 0.004; sp.ins = 0.003; sp.del = 0.003;
            return true;
        }
        case 4: {   // C4: heavy tail Pareto(alpha 1.2, min 1 kb) clip 100 kb, 10 % error, half diverge after a random cut
            double u = r.uniform(); if (u < 1e-12) u = 1e-12;
            double l = 1000.0 / std::pow(u, 1.0 / 1.2);
            sp.tlen = (uint32_t)std::min(100000.0, l);
            sp.sub = 0.04; sp.ins = 0.03; sp.del = 0.03;
            if (r.uniform() < 0.5) sp.cut = (int64_t)(r.uniform() * sp.tlen);
            return true;
        }
    }
    return false;
}

// Generates one pair. q/t may be NULL (size pass). Returns lengths.
void gen_pair(int profile, uint64_t seed, uint64_t index, uint8_t* q, uint8_t* t, uint32_t& qlen, uint32_t& tlen)
{
    Rng r(seed, index);
    PairSpec sp;
    draw_spec(profile, r, sp);
    tlen = sp.tlen;
    const uint32_t tsub = (uint32_t)(sp.sub * 4294967296.0), tdel = (uint32_t)(sp.del * 4294967296.0), tins = (uint32_t)(sp.ins * 4294967296.0);
    uint32_t ql = 0;
    for (uint32_t i = 0; i < sp.tlen; i++) {
        const uint64_t x = r.next();
        const unsigned base = (unsigned)(x >> 62);
        if (t) t[i] = (uint8_t)ACGT[base];
        const uint32_t u1 = (uint32_t)x, u2 = (uint32_t)(x >> 30);
        if (sp.cut >= 0 && (int64_t)i >= sp.cut) {            // diverged: unrelated random sequence from the cut on
            if (q) q[ql] = (uint8_t)ACGT[(x >> 20) & 3];
            ql++;
            continue;
        }
        if (u1 >= tdel) {                                     // not deleted
            unsigned b = base;
            if (u1 - tdel < tsub) b = (base + 1 + (unsigned)((x >> 40) % 3)) & 3;   // substitution: one of the 3 other bases
            if (q) q[ql] = (uint8_t)ACGT[b];
            ql++;
        }
        if (u2 < tins) { if (q) q[ql] = (uint8_t)ACGT[(x >> 44) & 3]; ql++; }
    }
    for (uint32_t i = 0; i < sp.tail_len; i++) { const uint64_t x = r.next(); if (q) q[ql] = (uint8_t)ACGT[x >> 62]; ql++; }
    if (ql == 0) { if (q) q[0] = 'A'; ql = 1; }
    qlen = ql;
}

}  // namespace

extern "C" int agatha_synth_pairs(int32_t profile, uint64_t seed, uint64_t first_pair, uint64_t n_pairs,
                                  uint32_t* query_lens, uint32_t* target_lens, uint64_t* query_offsets, uint64_t* target_offsets,
                                  uint8_t* query_bases, uint64_t query_capacity, uint8_t* target_bases, uint64_t target_capacity,
                                  int32_t n_threads)
{

    if (profile < 1 || profile > 4) return -1;   // profile must be 1..4
    if (!query_lens || !target_lens) return -1;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#else
    n_threads = 1;
#endif
    if (!query_bases || !target_bases) {      // size pass
#pragma omp parallel for schedule(dynamic, 64) num_threads(n_threads)
        for (int64_t i = 0; i < (int64_t)n_pairs; i++) gen_pair(profile, seed, first_pair + (uint64_t)i, nullptr, nullptr, query_lens[i], target_lens[i]);
        if (query_offsets && target_offsets) {
            uint64_t qo = 0, to = 0;
            for (uint64_t i = 0; i < n_pairs; i++) { query_offsets[i] = qo; target_offsets[i] = to; qo += query_lens[i]; to += target_lens[i]; }
        }
        return 0;
    }
    if (!query_offsets || !target_offsets) return -1;
    if (n_pairs) {
        const uint64_t last = n_pairs - 1;
        if (query_offsets[last] + query_lens[last] > query_capacity || target_offsets[last] + target_lens[last] > target_capacity)
            return -2;   // base buffers too small
    }
#pragma omp parallel for schedule(dynamic, 64) num_threads(n_threads)
    for (int64_t i = 0; i < (int64_t)n_pairs; i++) {
        uint32_t ql, tl;
        gen_pair(profile, seed, first_pair + (uint64_t)i, query_bases + query_offsets[i], target_bases + target_offsets[i], ql, tl);
    }
    return 0;
}
