// ORACLE — TEST INFRASTRUCTURE ONLY (see model.hpp header).
//
// Device-schedule restatement.  The ARITHMETIC of one update is the reference's
// (GibbsSimple::conditional = gibbs-simple.go:171-258, UniformSampler::weighted_sample =
// sampler.go:107-123, marginal count = chain.go:231-236); only the two things the device
// build changes on purpose are substituted:
//   * the variable ORDER of a sweep is supplied by the caller (the device's colour-sorted
//     schedule) instead of the reference's random scan (sampler.go:135-174);
//   * the uniform of each draw comes from the device's counter-based Philox stream
//     U(seed; chain, sweep, var) instead of the shared MT19937 channel (rand/rand.go).
// Updating the variables of one colour sequentially gives the same state as updating them
// concurrently iff the colouring is proper, so a bit-exact match of this restatement with
// the device trajectory validates gathers, index math, floor, inverse CDF, Philox,
// counting AND the colouring.
#pragma once
#include <cmath>

#include "sampler.hpp"

namespace oracle {

// Stream tags — MUST match grample_b200/csrc/philox.cuh
enum : uint32_t { kTagDraw24 = 1, kTagDraw53 = 2, kTagInit = 3, kTagScan = 4, kTagDraw16Hi = 6, kTagDraw16Lo = 7,
                  kTagPlaneA = 8, kTagPlaneB = 9, kTagTie24 = 10 };
constexpr int kBitsPlanes = 33;  // `bits` value selecting the 32-bit draw of the bit-sliced table sweep (GB_TABLE_BITS)

inline double philox_uniform(uint64_t seed, uint32_t chain, uint32_t sweep, uint32_t var, int bits) {
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t out[4];
    if (bits == 53) {
        uint32_t ctr[4] = {var, sweep, chain >> 1, kTagDraw53};
        Philox4x32::gen(ctr, key, out);
        int a = 2 * (int)(chain & 1);
        uint64_t x = (((uint64_t)out[a] << 32) | out[a + 1]) >> 11;
        return (double)x * (1.0 / 9007199254740992.0);
    }
    if (bits == kBitsPlanes) {
        // GB_TABLE_BITS: the draws of the 32 chains of one state word are bit-sliced over Philox output words
        // (philox.cuh): bit b of the top byte = bit p of plane word b; the low 24 bits come from a call of their own
        const uint32_t w = chain >> 5, p = chain & 31;
        uint32_t a[4], b[4], t[4];
        uint32_t ca[4] = {var, sweep, w, kTagPlaneA}, cb[4] = {var, sweep, w, kTagPlaneB}, ct[4] = {var, sweep, w, kTagTie24 | ((p >> 2) << 8)};
        Philox4x32::gen(ca, key, a);
        Philox4x32::gen(cb, key, b);
        Philox4x32::gen(ct, key, t);
        uint32_t hi8 = 0;
        for (int i = 0; i < 4; i++) {
            hi8 |= ((a[i] >> p) & 1u) << (7 - i);
            hi8 |= ((b[i] >> p) & 1u) << (3 - i);
        }
        const uint32_t word = (hi8 << 24) | (t[p & 3] >> 8);
        return (double)word * (1.0 / 4294967296.0);
    }
    // 32-bit draw of the table-mode sweep: hi and lo halves come from two calls shared by 8 chains
    uint32_t hi[4], lo[4];
    uint32_t ch[4] = {var, sweep, chain >> 3, kTagDraw16Hi}, cl[4] = {var, sweep, chain >> 3, kTagDraw16Lo};
    Philox4x32::gen(ch, key, hi);
    Philox4x32::gen(cl, key, lo);
    const uint32_t i = chain & 7, sh = 16 * (i & 1);
    const uint32_t word = (((hi[i >> 1] >> sh) & 0xffffu) << 16) | ((lo[i >> 1] >> sh) & 0xffffu);
    (void)out;
    return (double)word * (1.0 / 4294967296.0);
}

inline int philox_init_value(uint64_t seed, uint32_t chain, uint32_t var, int card) {
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t ctr[4] = {var, 0u, chain >> 2, kTagInit};
    uint32_t out[4];
    Philox4x32::gen(ctr, key, out);
    return (int)(((uint64_t)out[chain & 3] * (uint64_t)card) >> 32);
}

struct FixedUniform : Generator {
    double next = 0.0;
    double float64() override { return next; }
    int64_t int63() override { throw Error("FixedUniform: only Float64 draws are defined"); }
};

// Runs `n_sweeps` systematic sweeps over `order` for one chain.  `state` is updated in
// place; counts[off[v] + k] is incremented for every recorded draw when `record`.
// var_bits (optional, per variable): draw width of that variable — the device's hybrid mode draws 32 bits
// for tabulated variables and 53 bits for the others; nullptr = `bits` for every variable.
inline void sweep_chain(GibbsSimple& gs, const std::vector<int>& order, uint64_t seed, uint32_t chain,
                        uint32_t sweep0, uint32_t n_sweeps, int bits, bool record, int* state,
                        const std::vector<int>& count_off, double* counts, const int* var_bits = nullptr,
                        bool rao_blackwell = false) {
    FixedUniform fu;
    UniformSampler us(&fu, 1);
    std::vector<double> w;
    for (uint32_t s = 0; s < n_sweeps; s++) {
        for (int v : order) {
            gs.conditional(v, state, w);
            fu.next = philox_uniform(seed, chain, sweep0 + s, (uint32_t)v, var_bits ? var_bits[v] : bits);
            int k;
            try {
                k = us.weighted_sample((int64_t)w.size(), w.data(), w.size());
            } catch (const Error&) {
                k = (int)w.size() - 1;  // device convention for the (measure-zero) fall-through
            }
            state[v] = k;
            if (record && !gs.pgm->vars[v].collapsed) {
                if (rao_blackwell) {  // the device's GB_CHAINS_RAO_BLACKWELL bins: round(p_k * 2^24) for every value
                    double tot = 0.0;
                    for (double e : w) tot += e;
                    const double scale = 16777216.0 / tot;
                    for (size_t i = 0; i < w.size(); i++) counts[count_off[v] + i] += std::floor(w[i] * scale + 0.5);
                } else {
                    counts[count_off[v] + k] += 1.0;
                }
            }
        }
    }
}

// Random-scan parity mode of the device (k_random_scan): the reference's schedule with the
// device's Philox stream — one call per (chain, step): word 0 picks the variable by
// multiply-shift, words 2..3 form the 53-bit uniform of the value draw.
inline void scan_chain(GibbsSimple& gs, const std::vector<int>& order, uint64_t seed, uint32_t chain, uint64_t step0,
                       int64_t n_steps, bool record, int* state, const std::vector<int>& count_off, double* counts) {
    FixedUniform fu;
    UniformSampler us(&fu, 1);
    std::vector<double> w;
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (int64_t s = 0; s < n_steps; s++) {
        const uint64_t step = step0 + (uint64_t)s;
        uint32_t ctr[4] = {(uint32_t)step, (uint32_t)(step >> 32), chain, kTagScan}, out[4];
        Philox4x32::gen(ctr, key, out);
        const int v = order[(size_t)(((uint64_t)out[0] * (uint64_t)order.size()) >> 32)];
        gs.conditional(v, state, w);
        const uint64_t x = (((uint64_t)out[2] << 32) | out[3]) >> 11;
        fu.next = (double)x * (1.0 / 9007199254740992.0);
        int k;
        try {
            k = us.weighted_sample((int64_t)w.size(), w.data(), w.size());
        } catch (const Error&) {
            k = (int)w.size() - 1;
        }
        state[v] = k;
        if (record && !gs.pgm->vars[v].collapsed) counts[count_off[v] + k] += 1.0;
    }
}

}  // namespace oracle
