// sm_100a kernels of the Gibbs hot path.  Reference lines each kernel replaces (paths
// relative to the reference root):
//   K1/K2 k_sweep_colour      sampler/gibbs-simple.go:163-271 (SampleVar) + sampler/sampler.go:90-130
//                             (WeightedSample) + sampler/chain.go:221-246 (oneSample: marginal count,
//                             history) for every (variable of one colour, chain); the collapsed
//                             sampler (gibbs-collapsed.go:317-334) is the same arithmetic over a
//                             model whose factor set excludes the collapsed variables.
//   K3    k_collapse          sampler/gibbs-collapsed.go:205-260 (sum the variable out of its blanket)
//   K4    k_chain_dist        sampler/chain.go:253-290 (ChainDist) + model/error.go:81-249
//   K5    k_conditional       sampler/gibbs-simple.go:171-258 for caller-supplied states
//   K6    k_init_state        sampler/gibbs-simple.go:103-111 (uniform start, FixedVal honoured)
//   table modes (same SampleVar + WeightedSample, the float64 conditional evaluated once per neighbour configuration):
//         k_build_thresholds, k_sweep_tab (byte state, per colour), k_sweep_tab_resident (small models), and — in bits.cuh —
//         k_sweep_bits (bit-packed state, bit-sliced compare: the bench kernel)
//   merge k_merge_counts / k_merge_finalize   sampler/chain.go:96-148 (MergeChains) as integer sums + one rounding
//
// Data layout: state[var][chain] uint8, chain fastest, rows padded to a multiple of 4 chains so a
// thread reads the states of 4 consecutive chains of one neighbour with one 32-bit load and a
// warp's loads of one neighbour are contiguous (coalesced).  One Philox call serves the 4 chains.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <type_traits>

#include "philox.cuh"

namespace gb {

constexpr int kNeighborVarMaxDev = 23;  // blanket variables K3 can enumerate: model.maxTabSize = 2^23 bounds an explicit Collapse(var) to 23 binary neighbours (sampler.NeighborVarMax = 12 only limits the random pick and Adapt)
constexpr int kMaxCardDev = 64;         // GB_MAX_CARD

struct DevModel {
    int32_t n_vars, total_card, max_card;
    const int32_t* card;        // [n_vars]
    const int32_t* card_off;    // [n_vars+1]
    const int32_t* fixed;       // [n_vars]
    const int32_t* prog_off;    // [n_vars], -1 when the variable has no program
    const int32_t* prog;        // update programs (host_model.hpp::build_programs)
    const int32_t* pw_off;      // [n_vars] first pairwise record of the variable, -1 = none (a factor of arity > 2)
    const int4* pw_rec;         // {tab_off, stride of v, other variable, its stride} per factor
    const int4* pos_rec;        // [n_order] per sweep position {variable, cardinality, pw_off, number of factors}
    const double* tab64;        // log-space tables
    const float* tab32;
    int32_t n_tab;              // table entries, padded to a multiple of 4 (16-byte bulk-copy granularity)
    const int32_t* entry_var;   // [total_card] variable of each marginal entry
};

struct DevGroup {
    uint8_t* state;              // [n_vars][n_pad]
    unsigned long long* counts;  // [total_card]
    uint16_t* hist;              // [2][total_card][n_pad] or nullptr
    int32_t n_chains, n_pad;
    uint64_t first_chain;        // global id of local chain 0 (multiple of 4)
    uint32_t seed_lo, seed_hi;
    int32_t rb;                  // GB_CHAINS_RAO_BLACKWELL: counts hold fixed-point conditional probabilities (units of 2^-24)
};

struct DevTab {
    const int32_t* tp_off;  // [n_vars]
    const int32_t* tprog;   // per var: [n_nbr, thr_off, (nbr_var, stride) * n_nbr]
    uint32_t* thr;          // [n_thresholds]
    const int32_t* trec;    // [n_order][kTabRec] fixed-size record per sweep position
    int32_t n_order, n_thr;
};


__device__ __forceinline__ Philox4 philox_wide(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return Philox4{c0, c1, c2, c3};
}


template <typename Real>
__device__ __forceinline__ const Real* tables_of(const DevModel& m);
template <>
__device__ __forceinline__ const double* tables_of<double>(const DevModel& m) { return m.tab64; }
template <>
__device__ __forceinline__ const float* tables_of<float>(const DevModel& m) { return m.tab32; }

// Table reads: through the non-coherent global path (G = true) or plain loads when the tables have been
// staged in shared memory (the compiler then emits LDS).  With G = true a PREFIX of the tables may still be
// staged (StagedPrefix): a collapsed variant keeps the model's small factors first and appends the one large
// factor over the collapsed variable's blanket (up to 11^6 entries on ObjectDetection_11), so the small factors
// are served from shared memory and only the large one goes through L1/L2.  The prefix ends on a factor
// boundary, hence the test is per factor (tab_off < n).
template <typename Real>
struct StagedPrefix {
    const Real* s_tab;  // shared-memory copy of tab[0 .. n)
    int32_t n;          // entries staged (0 = none)
};
template <bool G, typename Real>
__device__ __forceinline__ Real ld_tab(const Real* p) {
    if constexpr (G) return __ldg(p);
    else return *p;
}

// gibbs-simple.go:227-258: log-weights -> floored un-normalised weights, in place.
// float64 follows the reference literally (min-shift only when min < -8, sequential floor with
// a running total).  float32 uses a max-shift instead (the floor is scale-invariant; min-shift
// would overflow float32 for wide conditionals) and a multiply in place of the divide.
// `card` may be smaller than MAXC (entries past it are ignored); callers that know card == MAXC pass it
// as a constant so that the `k < card` predicates fold away.
//
// The float64 floor test `e[k] / tot < 1e-6` (gibbs-simple.go:249) is evaluated without the division
// whenever that is provably the same decision: the running total only grows during the floor loop, by
// less than (1 + 1e-6)^64 < 1 + 1e-4, so an e[k] outside [1e-6 tot0 (1 - 1e-4), 1e-6 tot0 (1 + 1e-4)]
// (tot0 = total before the loop) compares the same way against every running total, correctly rounded
// quotient included.  Only a weight inside that band (or a non-finite total) takes the literal loop with the
// division.  The fast loop is branch-free: adding d = 0.0 to a non-negative double leaves it unchanged.
template <typename Real, int MAXC>
__device__ __forceinline__ void stabilise_exp_floor(Real (&w)[MAXC], const int card) {
    if constexpr (std::is_same<Real, double>::value) {
        double mn = w[0];
#pragma unroll
        for (int k = 1; k < MAXC; k++)
            if (k < card && w[k] < mn) mn = w[k];
        if (mn < -8.0) {
#pragma unroll
            for (int k = 0; k < MAXC; k++)
                if (k < card) w[k] = w[k] - (mn - 1.5);
        }
        double tot = 0.0;
#pragma unroll
        for (int k = 0; k < MAXC; k++)
            if (k < card) {
                const double e = exp(w[k]);
                tot += e;
                w[k] = e;
            }
        // (binary .. quaternary variables keep the literal loop: two to four divisions cost less than the band test)
        bool literal = MAXC < 8 || !(tot < 1.7e308);
        if constexpr (MAXC >= 8) {
            const double lo = tot * (1e-6 * (1.0 - 1e-4)), hi = tot * (1e-6 * (1.0 + 1e-4));
#pragma unroll
            for (int k = 0; k < MAXC; k++)
                if (k < card) literal |= (w[k] >= lo) & (w[k] <= hi);
        }
        if (!literal) {
            const double mid = tot * 1e-6;
#pragma unroll
            for (int k = 0; k < MAXC; k++)
                if (k < card) {
                    const double d = w[k] < mid ? tot * 1e-6 : 0.0;
                    tot += d;
                    w[k] += d;
                }
        } else {
#pragma unroll (MAXC < 8 ? MAXC : 1)
            for (int k = 0; k < MAXC; k++)
                if (k < card && w[k] / tot < 1e-6) {
                    const double d = tot * 1e-6;
                    tot += d;
                    w[k] += d;
                }
        }
    } else {
        float mx = w[0];
#pragma unroll
        for (int k = 1; k < MAXC; k++)
            if (k < card) mx = fmaxf(mx, w[k]);
        // exp(w - mx) as one FFMA + ex2.approx.ftz: arguments are <= 0 and a result flushed to zero is far below
        // the floor anyway.  The floor adds 1e-6 * (total before the loop) to every weight below it: the
        // reference's running total differs from that by a factor < (1 + 1e-6)^card, three orders of magnitude
        // inside the float32 tolerance of the conditional (1e-4 relative).
        const float nmx = -mx * 1.4426950408889634f;
        float tot = 0.f;
#pragma unroll
        for (int k = 0; k < MAXC; k++)
            if (k < card) {
                float e;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(w[k], 1.4426950408889634f, nmx)));
                tot += e;
                w[k] = e;
            }
        const float lim = tot * 1e-6f;
#pragma unroll
        for (int k = 0; k < MAXC; k++)
            if (k < card) w[k] += w[k] < lim ? lim : 0.f;
    }
}

// sampler.go:107-123: re-sum, r = U * tot, first k with r <= w[k].  The (measure-zero)
// fall-through that the reference turns into an error selects the last value here.
// float64: the reference's sequential subtraction, branch-free (r keeps being reduced after the hit; it is
// not used any more).  float32: value = number of prefix sums the draw exceeds (same law, no dependent
// subtract-compare chain).
template <typename Real, int MAXC>
__device__ __forceinline__ int inverse_cdf(const Real (&w)[MAXC], int card, Real u) {
    Real tot = 0;
#pragma unroll
    for (int k = 0; k < MAXC; k++)
        if (k < card) tot += w[k];
    Real r = u * tot;
    if constexpr (std::is_same<Real, double>::value && MAXC < 8) {
        int sel = card - 1;
        bool done = false;
#pragma unroll
        for (int k = 0; k < MAXC; k++)
            if (k < card && !done) {
                if (r <= w[k]) {
                    sel = k;
                    done = true;
                } else {
                    r -= w[k];
                }
            }
        return sel;
    } else if constexpr (std::is_same<Real, double>::value) {
        int sel = card - 1;
        bool done = false;
#pragma unroll
        for (int k = 0; k < MAXC; k++)
            if (k < card) {
                const bool le = r <= w[k];
                sel = (le && !done) ? k : sel;
                done |= le;
                r -= w[k];
            }
        return sel;
    } else {
        int sel = 0;
        float c = 0.f;
#pragma unroll
        for (int k = 0; k < MAXC - 1; k++)
            if (k < card - 1) {
                c += w[k];
                sel += r > c ? 1 : 0;
            }
        return sel;
    }
}

// Rao-Blackwell estimator (GB_CHAINS_RAO_BLACKWELL): instead of Marginal[value] += 1 (chain.go:235)
// every recorded update adds the conditional it sampled from, p_k = e[k] / sum(e), to ALL bins of the variable —
// the same expectation with less variance, at no extra memory traffic because the weights are already in
// registers.  Bins are 64-bit fixed point in units of 2^-24 (an integer sum: independent of the order in which
// chains arrive, so the merged estimate does not depend on scheduling or on the number of GPUs).  Lanes of a
// warp that update the same variable add up their contributions with a warp reduction first.
constexpr double kRbScale = 16777216.0;
struct NoRecord {  // plain counts: nothing to do inside the update (kernels without the estimator compile to the same code as before)
    __device__ __forceinline__ NoRecord(unsigned long long*, int) {}
    template <typename Real, int M>
    __device__ __forceinline__ void operator()(const int, const Real (&)[M], const int) const {}
    __device__ __forceinline__ void flush(const int) const {}
};
template <int MAXC>
struct RbRecord {
    unsigned long long* bins;  // counts + card_off[v]; nullptr = not recording
    int nvalid;                // chains of this work item that exist (the rest are padding)
    mutable unsigned acc[MAXC];  // this work item's fixed-point conditionals, summed over its chains (<= 4 x 2^24)
    __device__ __forceinline__ RbRecord(unsigned long long* b, int n) : bins(b), nvalid(n) {
#pragma unroll
        for (int k = 0; k < MAXC; k++) acc[k] = 0;
    }
    // called by the update with the floored weights of chain ci of the work item
    template <typename Real, int M>
    __device__ __forceinline__ void operator()(const int ci, const Real (&w)[M], const int card) const {
        static_assert(M <= MAXC, "weight vector longer than the accumulator");
        if (bins == nullptr || ci >= nvalid) return;
        Real tot = 0;
#pragma unroll
        for (int k = 0; k < M; k++)
            if (k < card) tot += w[k];
        const Real scale = (Real)kRbScale / tot;
#pragma unroll
        for (int k = 0; k < M; k++)
            if (k < card) acc[k] += (unsigned)(w[k] * scale + (Real)0.5);
    }
    // once per work item: lanes of the warp that updated the same variable reduce first, one atomic per bin
    __device__ __forceinline__ void flush(const int card) const {
        if (bins == nullptr) return;
        const unsigned peers = __match_any_sync(__activemask(), reinterpret_cast<unsigned long long>(bins));
        const bool leader = (int)(threadIdx.x & 31u) == __ffs(peers) - 1;
#pragma unroll
        for (int k = 0; k < MAXC; k++)
            if (k < card) {
                const unsigned sum = __reduce_add_sync(peers, acc[k]);
                if (leader && sum) atomicAdd(bins + k, (unsigned long long)sum);
            }
    }
};

// ------------------------------------------------------------------ K1/K2
// One variable x 4 consecutive chains: gather the Markov-blanket factor rows, stabilise,
// exponentiate, floor, inverse CDF.  `row` points at the first of the 4 chains in the row of
// variable 0, `stride` is the row stride in bytes (global layout: n_pad; shared-memory-resident
// layout: chains per CTA).  CW = chains whose weight vectors are held in registers at once.
template <typename Real, int MAXC, int CW, bool GT, bool EXACT, typename Rec>
__device__ __forceinline__ void lse_update_quad_impl(const DevModel& m, const Real* __restrict__ tab, const uint8_t* row,
                                                     const uint32_t stride, const int v, const int card_rt, const uint32_t chain0,
                                                     const uint32_t sweep, const uint32_t seed_lo, const uint32_t seed_hi,
                                                     int (&x)[4], const Rec& rb) {
    const int card = EXACT ? MAXC : card_rt;  // EXACT: the cardinality is the compile-time bound, predicates fold
    const int32_t* __restrict__ prog = m.prog + __ldg(m.prog_off + v);
    const int nf = __ldg(prog);
    Real u[4];
    if constexpr (std::is_same<Real, double>::value) {
        const Philox4 a = philox4x32_10((uint32_t)v, sweep, chain0 >> 1, kTagDraw53, seed_lo, seed_hi);
        const Philox4 b = philox4x32_10((uint32_t)v, sweep, (chain0 >> 1) + 1u, kTagDraw53, seed_lo, seed_hi);
        u[0] = u53(a.x, a.y); u[1] = u53(a.z, a.w); u[2] = u53(b.x, b.y); u[3] = u53(b.z, b.w);
    } else {
        const Philox4 a = philox4x32_10((uint32_t)v, sweep, chain0 >> 2, kTagDraw24, seed_lo, seed_hi);
        u[0] = (float)(a.x >> 8) * (1.0f / 16777216.0f);
        u[1] = (float)(a.y >> 8) * (1.0f / 16777216.0f);
        u[2] = (float)(a.z >> 8) * (1.0f / 16777216.0f);
        u[3] = (float)(a.w >> 8) * (1.0f / 16777216.0f);
    }
#pragma unroll
    for (int cb = 0; cb < 4; cb += CW) {
        Real w[CW][MAXC];
#pragma unroll
        for (int ci = 0; ci < CW; ci++)
#pragma unroll
            for (int k = 0; k < MAXC; k++) w[ci][k] = 0;
        const int32_t* __restrict__ p = prog + 1;
        for (int f = 0; f < nf; f++) {
            const int tab_off = __ldg(p), sv = __ldg(p + 1), no = __ldg(p + 2);
            p += 3;
            int b[CW];
#pragma unroll
            for (int ci = 0; ci < CW; ci++) b[ci] = tab_off;
            for (int o = 0; o < no; o++) {
                const int ov = __ldg(p), os = __ldg(p + 1);
                p += 2;
                const uint32_t s4 = *reinterpret_cast<const uint32_t*>(row + (size_t)ov * stride);
#pragma unroll
                for (int ci = 0; ci < CW; ci++) b[ci] += (int)((s4 >> (8 * (cb + ci))) & 0xffu) * os;
            }
#pragma unroll
            for (int k = 0; k < MAXC; k++)
                if (k < card) {
#pragma unroll
                    for (int ci = 0; ci < CW; ci++) w[ci][k] += ld_tab<GT>(tab + b[ci] + k * sv);
                }
        }
#pragma unroll
        for (int ci = 0; ci < CW; ci++) {
            stabilise_exp_floor<Real, MAXC>(w[ci], card);
            rb(cb + ci, w[ci], card);
            x[cb + ci] = inverse_cdf<Real, MAXC>(w[ci], card, u[cb + ci]);
        }
    }
}

// Dispatch on the variable's cardinality: binary and ternary variables (the bulk of the UAI problems) run
// bodies unrolled to exactly their cardinality instead of predicated-off iterations up to MAXC.  Same
// arithmetic in the same order, so trajectories do not change.
template <typename Real, int MAXC, int CW, bool GT = true, typename Rec = NoRecord>
__device__ __forceinline__ void lse_update_quad(const DevModel& m, const Real* __restrict__ tab, const uint8_t* row,
                                                const uint32_t stride, const int v, const int card, const uint32_t chain0,
                                                const uint32_t sweep, const uint32_t seed_lo, const uint32_t seed_hi,
                                                int (&x)[4], const Rec& rb) {
    if constexpr (MAXC > 2) {
        if (card == 2) {
            lse_update_quad_impl<Real, 2, CW, GT, true, Rec>(m, tab, row, stride, v, card, chain0, sweep, seed_lo, seed_hi, x, rb);
            return;
        }
    }
    if constexpr (MAXC > 3) {
        if (card == 3) {
            lse_update_quad_impl<Real, 3, CW, GT, true, Rec>(m, tab, row, stride, v, card, chain0, sweep, seed_lo, seed_hi, x, rb);
            return;
        }
    }
    if (card == MAXC) lse_update_quad_impl<Real, MAXC, CW, GT, true, Rec>(m, tab, row, stride, v, card, chain0, sweep, seed_lo, seed_hi, x, rb);
    else lse_update_quad_impl<Real, MAXC, CW, GT, false, Rec>(m, tab, row, stride, v, card, chain0, sweep, seed_lo, seed_hi, x, rb);
}

// One variable x ONE chain (same arithmetic and draws as lse_update_quad for that chain): used by the
// resident kernel for high-cardinality models, where one thread per chain gives 4x the parallelism
// and a quarter of the registers.  `cell` points at this chain's byte in the row of variable 0.
template <typename Real, int MAXC, bool GT, bool EXACT, typename Rec>
__device__ __forceinline__ int lse_update_one_impl(const DevModel& m, const Real* __restrict__ tab, const uint8_t* cell,
                                                   const uint32_t stride, const int v, const int card_rt, const uint32_t chain,
                                                   const uint32_t sweep, const uint32_t seed_lo, const uint32_t seed_hi,
                                                   const Rec& rb, const StagedPrefix<Real>& sp) {
    const int card = EXACT ? MAXC : card_rt;
    const int32_t* __restrict__ p = m.prog + __ldg(m.prog_off + v);
    const int nf = __ldg(p++);
    Real u;
    if constexpr (std::is_same<Real, double>::value) {
        const Philox4 a = philox4x32_10((uint32_t)v, sweep, chain >> 1, kTagDraw53, seed_lo, seed_hi);
        u = (chain & 1u) ? u53(a.z, a.w) : u53(a.x, a.y);
    } else {
        const Philox4 a = philox4x32_10((uint32_t)v, sweep, chain >> 2, kTagDraw24, seed_lo, seed_hi);
        const uint32_t wsel = (chain & 2u) ? ((chain & 1u) ? a.w : a.z) : ((chain & 1u) ? a.y : a.x);
        u = (float)(wsel >> 8) * (1.0f / 16777216.0f);
    }
    Real w[MAXC];
#pragma unroll
    for (int k = 0; k < MAXC; k++) w[k] = 0;
    for (int f = 0; f < nf; f++) {
        const int tab_off = __ldg(p), sv = __ldg(p + 1), no = __ldg(p + 2);
        p += 3;
        int b = tab_off;
        for (int o = 0; o < no; o++, p += 2) b += (int)cell[(size_t)__ldg(p) * stride] * __ldg(p + 1);
        if (GT && tab_off < sp.n) {
#pragma unroll
            for (int k = 0; k < MAXC; k++)
                if (k < card) w[k] += sp.s_tab[b + k * sv];
        } else {
#pragma unroll
            for (int k = 0; k < MAXC; k++)
                if (k < card) w[k] += ld_tab<GT>(tab + b + k * sv);
        }
    }
    stabilise_exp_floor<Real, MAXC>(w, card);
    rb(0, w, card);
    return inverse_cdf<Real, MAXC>(w, card, u);
}

// Pairwise fast path of the one-thread-per-chain body (host_model.hpp::build_programs, pw_rec): one 16-byte
// record per factor; the two strides a dense pairwise table can give the updated variable (1 = last in the
// scope, MAXC = first in the scope of an equal-cardinality pair) are compile-time, so the row's loads carry
// immediate offsets.  Same factor order and arithmetic as lse_update_one_impl: identical results.
template <typename Real, int MAXC, bool GT, typename Rec>
__device__ __forceinline__ int lse_update_one_pw(const DevModel& m, const Real* __restrict__ tab, const uint8_t* cell,
                                                 const uint32_t stride, const int v, const int32_t pw, const int nf,
                                                 const uint32_t chain, const uint32_t sweep, const uint32_t seed_lo,
                                                 const uint32_t seed_hi, const Rec& rb, const StagedPrefix<Real>& sp) {
    const int4* __restrict__ rec = m.pw_rec + pw;
    Real u;
    if constexpr (std::is_same<Real, double>::value) {
        const Philox4 a = philox4x32_10((uint32_t)v, sweep, chain >> 1, kTagDraw53, seed_lo, seed_hi);
        u = (chain & 1u) ? u53(a.z, a.w) : u53(a.x, a.y);
    } else {
        const Philox4 a = philox4x32_10((uint32_t)v, sweep, chain >> 2, kTagDraw24, seed_lo, seed_hi);
        const uint32_t wsel = (chain & 2u) ? ((chain & 1u) ? a.w : a.z) : ((chain & 1u) ? a.y : a.x);
        u = (float)(wsel >> 8) * (1.0f / 16777216.0f);
    }
    Real w[MAXC];
#pragma unroll
    for (int k = 0; k < MAXC; k++) w[k] = 0;
    int4 r = __ldg(rec);
    for (int f = 0; f < nf; f++) {
        const int4 nx = __ldg(rec + min(f + 1, nf - 1));  // next record in flight while this row is summed
        const int b = r.x + (int)cell[(size_t)(uint32_t)r.z * stride] * r.w;
        if (GT && r.x < sp.n) {  // factor inside the staged prefix: shared-memory row
            const Real* __restrict__ row = sp.s_tab + b;
            if (r.y == 1) {
#pragma unroll
                for (int k = 0; k < MAXC; k++) w[k] += row[k];
            } else if (r.y == MAXC) {
#pragma unroll
                for (int k = 0; k < MAXC; k++) w[k] += row[k * MAXC];
            } else {
#pragma unroll
                for (int k = 0; k < MAXC; k++) w[k] += row[k * r.y];
            }
        } else {
            const Real* __restrict__ row = tab + b;
            if (r.y == 1) {
#pragma unroll
                for (int k = 0; k < MAXC; k++) w[k] += ld_tab<GT>(row + k);
            } else if (r.y == MAXC) {
#pragma unroll
                for (int k = 0; k < MAXC; k++) w[k] += ld_tab<GT>(row + k * MAXC);
            } else {
#pragma unroll
                for (int k = 0; k < MAXC; k++) w[k] += ld_tab<GT>(row + k * r.y);
            }
        }
        r = nx;
    }
    stabilise_exp_floor<Real, MAXC>(w, MAXC);
    rb(0, w, MAXC);
    return inverse_cdf<Real, MAXC>(w, MAXC, u);
}

// cardinality buckets of the one-thread-per-chain body: exact unrolls for 2, 3, 11 (ObjectDetection) and MAXC;
// variables with only unary and pairwise factors take the record-driven fast path
template <typename Real, int MAXC, bool GT = true, typename Rec = NoRecord>
__device__ __forceinline__ int lse_update_one(const DevModel& m, const Real* __restrict__ tab, const uint8_t* cell,
                                              const uint32_t stride, const int v, const int card, const int32_t pw, const int nf,
                                              const uint32_t chain, const uint32_t sweep, const uint32_t seed_lo,
                                              const uint32_t seed_hi, const Rec& rb, const StagedPrefix<Real>& sp) {
    if constexpr (MAXC > 2) {
        if (card == 2) return lse_update_one_impl<Real, 2, GT, true, Rec>(m, tab, cell, stride, v, card, chain, sweep, seed_lo, seed_hi, rb, sp);
    }
    if constexpr (MAXC > 3) {
        if (card == 3) return lse_update_one_impl<Real, 3, GT, true, Rec>(m, tab, cell, stride, v, card, chain, sweep, seed_lo, seed_hi, rb, sp);
    }
    if constexpr (MAXC > 11) {
        if (card == 11) {
            if (pw >= 0) return lse_update_one_pw<Real, 11, GT, Rec>(m, tab, cell, stride, v, pw, nf, chain, sweep, seed_lo, seed_hi, rb, sp);
            return lse_update_one_impl<Real, 11, GT, true, Rec>(m, tab, cell, stride, v, card, chain, sweep, seed_lo, seed_hi, rb, sp);
        }
    }
    if (card == MAXC) {
        if (pw >= 0) return lse_update_one_pw<Real, MAXC, GT, Rec>(m, tab, cell, stride, v, pw, nf, chain, sweep, seed_lo, seed_hi, rb, sp);
        return lse_update_one_impl<Real, MAXC, GT, true, Rec>(m, tab, cell, stride, v, card, chain, sweep, seed_lo, seed_hi, rb, sp);
    }
    return lse_update_one_impl<Real, MAXC, GT, false, Rec>(m, tab, cell, stride, v, card, chain, sweep, seed_lo, seed_hi, rb, sp);
}

// Hybrid mode (GB_HYBRID): a binary variable with a threshold table (<= 4096 configurations of its free
// neighbours) is updated from the table — float64 conditional evaluated once per configuration, 32-bit
// draw, same tie rule and Philox fields as the table kernels — and every other variable by the float64
// log-sum-exp path.  One variable x 4 consecutive chains; tpo = the variable's offset in DevTab::tprog.
// Rao-Blackwell bins from a threshold table (GB_CHAINS_RAO_BLACKWELL under GB_TABLE / GB_HYBRID): a threshold is the
// cumulative float64 conditional on the 32-bit grid, T_j = floor(P(value <= j) 2^32), so the bins take p_0 = T_0 / 2^32,
// p_k = (T_k - T_{k-1}) / 2^32, p_last = 1 - T_last / 2^32, each rounded to the bins' unit of 2^-24 like RbRecord rounds
// the log-sum-exp weights: floor((T + 128) / 256) = floor(p 2^24 + 1/2), the same integer the float64 value rounds to
// (binary variables: identical bins; cumulative differences of floors may be one grid step off, a unit now and then)
__device__ __forceinline__ uint32_t rb_units(const uint64_t p32) { return (uint32_t)((p32 + 128u) >> 8); }

__device__ __forceinline__ void tab_update_quad(const DevTab& t, const int32_t tpo, const uint8_t* row, const uint32_t stride,
                                                const int v, const int card, const uint32_t chain0, const uint32_t sweep,
                                                const uint32_t seed_lo, const uint32_t seed_hi, int (&x)[4],
                                                unsigned long long* __restrict__ rb_bins = nullptr, const int rb_valid = 0) {
    const int32_t* __restrict__ tp = t.tprog + tpo;
    const int nn = __ldg(tp), thr_off = __ldg(tp + 1);
    uint32_t idx[4] = {0u, 0u, 0u, 0u};
    for (int i = 0; i < nn; i++) {
        const int ov = __ldg(tp + 2 + 2 * i);
        const uint32_t os = (uint32_t)__ldg(tp + 3 + 2 * i);
        const uint32_t s4 = *reinterpret_cast<const uint32_t*>(row + (size_t)ov * stride);
#pragma unroll
        for (int ci = 0; ci < 4; ci++) idx[ci] += ((s4 >> (8 * ci)) & 0xffu) * os;
    }
    // chains 8b .. 8b+7 share a call; this quad owns fields 4h .. 4h+3 (h = second quad of the block)
    const bool second = (chain0 >> 2) & 1u;
    const Philox4 a = philox_wide((uint32_t)v, sweep, chain0 >> 3, kTagDraw16Hi, seed_lo, seed_hi);
    const uint32_t wa[2] = {second ? a.z : a.x, second ? a.w : a.y};
    if (card > 2) {
        // ternary / quaternary variable: card - 1 cumulative thresholds per configuration, full 32-bit draws
        // (value = number of thresholds the draw exceeds)
        const Philox4 b = philox_wide((uint32_t)v, sweep, chain0 >> 3, kTagDraw16Lo, seed_lo, seed_hi);
        const uint32_t wb[2] = {second ? b.z : b.x, second ? b.w : b.y};
#pragma unroll
        for (int ci = 0; ci < 4; ci++) {
            const uint32_t u = (((wa[ci >> 1] >> (16 * (ci & 1))) & 0xffffu) << 16) | ((wb[ci >> 1] >> (16 * (ci & 1))) & 0xffffu);
            const uint32_t* __restrict__ T = t.thr + thr_off + idx[ci] * (uint32_t)(card - 1);
            int val = 0;
            uint64_t prev = 0;  // cumulative probability below the current value, in units of 2^-32
            for (int j = 0; j < card - 1; j++) {
                const uint32_t Tj = __ldg(T + j);
                val += u > Tj ? 1 : 0;
                if (rb_bins && ci < rb_valid) {
                    atomicAdd(rb_bins + j, (unsigned long long)rb_units((uint64_t)Tj - prev));
                    prev = (uint64_t)Tj;
                }
            }
            if (rb_bins && ci < rb_valid) atomicAdd(rb_bins + card - 1, (unsigned long long)rb_units(4294967296ull - prev));
            x[ci] = val;
        }
        return;
    }
    uint32_t T[4], hi[4];
    bool tie = false;
#pragma unroll
    for (int ci = 0; ci < 4; ci++) {
        T[ci] = __ldg(t.thr + thr_off + idx[ci]);
        hi[ci] = (wa[ci >> 1] >> (16 * (ci & 1))) & 0xffffu;
        x[ci] = hi[ci] > (T[ci] >> 16) ? 1 : 0;
        tie |= hi[ci] == (T[ci] >> 16);
    }
    if (rb_bins) {
        unsigned long long q0 = 0;
#pragma unroll
        for (int ci = 0; ci < 4; ci++)
            if (ci < rb_valid) q0 += rb_units((uint64_t)T[ci]);
        if (rb_valid > 0) {
            atomicAdd(rb_bins, q0);
            atomicAdd(rb_bins + 1, (unsigned long long)min(rb_valid, 4) * 16777216ull - q0);
        }
    }
    if (tie) {  // draw > threshold <=> high halves equal and lo16 > (T & 0xffff)
        const Philox4 b = philox_wide((uint32_t)v, sweep, chain0 >> 3, kTagDraw16Lo, seed_lo, seed_hi);
        const uint32_t wb[2] = {second ? b.z : b.x, second ? b.w : b.y};
#pragma unroll
        for (int ci = 0; ci < 4; ci++) {
            const uint32_t lo = (wb[ci >> 1] >> (16 * (ci & 1))) & 0xffffu;
            if (hi[ci] == (T[ci] >> 16) && lo > (T[ci] & 0xffffu)) x[ci] = 1;
        }
    }
}

__device__ __forceinline__ void hist_add(uint16_t* s_hist, const DevGroup& g, const int total_card, const int half,
                                         const int entry, const int CH, const int cc, const int lchain);

// One launch = one colour of one group.  Work item = (variable of the colour, quad of 4 chains);
// consecutive threads take consecutive quads of the same variable.
template <typename Real, int MAXC, int CW, bool RB = false>  // RB: Rao-Blackwell bins instead of counts (g.rb is set)
__global__ void __launch_bounds__(256)
k_sweep_colour(const DevModel m, const DevGroup g, const int32_t* __restrict__ vars, const int32_t n_vars_c,
               const uint32_t sweep, const int record, const int hist_half, const DevTab t, const int hybrid) {
    using Rec = typename std::conditional<RB, RbRecord<MAXC>, NoRecord>::type;
    const Real* __restrict__ tab = tables_of<Real>(m);
    const int32_t n_quads = g.n_pad >> 2;
    const int64_t total = (int64_t)n_vars_c * n_quads;
    const int lane = threadIdx.x & 31;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < total; base += (int64_t)gridDim.x * blockDim.x) {
        const int64_t item = base + threadIdx.x;
        const bool active = item < total;
        const unsigned mask = __ballot_sync(0xffffffffu, active);
        if (!active) continue;
        const int32_t j = (int32_t)(item / n_quads);
        const int32_t q = (int32_t)(item - (int64_t)j * n_quads);
        const int32_t v = __ldg(vars + j);
        const int32_t card = __ldg(m.card + v);
        const uint32_t chain0 = (uint32_t)(g.first_chain + 4u * (uint32_t)q);
        int x[4];
        const int32_t tpo = hybrid ? __ldg(t.tp_off + v) : -1;
        if (tpo >= 0)
            tab_update_quad(t, tpo, g.state + 4 * (size_t)q, (uint32_t)g.n_pad, v, card, chain0, sweep, g.seed_lo, g.seed_hi, x,
                            (RB && record) ? g.counts + __ldg(m.card_off + v) : nullptr, g.n_chains - 4 * q);
        else {
            const Rec rec(record ? g.counts + __ldg(m.card_off + v) : nullptr, g.n_chains - 4 * q);
            lse_update_quad<Real, MAXC, CW, true, Rec>(m, tab, g.state + 4 * (size_t)q, (uint32_t)g.n_pad, v, card, chain0, sweep,
                                                       g.seed_lo, g.seed_hi, x, rec);
            rec.flush(card);
        }
        const uint32_t packed = (uint32_t)x[0] | ((uint32_t)x[1] << 8) | ((uint32_t)x[2] << 16) | ((uint32_t)x[3] << 24);
        *reinterpret_cast<uint32_t*>(g.state + (size_t)v * g.n_pad + 4 * q) = packed;

        if (record) {
            const int nvalid = min(4, g.n_chains - 4 * q);  // >= 1 for every launched quad except pure padding
            const int32_t coff = __ldg(m.card_off + v);
            // chain.go:231-236: Marginal[value] += 1 — aggregated over the warp when it holds one variable
            const int v0 = __shfl_sync(mask, v, __ffs(mask) - 1);
            const bool uniform = __all_sync(mask, v == v0);
            if constexpr (RB) {
                // Rao-Blackwell estimator: the conditional was added to the bins inside the update
            } else if (uniform) {
                for (int k = 0; k < card; k++) {
                    int c = 0;
#pragma unroll
                    for (int ci = 0; ci < 4; ci++) c += (ci < nvalid && x[ci] == k) ? 1 : 0;
                    const int s = __reduce_add_sync(mask, c);
                    if (lane == __ffs(mask) - 1 && s) atomicAdd(g.counts + coff + k, (unsigned long long)s);
                }
            } else {
#pragma unroll
                for (int ci = 0; ci < 4; ci++)
                    if (ci < nvalid) atomicAdd(g.counts + coff + x[ci], 1ull);
            }
            // chain.go:237: ChainHistory[v].Add(value) — kept as per-chain half-window histograms
            if (hist_half >= 0 && g.hist) {
#pragma unroll
                for (int ci = 0; ci < 4; ci++)
                    if (ci < nvalid) hist_add(nullptr, g, m.total_card, hist_half, coff + x[ci], 0, 0, 4 * q + ci);
            }
        }
    }
}

// chain.go:237 (ChainHistory[v].Add(value)) as per-chain half-window histograms: in the CTA's shared memory
// when the launch keeps them there (s_hist = [2][total_card][CH] u16, flushed once at the end of the
// launch), else read-modify-write in global memory.  entry = card_off[v] + value; cc = chain within the CTA.
__device__ __forceinline__ void hist_add(uint16_t* s_hist, const DevGroup& g, const int total_card, const int half,
                                         const int entry, const int CH, const int cc, const int lchain) {
    if (s_hist) {  // two code paths instead of one generic pointer: LDS/STS or LDG/STG
        uint16_t* h = s_hist + ((size_t)half * total_card + entry) * CH + cc;
        *h = (uint16_t)(*h + 1);
    } else {
        // global memory: a fire-and-forget 32-bit reduction on the word that holds this chain's 16-bit counter (two
        // chains per word; a half window holds < 65536 samples, so the low lane never carries into the high one) —
        // no load to wait for, unlike a 2-byte read-modify-write
        unsigned int* w = reinterpret_cast<unsigned int*>(g.hist + ((size_t)half * total_card + entry) * g.n_pad + (lchain & ~1));
        atomicAdd(w, 1u << (16 * (lchain & 1)));
    }
}
// Table kernels: the 8 chains of a work unit are contiguous, and a binary variable has two histogram rows, so the 16
// read-modify-writes of 2 bytes become two 16-byte ones: bit i of `ones` / `zeros` is added to the 16-bit lane of
// chain i (lanes cannot carry into each other: a half window holds < 65536 samples).  Rows are 16-byte aligned
// (CH, n_pad and the unit's first chain are multiples of 8).
__device__ __forceinline__ uint4 spread_bits16(const uint32_t b) {  // bit i -> 16-bit lane i of a uint4
    uint4 r;
    r.x = ((b & 3u) | ((b & 3u) << 15)) & 0x00010001u;
    r.y = (((b >> 2) & 3u) | (((b >> 2) & 3u) << 15)) & 0x00010001u;
    r.z = (((b >> 4) & 3u) | (((b >> 4) & 3u) << 15)) & 0x00010001u;
    r.w = (((b >> 6) & 3u) | (((b >> 6) & 3u) << 15)) & 0x00010001u;
    return r;
}
__device__ __forceinline__ void add_row16(uint4* row, const uint4 inc) {
    uint4 h = *row;
    h.x += inc.x; h.y += inc.y; h.z += inc.z; h.w += inc.w;
    *row = h;
}
__device__ __forceinline__ void red_row16(uint16_t* row, const uint4 inc) {  // the same on global memory, without the load
    unsigned int* w = reinterpret_cast<unsigned int*>(row);
    if (inc.x) atomicAdd(w, inc.x);
    if (inc.y) atomicAdd(w + 1, inc.y);
    if (inc.z) atomicAdd(w + 2, inc.z);
    if (inc.w) atomicAdd(w + 3, inc.w);
}
__device__ __forceinline__ void hist_add8_binary(uint16_t* s_hist, const DevGroup& g, const int total_card, const int half,
                                                 const int coff, const int CH, const int cc0, const int lchain0,
                                                 const uint32_t ones, const uint32_t zeros) {
    if (s_hist) {
        uint16_t* r0 = s_hist + ((size_t)half * total_card + coff) * CH + cc0;
        add_row16(reinterpret_cast<uint4*>(r0), spread_bits16(zeros));
        add_row16(reinterpret_cast<uint4*>(r0 + CH), spread_bits16(ones));
    } else {
        uint16_t* r0 = g.hist + ((size_t)half * total_card + coff) * g.n_pad + lchain0;
        red_row16(r0, spread_bits16(zeros));
        red_row16(r0 + g.n_pad, spread_bits16(ones));
    }
}
__device__ __forceinline__ uint16_t* hist_begin(uint8_t* smem, const int32_t hist_off, const DevGroup& g, const int total_card,
                                                const int CH) {
    if (hist_off < 0 || !g.hist) return nullptr;
    uint16_t* s_hist = reinterpret_cast<uint16_t*>(smem + hist_off);
    for (int i = threadIdx.x; i < 2 * total_card * CH; i += blockDim.x) s_hist[i] = 0;
    return s_hist;
}
__device__ __forceinline__ void hist_flush(const uint16_t* s_hist, const DevGroup& g, const int total_card, const int CH,
                                           const int cta_chain) {
    if (!s_hist) return;
    for (int i = threadIdx.x; i < 2 * total_card * CH; i += blockDim.x) {
        const int e = i / CH, c = i - e * CH;
        if (s_hist[i] && cta_chain + c < g.n_chains) {
            uint16_t* h = g.hist + (size_t)e * g.n_pad + cta_chain + c;
            *h = (uint16_t)(*h + s_hist[i]);
        }
    }
}

// ------------------------------------------------------------------ K1/K2 resident
// Small models (the bundled UAI problems): the whole state of CH chains fits in shared memory, so
// ONE launch runs many sweeps — every colour of every sweep — with __syncthreads() between
// colours instead of one launch per colour (which is launch- and latency-bound for 60..461
// variables).  A CTA owns CH consecutive chains; marginal counts accumulate in shared memory and
// are flushed once per launch.  Same arithmetic and Philox stream as k_sweep_colour: the two
// paths produce identical trajectories.  Window schedule of (*Chain).AdvanceChain: sweeps
// [0, n_pre) record only, then n_half sweeps into the first and n_half into the second
// half-window histogram (n_half < 0: no histograms).
// TS = the model's log-space tables are staged in shared memory for the whole launch (every sweep's
// table gathers become LDS): one elected thread issues TMA bulk copies (cp.async.bulk, 16-byte
// granules) that complete on an mbarrier while the CTA loads its chains' state.  TS = false reads the
// tables through L1 (__ldg) — used when they do not fit next to the state.
__device__ __forceinline__ void mbar_init(const uint32_t bar, const uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(const uint32_t bar, const uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(const uint32_t dst, const void* src, const uint32_t bytes, const uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(const uint32_t bar, const uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar),
        "r"(parity)
        : "memory");
}

constexpr uint32_t kBulkChunk = 32768;  // bytes per cp.async.bulk (multiple of 16)

template <typename Real, int MAXC, int CW, bool TS, bool RB = false>  // CW = 0: one thread per chain (lse_update_one)
__global__ void __launch_bounds__(CW == 0 ? 512 : 256)
k_sweep_resident(const DevModel m, const DevGroup g, const int32_t* __restrict__ order,
                 const int32_t* __restrict__ colour_off, const int32_t n_colours, const int32_t ch_per_cta,
                 const uint32_t sweep0, const int32_t n_sweeps, const int record, const int32_t n_pre,
                 const int32_t n_half, const DevTab t, const int hybrid, const int32_t hist_off, const int32_t n_stage) {
    // n_stage = table entries staged in shared memory (multiple of 4): all of them when TS, else a prefix that ends on
    // a factor boundary (possibly empty) — see StagedPrefix
    using Rec = typename std::conditional<RB, RbRecord<MAXC>, NoRecord>::type;
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_bar;
    uint8_t* s_state = smem;                                                                  // [n_vars][CH]
    const size_t counts_off = ((size_t)m.n_vars * ch_per_cta + 15) & ~(size_t)15;
    unsigned int* s_counts = reinterpret_cast<unsigned int*>(smem + counts_off);              // [total_card]
    const Real* __restrict__ tab = tables_of<Real>(m);
    StagedPrefix<Real> sp{nullptr, 0};
    const bool staged = TS || n_stage > 0;
    if (staged) {
        Real* s_tab = reinterpret_cast<Real*>(smem + ((counts_off + (size_t)m.total_card * 4 + 15) & ~(size_t)15));  // [n_stage]
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
        if (threadIdx.x == 0) mbar_init(bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t bytes = (uint32_t)((n_stage + 3) & ~3) * (uint32_t)sizeof(Real);  // 16-byte granules (the device tables are padded)
            mbar_expect_tx(bar, bytes);
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_tab);
            for (uint32_t off = 0; off < bytes; off += kBulkChunk)
                tma_bulk_g2s(dst + off, reinterpret_cast<const uint8_t*>(tab) + off, min(kBulkChunk, bytes - off), bar);
        }
        if constexpr (TS) tab = s_tab;
        else sp = StagedPrefix<Real>{s_tab, n_stage};
    }
    const int CH = ch_per_cta;
    const int cta_chain = blockIdx.x * CH;  // local index of this CTA's first chain
    const int n_quads = CH >> 2;
    // load state (CH bytes per variable) and clear the counts
    for (int i = threadIdx.x; i < m.n_vars * n_quads; i += blockDim.x) {
        const int v = i / n_quads, q = i - v * n_quads;
        *reinterpret_cast<uint32_t*>(s_state + (size_t)v * CH + 4 * q) =
            *reinterpret_cast<const uint32_t*>(g.state + (size_t)v * g.n_pad + cta_chain + 4 * q);
    }
    for (int i = threadIdx.x; i < m.total_card; i += blockDim.x) s_counts[i] = 0;
    uint16_t* s_hist = hist_begin(smem, n_half >= 0 ? hist_off : -1, g, m.total_card, ch_per_cta);
    if (staged) mbar_wait((uint32_t)__cvta_generic_to_shared(&s_bar), 0);  // the tables have landed
    __syncthreads();
    for (int s = 0; s < n_sweeps; s++) {
        const uint32_t sweep = sweep0 + (uint32_t)s;
        const int hist_half = (n_half < 0 || s < n_pre) ? -1 : (s < n_pre + n_half ? 0 : 1);
        for (int col = 0; col < n_colours; col++) {
            const int c0 = __ldg(colour_off + col), nvc = __ldg(colour_off + col + 1) - c0;
            if constexpr (CW == 0) {
                const int ch_shift = 31 - __clz(CH);  // CH is a power of two (8..64)
                for (int item = threadIdx.x; item < nvc * CH; item += blockDim.x) {
                    const int j = item >> ch_shift, cc = item & (CH - 1);
                    const int4 pr = __ldg(m.pos_rec + c0 + j);  // variable, cardinality, pairwise records, factors
                    const int v = pr.x;
                    const int lchain = cta_chain + cc;
                    const uint32_t chain = (uint32_t)(g.first_chain + (uint64_t)lchain);
                    const Rec rec(record ? g.counts + __ldg(m.card_off + v) : nullptr, g.n_chains - lchain);
                    const int x = lse_update_one<Real, MAXC, !TS, Rec>(m, tab, s_state + cc, (uint32_t)CH, v, pr.y, pr.z, pr.w, chain,
                                                                       sweep, g.seed_lo, g.seed_hi, rec, sp);
                    rec.flush(pr.y);
                    s_state[(size_t)v * CH + cc] = (uint8_t)x;
                    if (record && lchain < g.n_chains) {
                        const int32_t coff = __ldg(m.card_off + v);
                        if constexpr (!RB) atomicAdd(&s_counts[coff + x], 1u);
                        if (hist_half >= 0 && g.hist) hist_add(s_hist, g, m.total_card, hist_half, coff + x, CH, cc, lchain);
                    }
                }
            } else {
                for (int item = threadIdx.x; item < nvc * n_quads; item += blockDim.x) {
                    const int j = item / n_quads, q = item - j * n_quads;
                    const int v = __ldg(order + c0 + j);
                    const int card = __ldg(m.card + v);
                    const int lchain = cta_chain + 4 * q;  // local chain index of the quad
                    const uint32_t chain0 = (uint32_t)(g.first_chain + (uint64_t)lchain);
                    int x[4];
                    const int32_t tpo = hybrid ? __ldg(t.tp_off + v) : -1;
                    if (tpo >= 0)
                        tab_update_quad(t, tpo, s_state + 4 * q, (uint32_t)CH, v, card, chain0, sweep, g.seed_lo, g.seed_hi, x,
                                        (RB && record) ? g.counts + __ldg(m.card_off + v) : nullptr, g.n_chains - lchain);
                    else {
                        const Rec rec(record ? g.counts + __ldg(m.card_off + v) : nullptr, g.n_chains - lchain);
                        lse_update_quad<Real, MAXC, (CW == 0 ? 1 : CW), !TS, Rec>(m, tab, s_state + 4 * q, (uint32_t)CH, v, card, chain0,
                                                                                  sweep, g.seed_lo, g.seed_hi, x, rec);
                        rec.flush(card);
                    }
                    *reinterpret_cast<uint32_t*>(s_state + (size_t)v * CH + 4 * q) =
                        (uint32_t)x[0] | ((uint32_t)x[1] << 8) | ((uint32_t)x[2] << 16) | ((uint32_t)x[3] << 24);
                    if (record) {
                        const int nvalid = max(0, min(4, g.n_chains - lchain));
                        const int32_t coff = __ldg(m.card_off + v);
#pragma unroll
                        for (int ci = 0; ci < 4; ci++)
                            if (ci < nvalid && !RB) atomicAdd(&s_counts[coff + x[ci]], 1u);
                        if (hist_half >= 0 && g.hist) {
#pragma unroll
                            for (int ci = 0; ci < 4; ci++)
                                if (ci < nvalid) hist_add(s_hist, g, m.total_card, hist_half, coff + x[ci], CH, 4 * q + ci, lchain + ci);
                        }
                    }
                }
            }
            __syncthreads();
        }
    }
    // write the state back and flush the counts
    for (int i = threadIdx.x; i < m.n_vars * n_quads; i += blockDim.x) {
        const int v = i / n_quads, q = i - v * n_quads;
        *reinterpret_cast<uint32_t*>(g.state + (size_t)v * g.n_pad + cta_chain + 4 * q) =
            *reinterpret_cast<const uint32_t*>(s_state + (size_t)v * CH + 4 * q);
    }
    if (record)
        for (int i = threadIdx.x; i < m.total_card; i += blockDim.x)
            if (s_counts[i]) atomicAdd(g.counts + i, (unsigned long long)s_counts[i]);
    hist_flush(s_hist, g, m.total_card, CH, cta_chain);
}

// ------------------------------------------------------------------ K1-table
// Tabulated-conditional sweep for models whose sampled variables are binary with <= 256
// neighbour configurations (host_model.hpp::build_tab_programs).  The conditional of
// gibbs-simple.go:171-258 is evaluated in float64 once per (variable, neighbour configuration)
// by k_build_thresholds and stored as the largest 32-bit draw that still selects value 0 under
// the reference's inverse-CDF rule (sampler.go:115-123: r = U*tot, r <= e0).  The sweep itself is
// integer work: gather neighbour bytes -> configuration index -> threshold -> compare.
// sampler.go:107-123 for a 32-bit draw u (U = u * 2^-32): re-sum, r = U * tot, first k with r <= w[k] (the fall-through
// selects the last value, as everywhere on the device).  Monotone non-decreasing in u.
template <int CARD>
__device__ __forceinline__ int tab_select(const double (&w)[CARD], const uint32_t u) {
    double tot = 0.0;
#pragma unroll
    for (int k = 0; k < CARD; k++) tot += w[k];
    double r = ((double)u * (1.0 / 4294967296.0)) * tot;
#pragma unroll
    for (int k = 0; k < CARD - 1; k++) {
        if (r <= w[k]) return k;
        r -= w[k];
    }
    return CARD - 1;
}
// thresholds of one configuration: thr[j] = the largest 32-bit draw that still selects a value <= j (j < CARD - 1), found
// by bisection on the reference's own predicate, so value = #{j : u > thr[j]} reproduces the inverse CDF exactly
template <int CARD>
__device__ __forceinline__ void tab_thresholds(const DevModel& m, const int32_t* __restrict__ prog, const int32_t* __restrict__ tp,
                                               const int nn, const int cfg, uint32_t* __restrict__ out) {
    double w[CARD];
#pragma unroll
    for (int k = 0; k < CARD; k++) w[k] = 0.0;
    const int nf = prog[0];
    const int32_t* p = prog + 1;
    for (int f = 0; f < nf; f++) {
        const int tab_off = p[0], sv = p[1], no = p[2];
        p += 3;
        int b = tab_off;
        for (int o = 0; o < no; o++, p += 2) {
            const int ov = p[0];
            int sv_o = m.fixed[ov];
            if (sv_o < 0) {
                for (int i = 0; i < nn; i++)
                    if (tp[2 + 2 * i] == ov) sv_o = (cfg / tp[3 + 2 * i]) % m.card[ov];
            }
            b += sv_o * p[1];
        }
#pragma unroll
        for (int k = 0; k < CARD; k++) w[k] += m.tab64[b + k * sv];
    }
    stabilise_exp_floor<double, CARD>(w, CARD);
#pragma unroll
    for (int j = 0; j < CARD - 1; j++) {
        uint32_t lo = 0u, hi = 0xffffffffu;  // invariant: select(lo) <= j (u = 0 always selects value 0: r = 0 <= w[0])
        while (lo < hi) {
            const uint32_t mid = lo + (uint32_t)(((uint64_t)hi - lo + 1) >> 1);
            if (tab_select<CARD>(w, mid) <= j) lo = mid;
            else hi = mid - 1u;
        }
        out[j] = lo;
    }
}

static __global__ void __launch_bounds__(128)
k_build_thresholds(const DevModel m, const DevTab t, const int32_t* __restrict__ order, const int32_t n_order) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_order; j += gridDim.x * blockDim.x) {
        const int v = order[j];
        if (t.tp_off[v] < 0) continue;  // hybrid mode: this variable is sampled by the log-sum-exp path
        const int32_t* tp = t.tprog + t.tp_off[v];
        const int nn = tp[0], thr_off = tp[1], card = m.card[v];
        int n_cfg = 1;
        for (int i = 0; i < nn; i++) n_cfg *= m.card[tp[2 + 2 * i]];
        const int32_t* prog = m.prog + m.prog_off[v];
        for (int cfg = 0; cfg < n_cfg; cfg++) {
            uint32_t* out = t.thr + thr_off + (size_t)cfg * (card - 1);
            if (card == 2) tab_thresholds<2>(m, prog, tp, nn, cfg, out);
            else if (card == 3) tab_thresholds<3>(m, prog, tp, nn, cfg, out);
            else tab_thresholds<4>(m, prog, tp, nn, cfg, out);
        }
    }
}

// CTA tile = (chunk of 2048 consecutive chains) x (VB consecutive sweep positions of the colour);
// every thread keeps the same 8 chains (one 64-bit state word per variable) for the whole tile, so
// the wave front of concurrently processed tiles stays L2-resident and shared neighbours of
// consecutive variables hit L1.  Per tile the position records and the 16-bit high halves of the
// thresholds are staged in shared memory.  One Philox call yields the high halves of 8 draws; the
// low halves are generated only when a high half ties with its threshold (probability 2^-16).
constexpr int kTabRec = 20;  // {v, thr_off, n_nbr | card << 8, card_off, nbr[8], stride[8]}; wide (n_nbr > 8): {.., tprog offset, true n_nbr}

// NN = neighbour slots read per variable (4 or 8); records pad unused slots with the variable
// itself at stride 0, so the loads are unconditional and branch-free.
// Neighbour rows belong to other colours, so they are read-only for the whole launch: the
// non-coherent global path is safe and keeps the loads out of the generic address space.
__device__ __forceinline__ uint2 ld_state8(const uint64_t addr) { return __ldg(reinterpret_cast<const uint2*>(addr)); }
__device__ __forceinline__ void st_state8(const uint64_t addr, const uint2 v) {
    asm volatile("st.global.v2.u32 [%0], {%1, %2};" ::"l"(addr), "r"(v.x), "r"(v.y) : "memory");
}

// Asynchronous 8-byte global -> shared copies (LDGSTS): completion is tracked per thread by commit
// groups, not by a register scoreboard, so several variables' neighbour words can be in flight at once.
__device__ __forceinline__ void cp_async8(const uint32_t smem_addr, const uint64_t gaddr) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr), "l"(gaddr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint2 lds_state8(const uint32_t smem_addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(smem_addr) : "memory");
    return v;
}

// NN = neighbour slots read per variable (4 or 8); records pad unused slots with the variable
// itself at stride 0, so the loads are unconditional and branch-free.  HIST = keep the per-chain
// half-window histograms (chain.go:237).  All shared-memory accesses go through the array symbols
// (not through generic pointers) so they compile to LDS with immediate bases.
// PF = prefetch depth in variables.  PF == 0: the next variable's words are loaded into a second
// register set (LDG; ptxas tracks both sets on one scoreboard, so every second variable waits for
// loads issued a few instructions earlier).  PF > 0: each thread copies the words of variable j + PF
// into its own slots of a shared-memory ring of PF + 1 stages with cp.async while variable j
// computes; the ring is private to the thread (no barrier), dynamic shared memory =
// (PF + 1) * NN * 256 * 8 bytes.
template <int VB, int NN, bool HIST, int PF>
__global__ void __launch_bounds__(256)
k_sweep_tab(const DevModel m, const DevTab t, const DevGroup g, const int32_t j_begin, const int32_t n_vars_c,
            const uint32_t sweep, const int record, const int hist_half) {
    __shared__ int4 s_rec[VB * (kTabRec / 4)];
    __shared__ unsigned int s_cnt[VB / 2];  // ones per variable: word 4*(j/8) + j%4 holds variables j (low 16 bits) and j+4 (high)
    __shared__ uint16_t s_thr[VB * (NN == 4 ? 16 : 256)];  // <= 2^NN configurations per variable
    const int units = g.n_pad >> 3;
    const int chunks = (units + 255) >> 8;
    const int n_vb = (n_vars_c + VB - 1) / VB;
    const int64_t n_tiles = (int64_t)chunks * n_vb;
    const uint32_t n_pad = (uint32_t)g.n_pad;
    const uint32_t seed_lo = g.seed_lo, seed_hi = g.seed_hi;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int chunk = (int)(tile / n_vb);
        const int vb = (int)(tile - (int64_t)chunk * n_vb);
        // A warp runs if ANY of its lanes owns chains; lanes past the end of a ragged last chunk redo the
        // last unit's work with their store and counts masked off, so warp-wide shuffles always see 32 lanes.
        const int unit_raw = chunk * 256 + threadIdx.x;
        const bool lane_valid = unit_raw < units;
        const bool active = chunk * 256 + (int)(threadIdx.x & ~31u) < units;
        const int unit = min(unit_raw, units - 1);
        const int nvalid = lane_valid ? max(0, min(8, g.n_chains - 8 * unit)) : 0;
        const uint32_t vmask = (nvalid >= 8) ? 0xffu : ((1u << nvalid) - 1u);
        const uint32_t chain_blk = (uint32_t)((g.first_chain >> 3) + (uint64_t)unit);
        const int nv = min(VB, n_vars_c - vb * VB);
        const int j0 = j_begin + vb * VB;
        {  // stage the tile's records and threshold high halves
            const int32_t* src = t.trec + (size_t)j0 * kTabRec;
            int32_t* dst = reinterpret_cast<int32_t*>(s_rec);
            for (int i = threadIdx.x; i < nv * kTabRec; i += 256) dst[i] = __ldg(src + i);
            const int ta = __ldg(src + 1);
            const int tb = (j0 + nv < t.n_order) ? __ldg(t.trec + (size_t)(j0 + nv) * kTabRec + 1) : t.n_thr;
            for (int i = threadIdx.x; i < tb - ta; i += 256) s_thr[i] = (uint16_t)(__ldg(t.thr + ta + i) >> 16);
            if (threadIdx.x < VB / 2) s_cnt[threadIdx.x] = 0;
        }
        __syncthreads();
        if (active) {
            const int thr_a = s_rec[0].y;
            const uint64_t my = reinterpret_cast<uint64_t>(g.state) + 8ull * (uint64_t)unit;
            constexpr unsigned amask = 0xffffffffu;
            const bool leader = (threadIdx.x & 31) == 0;

            auto load_nbrs = [&](const int j, uint2(&w)[NN]) {
                const uint64_t base = my;
                const int4 na = s_rec[j * 5 + 1];
                w[0] = ld_state8(base + (uint64_t)(uint32_t)na.x * n_pad);
                w[1] = ld_state8(base + (uint64_t)(uint32_t)na.y * n_pad);
                w[2] = ld_state8(base + (uint64_t)(uint32_t)na.z * n_pad);
                w[3] = ld_state8(base + (uint64_t)(uint32_t)na.w * n_pad);
                if constexpr (NN == 8) {
                    const int4 nb = s_rec[j * 5 + 2];
                    w[4] = ld_state8(base + (uint64_t)(uint32_t)nb.x * n_pad);
                    w[5] = ld_state8(base + (uint64_t)(uint32_t)nb.y * n_pad);
                    w[6] = ld_state8(base + (uint64_t)(uint32_t)nb.z * n_pad);
                    w[7] = ld_state8(base + (uint64_t)(uint32_t)nb.w * n_pad);
                }
            };
            // one variable x 8 chains: configuration indices, one Philox call, threshold compare, store
            auto update = [&](const int j, const uint2(&w)[NN], uint32_t& acc, const int slot) {
                const int4 hd = s_rec[j * 5];  // v, thr_off, n_nbr, card_off
                const int4 sa = s_rec[j * 5 + 3];
                uint32_t cfg_lo = w[0].x * (uint32_t)sa.x + w[1].x * (uint32_t)sa.y + w[2].x * (uint32_t)sa.z + w[3].x * (uint32_t)sa.w;
                uint32_t cfg_hi = w[0].y * (uint32_t)sa.x + w[1].y * (uint32_t)sa.y + w[2].y * (uint32_t)sa.z + w[3].y * (uint32_t)sa.w;
                if constexpr (NN == 8) {
                    const int4 sb = s_rec[j * 5 + 4];
                    cfg_lo += w[4].x * (uint32_t)sb.x + w[5].x * (uint32_t)sb.y + w[6].x * (uint32_t)sb.z + w[7].x * (uint32_t)sb.w;
                    cfg_hi += w[4].y * (uint32_t)sb.x + w[5].y * (uint32_t)sb.y + w[6].y * (uint32_t)sb.z + w[7].y * (uint32_t)sb.w;
                }
                const Philox4 a = philox_wide((uint32_t)hd.x, sweep, chain_blk, kTagDraw16Hi, seed_lo, seed_hi);
                const uint32_t wa[4] = {a.x, a.y, a.z, a.w};
                const int toff = hd.y - thr_a;
                uint32_t xbits = 0, dd[8];
                // NN == 4: at most 16 configurations, so the variable's thresholds live one per lane and
                // are fetched with a warp shuffle (no address arithmetic); NN == 8: shared-memory lookup
                uint32_t my_th = 0;
                if constexpr (NN == 4) my_th = s_thr[toff + (threadIdx.x & 15)];
#pragma unroll
                for (int i = 7; i >= 0; i--) {
                    const uint32_t idx = __byte_perm(i < 4 ? cfg_lo : cfg_hi, 0, 0x4440 + (i & 3));
                    uint32_t th;
                    if constexpr (NN == 4) th = __shfl_sync(amask, my_th, idx);
                    else th = s_thr[toff + idx];
                    const uint32_t hi = (i & 1) ? (wa[i >> 1] >> 16) : __byte_perm(wa[i >> 1], 0, 0x4410);
                    const uint32_t d = th - hi;            // sign bit set <=> hi > th  (both < 2^16)
                    xbits = __funnelshift_l(d, xbits, 1);  // xbits = (xbits << 1) | (hi > th)
                    dd[i] = d;
                }
                // a high half ties with its threshold <=> some d is zero <=> the unsigned minimum is zero
                const uint32_t tie = __vimin3_u32(__vimin3_u32(dd[0], dd[1], dd[2]), __vimin3_u32(dd[3], dd[4], dd[5]), min(dd[6], dd[7]));
                if (tie == 0) {  // resolve ties with the low halves: draw > threshold <=> lo16 > (T & 0xffff)
                    const Philox4 b = philox_wide((uint32_t)hd.x, sweep, chain_blk, kTagDraw16Lo, seed_lo, seed_hi);
                    const uint32_t wb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const uint32_t idx = ((i < 4 ? cfg_lo : cfg_hi) >> (8 * (i & 3))) & 0xffu;
                        const uint32_t T = __ldg(t.thr + hd.y + idx);
                        const uint32_t hi = (wa[i >> 1] >> (16 * (i & 1))) & 0xffffu;
                        const uint32_t lo = (wb[i >> 1] >> (16 * (i & 1))) & 0xffffu;
                        if (hi == (T >> 16) && lo > (T & 0xffffu)) xbits |= 1u << i;
                    }
                }
                uint2 outw;  // spread decision bits into state bytes: bit i -> byte i
                outw.x = ((xbits & 0xfu) * 0x00204081u) & 0x01010101u;
                outw.y = (((xbits >> 4) & 0xfu) * 0x00204081u) & 0x01010101u;
                if (lane_valid) st_state8(my + (uint64_t)(uint32_t)hd.x * n_pad, outw);
                if constexpr (HIST) {
                    if (record && hist_half >= 0 && lane_valid)
                        hist_add8_binary(nullptr, g, m.total_card, hist_half, hd.w, 0, 0, 8 * unit, xbits & vmask, ~xbits & vmask);
                }
                // chain.go:231-236: this thread's ones of the variable go into nibble `slot` of the group accumulator
                acc += (uint32_t)__popc(xbits & vmask) << (4 * slot);
            };

            // software pipeline: the next variable's neighbour words are in flight while this one
            // computes (variables of one colour are never neighbours, so the early loads cannot see
            // this tile's writes)
            // Variables are taken in groups of 8 so that the nibble slot is a compile-time constant;
            // per group the 8 nibbles are reduced over the warp as four words of two 16-bit fields
            // (<= 256 per field) and added to the CTA's shared counters by the warp leader.
            [[maybe_unused]] uint2 wA[NN], wB[NN];
            [[maybe_unused]] uint32_t pf_base = 0;
            if constexpr (PF > 0) {
                static_assert(8 % (PF + 1) == 0, "ring stages must divide the unroll factor");
                extern __shared__ __align__(16) uint8_t s_ring[];  // [PF + 1][NN][256] uint2
                pf_base = (uint32_t)__cvta_generic_to_shared(s_ring) + 8u * threadIdx.x;
            }
            [[maybe_unused]] auto issue = [&](const int j, const int stage) {
                const int4 na = s_rec[j * 5 + 1];
                cp_async8(pf_base + (stage * NN + 0) * 2048, my + (uint64_t)(uint32_t)na.x * n_pad);
                cp_async8(pf_base + (stage * NN + 1) * 2048, my + (uint64_t)(uint32_t)na.y * n_pad);
                cp_async8(pf_base + (stage * NN + 2) * 2048, my + (uint64_t)(uint32_t)na.z * n_pad);
                cp_async8(pf_base + (stage * NN + 3) * 2048, my + (uint64_t)(uint32_t)na.w * n_pad);
                if constexpr (NN == 8) {
                    const int4 nb = s_rec[j * 5 + 2];
                    cp_async8(pf_base + (stage * NN + 4) * 2048, my + (uint64_t)(uint32_t)nb.x * n_pad);
                    cp_async8(pf_base + (stage * NN + 5) * 2048, my + (uint64_t)(uint32_t)nb.y * n_pad);
                    cp_async8(pf_base + (stage * NN + 6) * 2048, my + (uint64_t)(uint32_t)nb.z * n_pad);
                    cp_async8(pf_base + (stage * NN + 7) * 2048, my + (uint64_t)(uint32_t)nb.w * n_pad);
                }
            };
            if constexpr (PF > 0) {
#pragma unroll
                for (int d = 0; d < PF; d++) {
                    if (d < nv) issue(d, d);
                    cp_async_commit();
                }
            } else {
                load_nbrs(0, wA);
            }
            for (int jg = 0; jg < nv; jg += 8) {
                uint32_t acc = 0;
                if constexpr (PF > 0) {
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const int j = jg + u;
                        if (j < nv) {
                            cp_async_wait<PF - 1>();  // variable j's words have landed (one group per variable)
#pragma unroll
                            for (int k = 0; k < NN; k++) wA[k] = lds_state8(pf_base + ((u % (PF + 1)) * NN + k) * 2048);
                            // refill the stage consumed by the PREVIOUS variable (its words are in registers)
                            if (j + PF < nv) issue(j + PF, (u + PF) % (PF + 1));
                            cp_async_commit();
                            update(j, wA, acc, u);
                        }
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < 8; u += 2) {
                        const int jj = jg + u;
                        if (jj < nv) {
                            if (jj + 1 < nv) load_nbrs(jj + 1, wB);
                            update(jj, wA, acc, u);
                            if (jj + 1 < nv) {
                                if (jj + 2 < nv) load_nbrs(jj + 2, wA);
                                update(jj + 1, wB, acc, u + 1);
                            }
                        }
                    }
                }
                if (record) {
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const unsigned s = __reduce_add_sync(amask, (acc >> (4 * k)) & 0x000f000fu);
                        if (leader && s) atomicAdd(&s_cnt[(jg >> 1) + k], s);
                    }
                }
            }
        }
        __syncthreads();
        if (record && threadIdx.x < nv) {
            const int32_t coff = s_rec[threadIdx.x * 5].w;
            const unsigned o = (s_cnt[(threadIdx.x >> 3) * 4 + (threadIdx.x & 3)] >> (4 * (threadIdx.x & 4))) & 0xffffu;
            const int valid = max(0, min(2048, g.n_chains - chunk * 2048));
            if (o) atomicAdd(g.counts + coff + 1, (unsigned long long)o);
            if (valid - (int)o) atomicAdd(g.counts + coff, (unsigned long long)(valid - (int)o));
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ K1-table resident
// Table mode for small models: like k_sweep_resident, a CTA keeps the state of CH (multiple of 8)
// consecutive chains in shared memory and runs every colour of every sweep of a round in one launch.
// Work item = (sweep position, unit of 8 chains); same records, thresholds, Philox stream and tie rule
// as k_sweep_tab, so the two paths produce identical trajectories.  Records and thresholds are read
// through L1 (they are a few KB and shared by every CTA).
// WIDE: some variable has more than 8 free neighbours or 256 configurations (variable-length records, 32-bit configuration
// indices).  MULTI: some sampled variable is ternary / quaternary (card - 1 cumulative thresholds per configuration,
// full 32-bit draws, value = number of thresholds the draw exceeds); binary variables keep the 16-bit fast path.
template <bool WIDE, bool MULTI, bool RB>  // RB: Rao-Blackwell bins instead of counts (g.rb is set)
__global__ void __launch_bounds__(256)
k_sweep_tab_resident(const DevModel m, const DevTab t, const DevGroup g, const int32_t* __restrict__ colour_off,
                     const int32_t n_colours, const int32_t ch_per_cta, const uint32_t sweep0, const int32_t n_sweeps,
                     const int record, const int32_t n_pre, const int32_t n_half, const int32_t hist_off) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* s_state = smem;  // [n_vars][CH]
    unsigned int* s_counts = reinterpret_cast<unsigned int*>(smem + (((size_t)m.n_vars * ch_per_cta + 15) & ~(size_t)15));  // [total_card]
    const int CH = ch_per_cta;
    const int cta_chain = blockIdx.x * CH;
    const int units = CH >> 3;
    const int unit_shift = 31 - __clz(units);
    for (int i = threadIdx.x; i < m.n_vars * units; i += blockDim.x) {
        const int v = i / units, q = i - v * units;
        *reinterpret_cast<uint2*>(s_state + (size_t)v * CH + 8 * q) =
            *reinterpret_cast<const uint2*>(g.state + (size_t)v * g.n_pad + cta_chain + 8 * q);
    }
    for (int i = threadIdx.x; i < m.total_card; i += blockDim.x) s_counts[i] = 0;
    // Shared-memory histograms of this kernel hold only the ONES of every (binary) sampled variable, indexed by sweep
    // position: [2][n_order][CH] u16 — half the footprint of both rows, and a warp's consecutive positions touch
    // consecutive 16-byte rows.  The zeros are (recorded sweeps of the half) - ones, added at the flush.
    uint16_t* s_hist = nullptr;
    if (n_half >= 0 && hist_off >= 0 && g.hist) {
        s_hist = reinterpret_cast<uint16_t*>(smem + hist_off);
        for (int i = threadIdx.x; i < 2 * t.n_order * (ch_per_cta >> 3); i += blockDim.x)
            reinterpret_cast<uint4*>(s_hist)[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    for (int s = 0; s < n_sweeps; s++) {
        const uint32_t sweep = sweep0 + (uint32_t)s;
        const int hist_half = (n_half < 0 || s < n_pre) ? -1 : (s < n_pre + n_half ? 0 : 1);
        for (int col = 0; col < n_colours; col++) {
            const int c0 = __ldg(colour_off + col), nvc = __ldg(colour_off + col + 1) - c0;
            for (int item = threadIdx.x; item < nvc * units; item += blockDim.x) {
                const int j = item >> unit_shift, q = item & (units - 1);  // units is a power of two (CH = 8..64)
                const int4* r = reinterpret_cast<const int4*>(t.trec + (size_t)(c0 + j) * kTabRec);
                const int4 hd = __ldg(r);  // v, thr_off, n_nbr | card << 8, card_off
                const int nn = hd.z & 0xff;
                [[maybe_unused]] const int card = hd.z >> 8;
                uint32_t cfg_lo = 0, cfg_hi = 0;
                [[maybe_unused]] uint32_t idxw[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
                const bool wide = WIDE && nn > 8;
                if (wide) {
                    // wide variable (more than 8 free neighbours or 256 configurations, e.g. the neighbours of a collapsed
                    // variable): the record points at its variable-length tprog entry and the index needs 32 bits
                    const int32_t* __restrict__ tp = t.tprog + __ldg(&r[1].x) + 2;
                    const int n_true = __ldg(&r[1].y);
                    for (int i = 0; i < n_true; i++) {
                        const int ov = __ldg(tp + 2 * i);
                        const uint32_t os = (uint32_t)__ldg(tp + 2 * i + 1);
                        const uint2 w = *reinterpret_cast<const uint2*>(s_state + (size_t)ov * CH + 8 * q);
#pragma unroll
                        for (int ci = 0; ci < 4; ci++) {
                            idxw[ci] += ((w.x >> (8 * ci)) & 0xffu) * os;
                            idxw[4 + ci] += ((w.y >> (8 * ci)) & 0xffu) * os;
                        }
                    }
                } else {
                    {
                        const int4 na = __ldg(r + 1), sa = __ldg(r + 3);
                        const int nb[4] = {na.x, na.y, na.z, na.w}, sb[4] = {sa.x, sa.y, sa.z, sa.w};
#pragma unroll
                        for (int i = 0; i < 4; i++) {  // slots past n_nbr hold the variable itself at stride 0
                            const uint2 w = *reinterpret_cast<const uint2*>(s_state + (size_t)nb[i] * CH + 8 * q);
                            cfg_lo += w.x * (uint32_t)sb[i];
                            cfg_hi += w.y * (uint32_t)sb[i];
                        }
                    }
                    if (nn > 4) {
                        const int4 na = __ldg(r + 2), sa = __ldg(r + 4);
                        const int nb[4] = {na.x, na.y, na.z, na.w}, sb[4] = {sa.x, sa.y, sa.z, sa.w};
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const uint2 w = *reinterpret_cast<const uint2*>(s_state + (size_t)nb[i] * CH + 8 * q);
                            cfg_lo += w.x * (uint32_t)sb[i];
                            cfg_hi += w.y * (uint32_t)sb[i];
                        }
                    }
                }
                const int lchain = cta_chain + 8 * q;
                const uint32_t chain_blk = (uint32_t)((g.first_chain + (uint64_t)lchain) >> 3);
                const Philox4 a = philox_wide((uint32_t)hd.x, sweep, chain_blk, kTagDraw16Hi, g.seed_lo, g.seed_hi);
                const uint32_t wa[4] = {a.x, a.y, a.z, a.w};
                if constexpr (MULTI) {
                    if (card > 2) {  // ternary / quaternary variable: value = number of cumulative thresholds the 32-bit draw exceeds
                        const int nvalid = record ? max(0, min(8, g.n_chains - lchain)) : 0;
                        // first on the draws' high halves (one Philox call, like the binary path); the low halves are
                        // generated only when a high half ties with a threshold's (probability ~ 2^-16 per comparison)
                        uint32_t Tk[8][3], hi16[8];
                        int val[8];
                        bool tie = false;
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            uint32_t idx = ((i < 4 ? cfg_lo : cfg_hi) >> (8 * (i & 3))) & 0xffu;
                            if constexpr (WIDE) idx = wide ? idxw[i] : idx;
                            const uint32_t* __restrict__ T = t.thr + hd.y + idx * (uint32_t)(card - 1);
                            hi16[i] = (wa[i >> 1] >> (16 * (i & 1))) & 0xffffu;
                            val[i] = 0;
#pragma unroll
                            for (int k = 0; k < 3; k++) {
                                Tk[i][k] = k < card - 1 ? __ldg(T + k) : 0xffffffffu;  // (a threshold no draw exceeds)
                                val[i] += hi16[i] > (Tk[i][k] >> 16) ? 1 : 0;
                                tie |= k < card - 1 && hi16[i] == (Tk[i][k] >> 16);
                            }
                        }
                        if (tie) {
                            const Philox4 b = philox_wide((uint32_t)hd.x, sweep, chain_blk, kTagDraw16Lo, g.seed_lo, g.seed_hi);
                            const uint32_t wb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                            for (int i = 0; i < 8; i++) {
                                const uint32_t u = (hi16[i] << 16) | ((wb[i >> 1] >> (16 * (i & 1))) & 0xffffu);
                                val[i] = 0;
#pragma unroll
                                for (int k = 0; k < 3; k++) val[i] += (k < card - 1 && u > Tk[i][k]) ? 1 : 0;
                            }
                        }
                        uint32_t outb[2] = {0u, 0u}, cpack = 0u;  // state bytes; per-value counts of the valid chains, one byte each
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            outb[i >> 2] |= (uint32_t)val[i] << (8 * (i & 3));
                            if (i < nvalid) {
                                cpack += 1u << (8 * val[i]);
                                if (hist_half >= 0 && g.hist)  // (the shared-memory histograms hold the ones of binary variables only)
                                    hist_add(nullptr, g, m.total_card, hist_half, hd.w + val[i], CH, 0, lchain + i);
                                if constexpr (RB) {
                                    uint64_t prev = 0;
                                    for (int k = 0; k < card - 1; k++) {
                                        atomicAdd(g.counts + hd.w + k, (unsigned long long)rb_units((uint64_t)Tk[i][k] - prev));
                                        prev = (uint64_t)Tk[i][k];
                                    }
                                    atomicAdd(g.counts + hd.w + card - 1, (unsigned long long)rb_units(4294967296ull - prev));
                                }
                            }
                        }
                        if constexpr (!RB) {
                            for (int k = 0; k < card; k++) {
                                const uint32_t n = (cpack >> (8 * k)) & 0xffu;
                                if (n) atomicAdd(&s_counts[hd.w + k], n);
                            }
                        }
                        *reinterpret_cast<uint2*>(s_state + (size_t)hd.x * CH + 8 * q) = make_uint2(outb[0], outb[1]);
                        continue;
                    }
                }
                // same compare as k_sweep_tab: d = threshold high half - draw high half (both < 2^16), sign bit set <=> the
                // draw is above the threshold; a zero d is a tie, found with 3-input unsigned minima
                uint32_t T[8], dd[8], xbits = 0;
                const uint32_t* __restrict__ thr_v = t.thr + hd.y;
#pragma unroll
                for (int i = 7; i >= 0; i--) {
                    uint32_t idx = __byte_perm(i < 4 ? cfg_lo : cfg_hi, 0, 0x4440 + (i & 3));
                    if constexpr (WIDE) idx = wide ? idxw[i] : idx;
                    T[i] = __ldg(thr_v + idx);
                    const uint32_t hi = (i & 1) ? (wa[i >> 1] >> 16) : __byte_perm(wa[i >> 1], 0, 0x4410);
                    const uint32_t d = (T[i] >> 16) - hi;
                    xbits = __funnelshift_l(d, xbits, 1);  // xbits = (xbits << 1) | (hi > threshold)
                    dd[i] = d;
                }
                const bool tie = __vimin3_u32(__vimin3_u32(dd[0], dd[1], dd[2]), __vimin3_u32(dd[3], dd[4], dd[5]), min(dd[6], dd[7])) == 0;
                if (tie) {  // draw > threshold <=> high halves equal and lo16 > (T & 0xffff)
                    const Philox4 b = philox_wide((uint32_t)hd.x, sweep, chain_blk, kTagDraw16Lo, g.seed_lo, g.seed_hi);
                    const uint32_t wb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const uint32_t hi = (wa[i >> 1] >> (16 * (i & 1))) & 0xffffu;
                        const uint32_t lo = (wb[i >> 1] >> (16 * (i & 1))) & 0xffffu;
                        if (hi == (T[i] >> 16) && lo > (T[i] & 0xffffu)) xbits |= 1u << i;
                    }
                }
                uint2 outw;
                outw.x = ((xbits & 0xfu) * 0x00204081u) & 0x01010101u;
                outw.y = (((xbits >> 4) & 0xfu) * 0x00204081u) & 0x01010101u;
                *reinterpret_cast<uint2*>(s_state + (size_t)hd.x * CH + 8 * q) = outw;
                if (record) {
                    const int nvalid = max(0, min(8, g.n_chains - lchain));
                    const uint32_t vmask = (nvalid >= 8) ? 0xffu : ((1u << nvalid) - 1u);
                    const int ones = __popc(xbits & vmask);
                    if constexpr (RB) {  // Rao-Blackwell bins: the conditionals themselves (thresholds), not the draws
                        unsigned long long q0 = 0;
#pragma unroll
                        for (int i = 0; i < 8; i++)
                            if (i < nvalid) q0 += rb_units((uint64_t)T[i]);
                        if (nvalid > 0) {
                            atomicAdd(g.counts + hd.w, q0);
                            atomicAdd(g.counts + hd.w + 1, (unsigned long long)nvalid * 16777216ull - q0);
                        }
                    } else {
                        if (ones) atomicAdd(&s_counts[hd.w + 1], (unsigned)ones);
                        if (nvalid - ones) atomicAdd(&s_counts[hd.w], (unsigned)(nvalid - ones));
                    }
                    if (hist_half >= 0 && g.hist) {
                        if (s_hist)
                            add_row16(reinterpret_cast<uint4*>(s_hist + ((size_t)hist_half * t.n_order + c0 + j) * CH + 8 * q),
                                      spread_bits16(xbits & vmask));
                        else
                            hist_add8_binary(nullptr, g, m.total_card, hist_half, hd.w, CH, 8 * q, lchain, xbits & vmask, ~xbits & vmask);
                    }
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < m.n_vars * units; i += blockDim.x) {
        const int v = i / units, q = i - v * units;
        *reinterpret_cast<uint2*>(g.state + (size_t)v * g.n_pad + cta_chain + 8 * q) =
            *reinterpret_cast<const uint2*>(s_state + (size_t)v * CH + 8 * q);
    }
    if (record)
        for (int i = threadIdx.x; i < m.total_card; i += blockDim.x)
            if (s_counts[i]) atomicAdd(g.counts + i, (unsigned long long)s_counts[i]);
    if (s_hist && record) {
        // recorded sweeps of this launch that fell into each half window (the schedule of the sweep loop above)
        const int b0 = max(0, n_pre), e0 = min(n_sweeps, n_pre + n_half);
        const int n_rec[2] = {max(0, e0 - b0), max(0, n_sweeps - max(0, n_pre + n_half))};
        for (int i = threadIdx.x; i < 2 * t.n_order * units; i += blockDim.x) {  // (half, position, unit of 8 chains): 16-byte rows
            const int e = i >> unit_shift, u = i & (units - 1);
            const int half = e >= t.n_order ? 1 : 0, pos = e - half * t.n_order;
            const int nvalid = min(8, g.n_chains - (cta_chain + 8 * u));
            if (n_rec[half] == 0 || nvalid <= 0) continue;  // nothing recorded / padding
            if constexpr (MULTI) {
                if ((__ldg(t.trec + (size_t)pos * kTabRec + 2) >> 8) != 2) continue;  // non-binary: recorded in global memory directly
            }
            const uint4 ones = *reinterpret_cast<const uint4*>(s_hist + (size_t)e * CH + 8 * u);
            const uint4 lanes = spread_bits16(nvalid >= 8 ? 0xffu : ((1u << nvalid) - 1u));  // 1 in the lanes of existing chains
            const uint32_t nr = (uint32_t)n_rec[half];
            uint4 zeros;  // per 16-bit lane: recorded sweeps - ones (no borrow: ones <= recorded sweeps)
            zeros.x = lanes.x * nr - ones.x; zeros.y = lanes.y * nr - ones.y; zeros.z = lanes.z * nr - ones.z; zeros.w = lanes.w * nr - ones.w;
            uint16_t* h = g.hist + ((size_t)half * m.total_card + __ldg(t.trec + (size_t)pos * kTabRec + 3)) * g.n_pad + cta_chain + 8 * u;
            add_row16(reinterpret_cast<uint4*>(h), zeros);
            add_row16(reinterpret_cast<uint4*>(h + g.n_pad), ones);
        }
    }
}

// ------------------------------------------------------------------ K6
static __global__ void __launch_bounds__(256) k_init_state(const DevModel m, const DevGroup g) {
    const int32_t n_quads = g.n_pad >> 2;
    const int64_t total = (int64_t)m.n_vars * n_quads;
    for (int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; item < total;
         item += (int64_t)gridDim.x * blockDim.x) {
        const int32_t v = (int32_t)(item / n_quads);
        const int32_t q = (int32_t)(item - (int64_t)v * n_quads);
        const int32_t fx = __ldg(m.fixed + v);
        uint32_t packed;
        if (fx >= 0) {
            packed = (uint32_t)fx * 0x01010101u;
        } else {
            const uint32_t card = (uint32_t)__ldg(m.card + v);
            const uint32_t chain0 = (uint32_t)(g.first_chain + 4u * (uint32_t)q);
            const Philox4 a = philox4x32_10((uint32_t)v, 0u, chain0 >> 2, kTagInit, g.seed_lo, g.seed_hi);
            packed = __umulhi(a.x, card) | (__umulhi(a.y, card) << 8) | (__umulhi(a.z, card) << 16) |
                     (__umulhi(a.w, card) << 24);
        }
        *reinterpret_cast<uint32_t*>(g.state + (size_t)v * g.n_pad + 4 * q) = packed;
    }
}

// ------------------------------------------------------------------ K1-scan (parity mode)
// The reference's own schedule: every step each chain picks ONE variable uniformly among the
// free, un-collapsed ones (sampler.go:135-174) and updates it (gibbs-simple.go:163-271).  One
// thread per chain, float64, one Philox call per (chain, step): words (x) -> variable index by
// multiply-shift (the reference uses Int31n's rejection sampling; the bias here is < n/2^32),
// words (z, w) -> the 53-bit uniform of the value draw.  Slow by design (divergent gathers): it
// exists so that small-model parity runs can use the reference's schedule.
template <int MAXC>
__global__ void __launch_bounds__(128)
k_random_scan(const DevModel m, const DevGroup g, const int32_t* __restrict__ order, const int32_t n_order,
              const uint64_t step0, const int64_t n_steps, const int record) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= g.n_chains) return;
    const uint32_t chain = (uint32_t)(g.first_chain + (uint64_t)ch);
    for (int64_t s = 0; s < n_steps; s++) {
        const uint64_t step = step0 + (uint64_t)s;
        const Philox4 r = philox4x32_10((uint32_t)step, (uint32_t)(step >> 32), chain, kTagScan, g.seed_lo, g.seed_hi);
        const int v = order[__umulhi(r.x, (uint32_t)n_order)];
        const int card = m.card[v];
        double w[MAXC];
#pragma unroll
        for (int k = 0; k < MAXC; k++) w[k] = 0.0;
        const int32_t* p = m.prog + m.prog_off[v];
        const int nf = *p++;
        for (int f = 0; f < nf; f++) {
            const int tab_off = p[0], sv = p[1], no = p[2];
            p += 3;
            int b = tab_off;
            for (int o = 0; o < no; o++, p += 2) b += (int)g.state[(size_t)p[0] * g.n_pad + ch] * p[1];
#pragma unroll
            for (int k = 0; k < MAXC; k++)
                if (k < card) w[k] += m.tab64[b + k * sv];
        }
        stabilise_exp_floor<double, MAXC>(w, card);
        const int x = inverse_cdf<double, MAXC>(w, card, u53(r.z, r.w));
        g.state[(size_t)v * g.n_pad + ch] = (uint8_t)x;
        if (record) atomicAdd(g.counts + m.card_off[v] + x, 1ull);
    }
}

// ------------------------------------------------------------------ K5
// states: int32 [n_states][n_vars]; out: double [n_states][kOutStride] floored weights e[k]
constexpr int kProbeStride = 64;
template <typename Real>
__global__ void __launch_bounds__(128)
k_conditional(const DevModel m, const int32_t n_states, const int32_t* __restrict__ states,
              const int32_t* __restrict__ vars, double* __restrict__ out) {
    const Real* __restrict__ tab = tables_of<Real>(m);
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_states) return;
    const int32_t* st = states + (size_t)s * m.n_vars;
    const int v = vars[s];
    const int card = m.card[v];
    Real w[kProbeStride];
    for (int k = 0; k < kProbeStride; k++) w[k] = 0;
    const int32_t* p = m.prog + m.prog_off[v];
    const int nf = *p++;
    for (int f = 0; f < nf; f++) {
        const int tab_off = p[0], sv = p[1], no = p[2];
        p += 3;
        int b = tab_off;
        for (int o = 0; o < no; o++, p += 2) b += st[p[0]] * p[1];
        for (int k = 0; k < card; k++) w[k] += tab[b + k * sv];
    }
    stabilise_exp_floor<Real, kProbeStride>(w, card);
    for (int k = 0; k < kProbeStride; k++) out[(size_t)s * kProbeStride + k] = k < card ? (double)w[k] : 0.0;
}

// ------------------------------------------------------------------ K3
struct CollapsePlan {
    int32_t n_b;            // blanket size without the collapsed variable (<= kNeighborVarMaxDev)
    int32_t n_f;            // factors touching the collapsed variable
    int32_t card_v;
    int64_t new_size;       // entries of the new table
    int32_t bcard[kNeighborVarMaxDev];
    int32_t bfixed[kNeighborVarMaxDev];
    const int32_t* f_tab_off;   // [n_f]
    const int32_t* f_stride_v;  // [n_f]
    const int32_t* f_stride_b;  // [n_f][n_b] stride of blanket position b in factor f (0 if absent)
};

static __global__ void __launch_bounds__(256)
k_collapse(const CollapsePlan pl, const double* __restrict__ tab, double* __restrict__ new_tab,
           double* __restrict__ marg /*[card_v], pre-set to 1e-12*/) {
    __shared__ double s_marg[kMaxCardDev];
    for (int k = threadIdx.x; k < pl.card_v; k += blockDim.x) s_marg[k] = 0.0;
    __syncthreads();
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < pl.new_size) {
        int d[kNeighborVarMaxDev];
        int64_t rem = e;
        bool reachable = true;  // VariableIter(honorFixed) never visits other values of fixed vars
        for (int b = pl.n_b - 1; b >= 0; b--) {
            d[b] = (int)(rem % pl.bcard[b]);
            rem /= pl.bcard[b];
            if (pl.bfixed[b] >= 0 && d[b] != pl.bfixed[b]) reachable = false;
        }
        double acc = 0.0;
        if (reachable) {
            for (int x = 0; x < pl.card_v; x++) {
                double s = 0.0;
                for (int f = 0; f < pl.n_f; f++) {
                    int idx = pl.f_tab_off[f] + x * pl.f_stride_v[f];
                    for (int b = 0; b < pl.n_b; b++) idx += d[b] * pl.f_stride_b[f * pl.n_b + b];
                    s += tab[idx];
                }
                const double val = exp(s);
                acc += val;
                atomicAdd(&s_marg[x], val);
            }
        }
        // Function.UseLogSpace on the new factor (function.go:126-142)
        new_tab[e] = log(acc < 1e-6 ? acc + 1e-6 : acc);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < pl.card_v; k += blockDim.x)
        if (s_marg[k] != 0.0) atomicAdd(marg + k, s_marg[k]);
}

// ------------------------------------------------------------------ K4
// model/error.go:81-249 on two length-card vectors given as accessors
template <typename FA, typename FB>
__device__ __forceinline__ double measure_dev(int which, int card, FA A, FB B) {
    double t1 = 0.0, t2 = 0.0;
    for (int c = 0; c < card; c++) { t1 += A(c); t2 += B(c); }
    if (t1 < 1e-12) t1 = 1e-12;
    if (t2 < 1e-12) t2 = 1e-12;
    double acc = 0.0;
    if (which == 0) {
        for (int c = 0; c < card; c++) {
            const double e = fabs(A(c) / t1 - B(c) / t2);
            if (c == 0 || e > acc) acc = e;
        }
        return acc;
    } else if (which == 1) {
        for (int c = 0; c < card; c++) acc += fabs(A(c) / t1 - B(c) / t2);
        return acc / (double)card;
    } else if (which == 2) {
        for (int c = 0; c < card; c++) {
            const double dd = sqrt(A(c) / t1) - sqrt(B(c) / t2);
            acc += dd * dd;
        }
        return sqrt(acc) / sqrt(2.0);
    }
    double k1 = 0.0, k2 = 0.0;
    for (int c = 0; c < card; c++) {
        const double p1 = A(c) / t1, p2 = B(c) / t2, mid = (p1 + p2) * 0.5;
        const double qq = mid < 1e-12 ? 1e-12 : mid;
        const double x1 = p1 < 1e-12 ? 1e-12 : p1, x2 = p2 < 1e-12 ? 1e-12 : p2;
        k1 += x1 * log2(x1 / qq);
        k2 += x2 * log2(x2 / qq);
    }
    return 0.5 * (k1 + k2);
}

// item = (variable, chain): within = d(hist1, hist2), between = d(merged, hist1 + hist2), every
// histogram bin seeded with 1e-8 (chain.go:264-287); sums over chains into wb[v], wb[n_vars+v].
static __global__ void __launch_bounds__(256)
k_chain_dist(const DevModel m, const DevGroup g, const double* __restrict__ merged,
             const uint8_t* __restrict__ skip, const int which, double* __restrict__ wb, const double tail_chains) {
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(wb + 2 * m.n_vars, tail_chains);  // chains behind these sums (summed over ranks with them)
    const int64_t total = (int64_t)m.n_vars * g.n_chains;
    for (int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; item < total;
         item += (int64_t)gridDim.x * blockDim.x) {
        const int v = (int)(item / g.n_chains);
        const int ch = (int)(item - (int64_t)v * g.n_chains);
        if (skip[v]) continue;
        const int card = m.card[v], coff = m.card_off[v];
        const uint16_t* h1 = g.hist + (size_t)coff * g.n_pad + ch;
        const uint16_t* h2 = g.hist + ((size_t)m.total_card + coff) * g.n_pad + ch;
        const size_t st = (size_t)g.n_pad;
        auto A1 = [&](int c) { return 1e-8 + (double)h1[c * st]; };
        auto A2 = [&](int c) { return 1e-8 + (double)h2[c * st]; };
        auto A12 = [&](int c) { return (1e-8 + (double)h1[c * st]) + (1e-8 + (double)h2[c * st]); };
        auto M = [&](int c) { return merged[coff + c]; };
        const double within = measure_dev(which, card, A1, A2);
        const double between = measure_dev(which, card, M, A12);
        atomicAdd(wb + v, within);
        atomicAdd(wb + m.n_vars + v, between);
    }
}

// this device's contribution to MergeChains (chain.go:131-144): every chain starts at the
// uniform marginal 1/card (model/variable.go:45) and adds its counts
static __global__ void __launch_bounds__(256)
k_merge_partial(const DevModel m, const unsigned long long* __restrict__ counts, const double n_chains,
                const uint8_t* __restrict__ skip, double* __restrict__ out, const double count_unit) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.total_card) return;
    const int v = m.entry_var[i];
    if (skip[v] & 1) return;
    out[i] += n_chains * (1.0 / (double)m.card[v]) + (double)counts[i] * count_unit;  // 1, or 2^-24 for the Rao-Blackwell bins
}

// MergeChains, integer form (multi-GPU and asynchronous path): the sample counts of this device's groups are summed as
// 64-bit integers (collapsed-in-any variables stay 0); the tail carries this device's chain count and TotalSampleCount.
// After the (optional) NCCL sum over ranks, k_merge_finalize adds the chains' uniform start mass (model/variable.go:45)
// in one rounding per entry, so the merged marginals are bit-identical however the chains are sharded over devices.
static __global__ void __launch_bounds__(256)
k_merge_counts(const DevModel m, const unsigned long long* __restrict__ counts, const uint8_t* __restrict__ skip,
               unsigned long long* __restrict__ sum, const unsigned long long tail_chains, const unsigned long long tail_samples) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        sum[m.total_card] += tail_chains;
        sum[m.total_card + 1] += tail_samples;
    }
    if (i >= m.total_card) return;
    if (skip[m.entry_var[i]] & 1) return;
    sum[i] += counts[i];
}
static __global__ void __launch_bounds__(256)
k_merge_finalize(const DevModel m, const unsigned long long* __restrict__ sum, const uint8_t* __restrict__ skip,
                 double* __restrict__ out, const double count_unit) {
    const double n_chains = (double)sum[m.total_card];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m.total_card; i += gridDim.x * blockDim.x) {
        const int v = m.entry_var[i];
        out[i] = (skip[v] & 1) ? 0.0 : n_chains * (1.0 / (double)m.card[v]) + (double)sum[i] * count_unit;
    }
}

}  // namespace gb
