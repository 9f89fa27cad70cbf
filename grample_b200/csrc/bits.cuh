// Bit-sliced table-mode sweep (GB_TABLE_BITS): the same update as k_sweep_tab — the reference's float64 conditional
// (sampler/gibbs-simple.go:171-258) evaluated once per (variable, neighbour configuration), stored as a 32-bit
// inverse-CDF threshold T, value 0 drawn iff a uniform 32-bit draw u <= T (sampler/sampler.go:115-123) — for models
// whose sampled variables are binary with at most 4 free binary neighbours (the Ising / Grids problems), with the
// chain state packed ONE BIT PER CHAIN:
//
//   bits[var][chain >> 5], chain = bit (chain & 31)          (1 M variables x 65536 chains = 8 GiB instead of 64)
//
// A thread owns W 32-bit words (W x 32 chains) of every variable of its tile, so a warp's load of one neighbour is one
// coalesced 128-byte line per word and nothing is unpacked: the whole update is done on 32 chains at a time with
// bitwise logic.
//
//   * the four neighbour words n0..n3 ARE the configuration index, bit-sliced (cfg = n0 + 2 n1 + 4 n2 + 8 n3);
//   * the draw is bit-sliced too: one Philox4x32-10 call yields four 32-bit words = four bit PLANES of the draws of
//     32 chains (plane b holds bit b of every chain's draw), so two calls give the top 8 bits of 32 draws;
//   * for plane b the threshold bit of every chain, Tsel_b = bit b of T[cfg], is a 4-input boolean function of
//     (n0, n1, n2, n3) whose truth table is bit b of the variable's 16 thresholds: a 16:1 multiplexer tree — level 0
//     on the FMA pipe (r_k = n0 * a_k + b_k with a_k in {0, 1, -1}, b_k in {0, -1} selects one of 0, ~0, n0, ~n0),
//     levels 1..3 as 3-input LOP3 multiplexers on n1, n2, n3: 8 IMAD + 7 LOP3 per plane per 32 chains; the
//     coefficients are warp-uniform and come from shared memory (built per tile from the thresholds);
//   * the comparison runs MSB-first on the planes: gt |= eq & d & ~Tsel, eq &= ~(d ^ Tsel) — 2 LOP3 per plane per 32
//     chains.  After 8 planes a chain is undecided (eq) with probability 2^-8; those are pushed to a shared-memory
//     queue and resolved after the tile by whichever threads are free, one item per thread, with the remaining 24
//     bits of the draw and the full 32-bit threshold — exactly the comparison u > T on the 32-bit draw
//     u = (top 8 bits from the planes) << 24 | (24 tie bits), so the law is the same as k_sweep_tab's.
//
// Stream (restated by oracle/sweep.hpp, bits = 33): key = seed, chain word w = global chain >> 5, position p = chain & 31
//   planes 7..4 = words x, y, z, w of Philox(var, sweep, w, kTagPlaneA); planes 3..0 of Philox(var, sweep, w, kTagPlaneB);
//   tie bits of chain p = word (p & 3) of Philox(var, sweep, w, kTagTie24 | (p >> 2) << 8), shifted right by 8.
#pragma once
#include "kernels.cuh"

namespace gb {

constexpr int kBitsVB = 32;     // sweep positions per tile (half of a locality-ordered 64-position patch)
// deferred ties per tile: 6 per thread-word (expected kBitsVB * (1 - (255 / 256)^32) = 3.77, sd 0.08 over a tile; overflow resolves inline)
constexpr int kBitsRec = 8;     // {v, card_off, thr_off, cfg mask, nbr[4]}

// initial state, identical to k_init_state's values (Philox kTagInit: chain >> 2 per call, umulhi(word, 2))
static __global__ void __launch_bounds__(256) k_init_bits(const DevModel m, const DevGroup g, uint32_t* __restrict__ bits, const int32_t n_words) {
    const int64_t total = (int64_t)m.n_vars * n_words;
    for (int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; item < total; item += (int64_t)gridDim.x * blockDim.x) {
        const int32_t v = (int32_t)(item / n_words);
        const int32_t w = (int32_t)(item - (int64_t)v * n_words);
        const int32_t fx = __ldg(m.fixed + v);
        uint32_t word = 0;
        if (fx >= 0) {
            word = fx ? 0xffffffffu : 0u;
        } else {
            const uint32_t blk0 = (uint32_t)((g.first_chain >> 2) + 8ull * (uint64_t)w);
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const Philox4 a = philox4x32_10((uint32_t)v, 0u, blk0 + (uint32_t)q, kTagInit, g.seed_lo, g.seed_hi);
                word |= (a.x >> 31) << (4 * q) | (a.y >> 31) << (4 * q + 1) | (a.z >> 31) << (4 * q + 2) | (a.w >> 31) << (4 * q + 3);  // umulhi(word, 2)
            }
        }
        bits[(size_t)v * n_words + w] = word;
    }
}

// Philox4x32-10 with the ten round keys (key + r * Weyl constants) precomputed on the host and passed as a kernel
// parameter: they sit in the constant bank and feed the round's 3-input XOR directly, so the key schedule costs no
// instruction (left to the compiler it is either hoisted into 20 registers or recomputed every call).
struct PhiloxKeys {
    uint32_t k0[10], k1[10];
};
inline PhiloxKeys philox_keys(const uint32_t seed_lo, const uint32_t seed_hi) {
    PhiloxKeys k;
    for (int r = 0; r < 10; r++) {
        k.k0[r] = seed_lo + (uint32_t)r * 0x9E3779B9u;
        k.k1[r] = seed_hi + (uint32_t)r * 0xBB67AE85u;
    }
    return k;
}
__device__ __forceinline__ Philox4 philox_keyed(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys& k) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k.k0[r], n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k.k1[r];
        c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
    }
    return Philox4{c0, c1, c2, c3};
}

__device__ __forceinline__ uint32_t mux32(const uint32_t sel, const uint32_t hi, const uint32_t lo) { return (sel & hi) | (~sel & lo); }

// one undecided word of one variable: decide every chain in `eq` with the tie bits; returns the bits to set
__device__ __forceinline__ uint32_t bits_resolve(const DevTab& t, const int32_t* __restrict__ rec, const uint32_t* __restrict__ bits,
                                                const int32_t n_words, const int32_t wi, const uint32_t gw, const uint32_t sweep,
                                                const uint32_t seed_lo, const uint32_t seed_hi, uint32_t eq) {
    const uint32_t v = (uint32_t)rec[0];
    const uint32_t n0 = __ldg(bits + (size_t)rec[4] * n_words + wi), n1 = __ldg(bits + (size_t)rec[5] * n_words + wi);
    const uint32_t n2 = __ldg(bits + (size_t)rec[6] * n_words + wi), n3 = __ldg(bits + (size_t)rec[7] * n_words + wi);
    uint32_t add = 0;
    while (eq) {
        const int p = __ffs(eq) - 1;
        eq &= eq - 1;
        const uint32_t cfg = (((n0 >> p) & 1u) | ((n1 >> p) & 1u) << 1 | ((n2 >> p) & 1u) << 2 | ((n3 >> p) & 1u) << 3) & (uint32_t)rec[3];
        const uint32_t T = __ldg(t.thr + rec[2] + cfg);
        const Philox4 r = philox4x32_10(v, sweep, gw, kTagTie24 | ((uint32_t)(p >> 2) << 8), seed_lo, seed_hi);
        const uint32_t word = (p & 2) ? ((p & 1) ? r.w : r.z) : ((p & 1) ? r.y : r.x);
        if ((word >> 8) > (T & 0x00ffffffu)) add |= 1u << p;
    }
    return add;
}

// CTA tile = (chunk of 256 * W consecutive state words = 8192 * W chains) x (kBitsVB consecutive sweep positions of the
// colour); tiles are handed out by an atomic counter (`tile_counter[0]`, zero at launch; `tile_counter[1]` counts the
// CTAs that have left), so a CTA that starts late — e.g. behind the NCCL kernel of an overlapped merge — simply takes
// fewer tiles.
// W state words per thread, NT threads per CTA.  S = 1: a chunk is NT * W words (32 NT W chains) and every thread walks all
// positions of the tile; S = 2: the two halves of the CTA take the two halves of the tile's positions for a chunk of
// NT * W / 2 words — the shape for populations of one such chunk (8192 chains per GPU = the 8-way split of 65536), which
// keeps 256-thread CTAs and their register allocation instead of a separately compiled 128-thread kernel.
template <int W, int NT, int S = 1>
__global__ void __launch_bounds__(NT, NT == 128 ? (W == 2 ? 6 : 8) : 0)  // 256 threads: the compiler's own choice (80 / 48 registers); 128: the same warps per SM
k_sweep_bits(const DevModel m, const DevTab t, const DevGroup g, uint32_t* __restrict__ bits, const int32_t n_words,
             const int32_t j_begin, const int32_t n_vars_c, const uint32_t sweep, const int record,
             unsigned int* __restrict__ tile_counter, const PhiloxKeys keys) {
    constexpr int VB = kBitsVB;
    __shared__ __align__(16) int2 s_coef[VB * 8 * 8];   // [position][plane 7..0][pair k]: {a_k, b_k}
    __shared__ __align__(16) int32_t s_rec[VB * kBitsRec];
    __shared__ __align__(16) const uint32_t* s_row[VB * 4];         // state rows of the 4 neighbours of every position
    __shared__ uint8_t s_t8[VB * 16];                    // top byte of the 16 thresholds of every position
    __shared__ unsigned int s_cnt[VB];                   // ones per position over the tile's chains
    constexpr int kBitsQueue = 6 * NT * W;
    __shared__ uint2 s_q[kBitsQueue];                    // deferred ties: {position << 16 | word slot, eq}
    __shared__ unsigned int s_qn;
    __shared__ int s_tile;
    constexpr int NTS = NT / S;        // threads that share a position
    const int chunk_words = NTS * W;
    const int chunks = (n_words + chunk_words - 1) / chunk_words;
    const int n_vb = (n_vars_c + VB - 1) / VB;
    const int64_t n_tiles = (int64_t)chunks * n_vb;
    const uint32_t seed_lo = g.seed_lo, seed_hi = g.seed_hi;
    const uint32_t gw0 = (uint32_t)(g.first_chain >> 5);
    const int tid = threadIdx.x;
    const int lane_t = tid % NTS, part = tid / NTS;  // word slot within the chunk; which part of the tile's positions
    for (;;) {
        if (tid == 0) s_tile = (int)atomicAdd(tile_counter, 1u);
        __syncthreads();
        const int64_t tile = s_tile;
        if (tile >= n_tiles) {
            // the last CTA to leave re-arms the group's counter pair {next tile, CTAs done} for the group's next launch
            if (tid == 0 && atomicAdd(tile_counter + 1, 1u) == gridDim.x - 1) {
                tile_counter[0] = 0u;
                tile_counter[1] = 0u;
            }
            break;
        }
        const int chunk = (int)(tile / n_vb), vb = (int)(tile - (int64_t)chunk * n_vb);
        const int nv = min(VB, n_vars_c - vb * VB);
        const int j0 = j_begin + vb * VB;
        // ---- stage the tile: records, threshold top bytes, multiplexer coefficients
        for (int i = tid; i < nv * kBitsRec; i += NT) {
            const int j = i >> 3, f = i & 7;
            const int32_t* r = t.trec + (size_t)(j0 + j) * kTabRec;  // {v, thr_off, n_nbr, card_off, nbr[8], stride[8]}
            int32_t val;
            if (f == 0) val = __ldg(r);
            else if (f == 1) val = __ldg(r + 3);
            else if (f == 2) val = __ldg(r + 1);
            else if (f == 3) val = (1 << (__ldg(r + 2) & 0xff)) - 1;
            else val = __ldg(r + f);  // nbr[f - 4]; slots past n_nbr hold the variable itself (its bits are masked out of cfg)
            s_rec[i] = val;
        }
        if (tid < VB) s_cnt[tid] = 0;
        if (tid == 0) s_qn = 0;
        __syncthreads();
        if (tid < nv * 4) s_row[tid] = bits + (size_t)s_rec[(tid >> 2) * kBitsRec + 4 + (tid & 3)] * n_words;
        for (int i = tid; i < nv * 16; i += NT) {
            const int j = i >> 4, c = i & 15;
            s_t8[i] = (uint8_t)(__ldg(t.thr + s_rec[j * kBitsRec + 2] + (c & s_rec[j * kBitsRec + 3])) >> 24);
        }
        __syncthreads();
        for (int q = tid; q < nv * 64; q += NT) {  // pair q = (position, plane slot, k): plane slot 0 = bit 7
            const int j = q >> 6, b = 7 - ((q >> 3) & 7), k = q & 7;
            const int t0 = (s_t8[j * 16 + 2 * k] >> b) & 1, t1 = (s_t8[j * 16 + 2 * k + 1] >> b) & 1;
            s_coef[q] = make_int2(t1 - t0, -t0);  // n0 * a + b = {0, ~0, n0, ~n0} for (t0, t1) = {00, 11, 01, 10}
        }
        __syncthreads();

        // ---- this thread's words of the chunk
        int32_t wi[W];
        uint32_t valid[W];
#pragma unroll
        for (int u = 0; u < W; u++) {
            wi[u] = chunk * chunk_words + u * NTS + lane_t;
            const int64_t first = 32ll * wi[u];
            const int64_t left = (int64_t)g.n_chains - first;
            valid[u] = wi[u] >= n_words || left <= 0 ? 0u : (left >= 32 ? 0xffffffffu : ((1u << (int)left) - 1u));
            if (wi[u] >= n_words) wi[u] = n_words - 1;  // padding threads redo the last word, stores masked off
        }
        const bool any_word = chunk * chunk_words + (lane_t & ~31) < n_words;  // the warp owns at least one real word
        const int j_lo = min(nv, part * (VB / S)), j_hi = min(nv, (part + 1) * (VB / S));  // this thread's positions of the tile
        if (any_word) {
            // neighbour words of the NEXT position are in flight while this one computes (rows of the other colour:
            // read-only for the whole launch)
            uint32_t nx[W][4];
            auto load_nbrs = [&](const int j) {
                const ulonglong2 lo = *reinterpret_cast<const ulonglong2*>(&s_row[j * 4]), hi = *reinterpret_cast<const ulonglong2*>(&s_row[j * 4 + 2]);
                const uint32_t* r0 = reinterpret_cast<const uint32_t*>(lo.x);
                const uint32_t* r1 = reinterpret_cast<const uint32_t*>(lo.y);
                const uint32_t* r2 = reinterpret_cast<const uint32_t*>(hi.x);
                const uint32_t* r3 = reinterpret_cast<const uint32_t*>(hi.y);
#pragma unroll
                for (int u = 0; u < W; u++) {
                    nx[u][0] = __ldg(r0 + wi[u]);
                    nx[u][1] = __ldg(r1 + wi[u]);
                    nx[u][2] = __ldg(r2 + wi[u]);
                    nx[u][3] = __ldg(r3 + wi[u]);
                }
            };
            load_nbrs(min(j_lo, nv - 1));
            for (int jg = j_lo; jg < j_hi; jg += 2) {
                uint32_t acc = 0;  // ones of positions jg (low half) and jg + 1 (high half) over this thread's chains
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int j = jg + h;
                    if (j < j_hi) {
                        const int4 ra = *reinterpret_cast<const int4*>(&s_rec[j * kBitsRec]);      // v, card_off, thr_off, cfg mask
                        uint32_t n[W][4], gt[W], eq[W];
#pragma unroll
                        for (int u = 0; u < W; u++) {
#pragma unroll
                            for (int i = 0; i < 4; i++) n[u][i] = nx[u][i];
                            gt[u] = 0u;
                            eq[u] = 0xffffffffu;
                        }
                        load_nbrs(min(j + 1, j_hi - 1));
                        uint32_t* const own = bits + (size_t)ra.x * n_words;
#pragma unroll
                        for (int half = 0; half < 2; half++) {
                            uint32_t d[W][4];
#pragma unroll
                            for (int u = 0; u < W; u++) {
                                const Philox4 r = philox_keyed((uint32_t)ra.x, sweep, gw0 + (uint32_t)wi[u], half ? kTagPlaneB : kTagPlaneA, keys);
                                d[u][0] = r.x; d[u][1] = r.y; d[u][2] = r.z; d[u][3] = r.w;
                            }
#pragma unroll
                            for (int pb = 0; pb < 4; pb++) {
                                const int4* cf = reinterpret_cast<const int4*>(&s_coef[(j * 8 + half * 4 + pb) * 8]);
                                const int4 c0 = cf[0], c1 = cf[1], c2 = cf[2], c3 = cf[3];
#pragma unroll
                                for (int u = 0; u < W; u++) {
                                    const uint32_t x0 = n[u][0];
                                    const uint32_t r0 = x0 * (uint32_t)c0.x + (uint32_t)c0.y, r1 = x0 * (uint32_t)c0.z + (uint32_t)c0.w;
                                    const uint32_t r2 = x0 * (uint32_t)c1.x + (uint32_t)c1.y, r3 = x0 * (uint32_t)c1.z + (uint32_t)c1.w;
                                    const uint32_t r4 = x0 * (uint32_t)c2.x + (uint32_t)c2.y, r5 = x0 * (uint32_t)c2.z + (uint32_t)c2.w;
                                    const uint32_t r6 = x0 * (uint32_t)c3.x + (uint32_t)c3.y, r7 = x0 * (uint32_t)c3.z + (uint32_t)c3.w;
                                    const uint32_t s0 = mux32(n[u][1], r1, r0), s1 = mux32(n[u][1], r3, r2);
                                    const uint32_t s2 = mux32(n[u][1], r5, r4), s3 = mux32(n[u][1], r7, r6);
                                    const uint32_t T = mux32(n[u][3], mux32(n[u][2], s3, s2), mux32(n[u][2], s1, s0));
                                    const uint32_t dd = d[u][pb];
                                    gt[u] |= eq[u] & dd & ~T;
                                    eq[u] &= ~(dd ^ T);
                                }
                            }
                        }
#pragma unroll
                        for (int u = 0; u < W; u++) {
                            if (valid[u]) own[wi[u]] = gt[u];
                            acc += (uint32_t)__popc(gt[u] & valid[u]) << (16 * h);
                            if (eq[u] & valid[u]) {  // undecided after 8 planes (2^-8 per chain): resolve after the tile
                                const unsigned slot = atomicAdd(&s_qn, 1u);
                                if (slot < (unsigned)kBitsQueue) {
                                    s_q[slot] = make_uint2((uint32_t)j << 16 | (uint32_t)(u * NTS + lane_t), eq[u] & valid[u]);
                                } else {
                                    const uint32_t add = bits_resolve(t, &s_rec[j * kBitsRec], bits, n_words, wi[u], gw0 + (uint32_t)wi[u], sweep,
                                                                      seed_lo, seed_hi, eq[u] & valid[u]);
                                    if (add) {
                                        own[wi[u]] = gt[u] | add;
                                        acc += (uint32_t)__popc(add) << (16 * h);
                                    }
                                }
                            }
                        }
                    }
                }
                if (record) {  // chain.go:231-236: two positions per warp reduction (16-bit fields, <= 32 * 32 * W each)
                    const unsigned s = __reduce_add_sync(0xffffffffu, acc);
                    if ((tid & 31) == 0) {
                        if (s & 0xffffu) atomicAdd(&s_cnt[jg], s & 0xffffu);
                        if (s >> 16) atomicAdd(&s_cnt[jg + 1], s >> 16);
                    }
                }
            }
        }
        __syncthreads();
        // ---- deferred ties: one queued word per thread
        const int nq = (int)min(s_qn, (unsigned)kBitsQueue);
        for (int i = tid; i < nq; i += NT) {
            const uint2 it = s_q[i];
            const int j = (int)(it.x >> 16), slot = (int)(it.x & 0xffffu);
            const int32_t w = chunk * chunk_words + slot;
            const uint32_t add = bits_resolve(t, &s_rec[j * kBitsRec], bits, n_words, w, gw0 + (uint32_t)w, sweep, seed_lo, seed_hi, it.y);
            if (add) {
                atomicOr(bits + (size_t)s_rec[j * kBitsRec] * n_words + w, add);
                if (record) atomicAdd(&s_cnt[j], (unsigned)__popc(add));
            }
        }
        __syncthreads();
        if (record && tid < nv) {
            const int32_t coff = s_rec[tid * kBitsRec + 1];
            const unsigned o = s_cnt[tid];
            const int64_t first = 32ll * chunk * chunk_words;
            const int valid_chains = (int)max((int64_t)0, min((int64_t)32 * chunk_words, (int64_t)g.n_chains - first));
            if (o) atomicAdd(g.counts + coff + 1, (unsigned long long)o);
            if (valid_chains - (int)o) atomicAdd(g.counts + coff, (unsigned long long)(valid_chains - (int)o));
        }
        __syncthreads();
    }
}

}  // namespace gb
