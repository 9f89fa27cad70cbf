// Counter-based Philox4x32-10 stream of the device sampler (replaces the reference's shared
// MT19937-64 channel, rand/rand.go:12-105, as BASELINE.json's north_star prescribes).
//
// Stream definition (the oracle restates it in oracle/sweep.hpp — keep in sync):
//   key     = (seed_lo, seed_hi)
//   counter = (var, sweep, chain_block, tag)
//   tag kTagDraw24: chain_block = chain >> 2, the 4 output words serve chains 4b .. 4b+3,
//                   U = (word >> 8) * 2^-24 (float32 sweeps; statistical parity only)
//   tag kTagDraw16Hi / kTagDraw16Lo: chain_block = chain >> 3; chain 8b+i takes the 16-bit field
//                   (word[i >> 1] >> 16*(i & 1)) & 0xffff of each call; the 32-bit draw is
//                   word32 = (hi16 << 16) | lo16, U = word32 * 2^-32.  The table-mode sweep only
//                   evaluates the Lo call when hi16 alone does not decide the comparison.
//   tags kTagPlaneA / kTagPlaneB / kTagTie24 (bit-sliced table sweep, bits.cuh): chain_block = chain >> 5 (one 32-chain
//                   state word), p = chain & 31.  The 32-bit draw of chain p is u = hi8 << 24 | lo24 with
//                   bit 7..4 of hi8 = bit p of words x, y, z, w of the PlaneA call, bit 3..0 = bit p of the PlaneB
//                   call's words, lo24 = word (p & 3) of the call with tag kTagTie24 | (p >> 2) << 8, shifted right by 8
//                   (only evaluated when hi8 equals the threshold's top byte).
//   tag kTagDraw53: chain_block = chain >> 1, words (0,1) -> chain 2b, (2,3) -> chain 2b+1,
//                   x = ((w_a << 32) | w_b) >> 11, U = x * 2^-53   (same range as Go's Float64)
//   tag kTagInit  : chain_block = chain >> 2, value = (word * card) >> 32
// `chain` is the GLOBAL chain id, so a trajectory does not depend on how chains are sharded.
#pragma once
#include <cstdint>

namespace gb {

enum : uint32_t { kTagDraw24 = 1, kTagDraw53 = 2, kTagInit = 3, kTagScan = 4, kTagCollapse = 5, kTagDraw16Hi = 6, kTagDraw16Lo = 7,
                  kTagPlaneA = 8, kTagPlaneB = 9, kTagTie24 = 10 };

struct Philox4 {
    uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                           uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
#ifdef __CUDA_ARCH__
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return Philox4{c0, c1, c2, c3};
}

__host__ __device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
    const uint64_t x = (((uint64_t)a << 32) | (uint64_t)b) >> 11;
    return (double)x * (1.0 / 9007199254740992.0);
}

}  // namespace gb
