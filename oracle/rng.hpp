// ORACLE — TEST INFRASTRUCTURE ONLY (see model.hpp header).
//
// Random sources for the CPU restatement.
//  * MT19937-64: the reference draws from github.com/seehuhn/mt19937 v1.0.0 (go.mod:7),
//    whose source is NOT under /root/reference.  It is the standard Matsumoto/Nishimura
//    MT19937-64; restated from the published algorithm and pinned by the reference's
//    golden vector rand/rand_test.go:17-39 (init_by_array64 {0x12345,0x23456,0x34567,0x45678},
//    outputs masked to 63 bits).  The Seed(int64) path (rand/rand.go:27-28) has no
//    known-answer test in the reference: standard init_genrand64 is assumed.
//  * Generator: rand/rand.go:52-105 (Int63/Int63n/Int31/Int31n/Float64, copies of Go stdlib).
//    The reference serves draws through a 1024-deep channel from one goroutine; the draw
//    ORDER within one consumer is identical, the channel itself is not restated.
//  * Philox4x32-10: the device sampler's counter-based stream, restated so that the
//    sweep-mode oracle can reproduce the device trajectory draw for draw.  Pinned by the
//    Random123 known-answer vectors (tests/test_oracle_rng.py).
#pragma once
#include <cstdint>
#include <vector>

#include "model.hpp"

namespace oracle {

struct MT19937_64 {
    static constexpr int NN = 312, MM = 156;
    static constexpr uint64_t MATRIX_A = 0xB5026F5AA96619E9ULL;
    static constexpr uint64_t UM = 0xFFFFFFFF80000000ULL, LM = 0x7FFFFFFFULL;
    uint64_t mt[NN];
    int mti = NN + 1;

    void seed(uint64_t s) {  // init_genrand64
        mt[0] = s;
        for (mti = 1; mti < NN; mti++)
            mt[mti] = 6364136223846793005ULL * (mt[mti - 1] ^ (mt[mti - 1] >> 62)) + (uint64_t)mti;
    }
    void seed_from_slice(const std::vector<uint64_t>& key) {  // init_by_array64
        seed(19650218ULL);
        uint64_t i = 1, j = 0;
        uint64_t klen = key.size();
        uint64_t k = NN > klen ? NN : klen;
        for (; k; k--) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 62)) * 3935559000370003845ULL)) + key[j] + j;
            i++; j++;
            if (i >= NN) { mt[0] = mt[NN - 1]; i = 1; }
            if (j >= klen) j = 0;
        }
        for (k = NN - 1; k; k--) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 62)) * 2862933555777941757ULL)) - i;
            i++;
            if (i >= NN) { mt[0] = mt[NN - 1]; i = 1; }
        }
        mt[0] = 1ULL << 63;
    }
    uint64_t next_u64() {
        static const uint64_t mag01[2] = {0ULL, MATRIX_A};
        if (mti >= NN) {
            if (mti == NN + 1) seed(5489ULL);
            int i;
            uint64_t x;
            for (i = 0; i < NN - MM; i++) {
                x = (mt[i] & UM) | (mt[i + 1] & LM);
                mt[i] = mt[i + MM] ^ (x >> 1) ^ mag01[(int)(x & 1ULL)];
            }
            for (; i < NN - 1; i++) {
                x = (mt[i] & UM) | (mt[i + 1] & LM);
                mt[i] = mt[i + (MM - NN)] ^ (x >> 1) ^ mag01[(int)(x & 1ULL)];
            }
            x = (mt[NN - 1] & UM) | (mt[0] & LM);
            mt[NN - 1] = mt[MM - 1] ^ (x >> 1) ^ mag01[(int)(x & 1ULL)];
            mti = 0;
        }
        uint64_t x = mt[mti++];
        x ^= (x >> 29) & 0x5555555555555555ULL;
        x ^= (x << 17) & 0x71D67FFFEDA60000ULL;
        x ^= (x << 37) & 0xFFF7EEE000000000ULL;
        x ^= (x >> 43);
        return x;
    }
};

// rand/rand.go:12-105.  Virtual so the sweep-mode oracle can substitute a Philox source.
struct Generator {
    MT19937_64 mt;
    virtual ~Generator() = default;
    Generator() { mt.seed(5489ULL); }
    explicit Generator(int64_t seed) { mt.seed((uint64_t)seed); }  // rand.go:47-49
    explicit Generator(const std::vector<uint64_t>& seed) {        // rand.go:19-44
        if (seed.empty()) throw Error("Invalid generator seed array");
        if (seed.size() == 1) mt.seed(seed[0]);
        else mt.seed_from_slice(seed);
    }
    virtual int64_t int63() { return (int64_t)(mt.next_u64() & 0x7fffffffffffffffULL); }  // rand.go:52 + rand_test.go:33
    int64_t int63n(int64_t n) {  // rand.go:57-73
        if (n <= 0) throw Error("invalid argument to Int63n");
        if ((n & (n - 1)) == 0) return int63() & (n - 1);
        int64_t max = (int64_t)((1ULL << 63) - 1 - (1ULL << 63) % (uint64_t)n);
        int64_t v = int63();
        while (v > max) v = int63();
        return v % n;
    }
    int32_t int31() { return (int32_t)(int63() >> 32); }  // rand.go:76-78
    int32_t int31n(int32_t n) {                            // rand.go:81-98
        if (n <= 0) throw Error("invalid argument to Int31n");
        if ((n & (n - 1)) == 0) return int31() & (n - 1);
        int32_t max = (int32_t)((1U << 31) - 1 - (1U << 31) % (uint32_t)n);
        int32_t v = int31();
        while (v > max) v = int31();
        return v % n;
    }
    virtual double float64() { return (double)int63n(1LL << 53) / (double)(1LL << 53); }  // rand.go:102-105
};

// ---------------------------------------------------------------- Philox4x32-10
struct Philox4x32 {
    static inline void round(uint32_t c[4], const uint32_t k[2]) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    static inline void gen(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
        uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
        uint32_t k[2] = {key[0], key[1]};
        for (int r = 0; r < 10; r++) {
            round(c, k);
            k[0] += 0x9E3779B9u;
            k[1] += 0xBB67AE85u;
        }
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
    }
};

}  // namespace oracle
