// Minibatch permutation for the on-device epoch scheduler (the reference has no minibatches -- Chief.py:58-64
// feeds whole rollouts -- so the scheme is this repo's; the parity tests hold a numpy restatement of it).
//
// A keyed bijection of [0, n): a 4-round Feistel network over ceil(log2 n) bits (unbalanced halves when the bit count is
// odd) with cycle walking for n that is not a power of two.  Index i of epoch e is computed in O(1) by whichever thread
// needs it: no sort, no permutation array, no host round trip.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define DPPO_HD __host__ __device__ __forceinline__
#else
#define DPPO_HD inline
#endif

namespace dppo {

struct PermKey {
    uint32_t k[4];
    uint32_t n;      // domain size
    int bits;        // max(2, ceil(log2 n))
};

DPPO_HD uint32_t perm_mix(uint32_t x, uint32_t key) {
    x ^= key; x *= 0x9E3779B1u; x ^= x >> 15; x *= 0x85EBCA77u; x ^= x >> 13; x *= 0xC2B2AE3Du; x ^= x >> 16;
    return x;
}

DPPO_HD uint32_t perm_index(const PermKey& p, uint32_t i) {
    const int a0 = p.bits / 2, b0 = p.bits - a0;      // left half a0 bits, right half b0 bits
    do {
        int la = a0, lb = b0;
        uint32_t L = i >> b0, R = i & ((1u << b0) - 1u);
        #pragma unroll
        for (int r = 0; r < 4; ++r) {
            const uint32_t nl = R, nr = L ^ (perm_mix(R, p.k[r]) & ((1u << la) - 1u));
            L = nl; R = nr;
            const int t = la; la = lb; lb = t;
        }
        i = (L << lb) | R;                              // after 4 rounds la,lb are back to a0,b0
    } while (i >= p.n);
    return i;
}

inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

inline PermKey make_perm_key(uint32_t n, uint64_t seed, uint64_t epoch) {
    PermKey p;
    p.n = n;
    int bits = 2;
    while (bits < 31 && (1u << bits) < n) ++bits;
    p.bits = bits;
    uint64_t s = splitmix64(seed ^ (epoch * 0xD1B54A32D192ED03ull));
    for (int r = 0; r < 4; ++r) { s = splitmix64(s); p.k[r] = (uint32_t)(s >> 32); }
    return p;
}

}  // namespace dppo
