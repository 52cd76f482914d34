// csrc/kmc_philox.cuh -- counter-based Philox4x32-10 stream keyed by (molecule, partner, step, slot).
// Replaces rand2() (main.cpp:2313-2326: a freshly time-seeded mt19937_64 per call, 89 % of the
// reference's run time) and the std::rand() behind random_shuffle (main.cpp:1285/1345/1413/1597).
// Only the distribution (iid U[0,1), 53 bits) is kept from the reference; the keying makes every draw
// independent of evaluation order, which is what lets the sweep run in parallel and still replay.
#pragma once
#include <stdint.h>

namespace kmc {

enum DrawSlot : uint32_t {
    SLOT_MOVE0 = 0,          // 0..5 unit move draws   main.cpp:585,587,611 | 693,695,726 | 909-911,942-944 | 990,991,1089
    SLOT_RL_ON = 8,          // main.cpp:1919   molecule = receptor i, partner = 4*j + k
    SLOT_MONO_CIS_ON = 9,    // main.cpp:1985   molecule = i, partner = j
    SLOT_CIS_ON = 10,        // main.cpp:2039
    SLOT_RL_OFF = 11,        // main.cpp:2070
    SLOT_MONO_CIS_OFF = 12,  // main.cpp:2105
    SLOT_CIS_OFF = 13,       // main.cpp:2128
    SLOT_SHUFFLE = 16,       // random_shuffle: molecule = root ligand, partner = running rand() count
    SLOT_INIT = 32           // initial-configuration generator (host)
};

// ROLLED: the ten rounds as a loop (they depend on each other anyway). The streaming kernels unroll them; the fused small-system
// kernel, whose CTAs are in different stages of the step at any moment, is bound by its instruction-cache footprint (32 KB L1.5:
// sm__icc_request_hit_rate 62 % with every block inlined and unrolled) and takes the compact form: 36.6 -> 34.4 us per step at
// 1 024 replicas (the membrane's proposal kernels lose 2-3 us with it, so they keep the unrolled one).
template <bool ROLLED = false>
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll (ROLLED ? 1 : 10)
    for (int r = 0; r < 10; r++) {
#ifdef __CUDA_ARCH__
        uint32_t h0 = __umulhi(M0, c[0]), l0 = M0 * c[0], h1 = __umulhi(M1, c[2]), l1 = M1 * c[2];
#else
        uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
        uint32_t h0 = (uint32_t)(p0 >> 32), l0 = (uint32_t)p0, h1 = (uint32_t)(p1 >> 32), l1 = (uint32_t)p1;
#endif
        uint32_t n0 = h1 ^ c[1] ^ k0, n2 = h0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = l1; c[2] = n2; c[3] = l0;
        k0 += W0; k1 += W1;
    }
}

// U[0,1) with 53 random bits. One Philox block serves TWO draws: slots 2b and 2b+1 share the block whose counter carries the even
// slot number, slot 2b takes the words (x1:x0), slot 2b+1 the words (x3:x2). keyed_uniform2 returns both halves of a block
// (the proposal kernels need slots 0..5 of one molecule: three blocks instead of six).
__host__ __device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
    const uint64_t bits = ((uint64_t)hi << 32) | lo;
    return (double)(bits >> 11) * (1.0 / 9007199254740992.0);
}
template <bool ROLLED = false>
__host__ __device__ __forceinline__ void keyed_uniform2(uint64_t seed, uint32_t mol, uint32_t partner, uint64_t step, uint32_t evenSlot, double &ua, double &ub) {
    uint32_t c[4] = {mol, partner, (uint32_t)step, evenSlot | ((uint32_t)(step >> 32) << 8)};
    philox4x32_10<ROLLED>(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    ua = u53(c[1], c[0]); ub = u53(c[3], c[2]);
}
template <bool ROLLED = false>
__host__ __device__ __forceinline__ double keyed_uniform(uint64_t seed, uint32_t mol, uint32_t partner, uint64_t step, uint32_t slot) {
    uint32_t c[4] = {mol, partner, (uint32_t)step, (slot & ~1u) | ((uint32_t)(step >> 32) << 8)};
    philox4x32_10<ROLLED>(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    return (slot & 1u) ? u53(c[3], c[2]) : u53(c[1], c[0]);
}
// 31-bit integer, the stand-in for libc rand() (RAND_MAX = 2^31-1)
__host__ __device__ __forceinline__ int keyed_rand31(uint64_t seed, uint32_t mol, uint32_t count, uint64_t step) {
    uint32_t c[4] = {mol, count, (uint32_t)step, (uint32_t)SLOT_SHUFFLE | ((uint32_t)(step >> 32) << 8)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    return (int)(c[0] >> 1);
}

}  // namespace kmc
