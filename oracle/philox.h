// oracle/philox.h -- TEST INFRASTRUCTURE. Philox4x32-10 (Salmon et al., SC'11) restated for the
// CPU oracle; the product has its own device implementation in csrc/kmc_philox.cuh and the two
// are only ever compared through their outputs. Checked against the Random123 known-answer
// vectors in tests/test_oracle_philox.py.
//
// Draw keying shared by oracle and product (DESIGN.md "Random streams"):
//   key     = 64-bit seed (lo, hi)
//   counter = (molecule id [1-based, reference numbering], partner code, step, slot)
//   U[0,1)  = ((x1:x0) >> 11) * 2^-53 for an even slot, ((x3:x2) >> 11) * 2^-53 for the odd slot that shares its block
//             (counter word 3 carries the even slot number; 53 random bits, like generate_canonical<double,53>)
//   rand()  = x0 >> 1                             (31 bits, RAND_MAX = 2^31-1)
#pragma once
#include <cstdint>

namespace kmco {

struct Philox4 { uint32_t v[4]; };

static inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    return Philox4{{c0, c1, c2, c3}};
}

enum Slot : uint32_t {
    SLOT_MOVE0 = 0,           // 0..5: translation / rotation draws of a unit (main.cpp:585-611, 693-726, 909-944, 990-1089)
    SLOT_RL_ON = 8,           // main.cpp:1919  molecule = receptor i, partner = 4*j + k
    SLOT_MONO_CIS_ON = 9,     // main.cpp:1985  molecule = i, partner = j
    SLOT_CIS_ON = 10,         // main.cpp:2039
    SLOT_RL_OFF = 11,         // main.cpp:2070
    SLOT_MONO_CIS_OFF = 12,   // main.cpp:2105
    SLOT_CIS_OFF = 13,        // main.cpp:2128
    SLOT_SHUFFLE = 16,        // main.cpp:1285/1345/1413/1597  molecule = root ligand, partner = running rand() count
};

// one Philox block serves two draws: slots 2b and 2b+1 share the block whose counter carries the even slot number;
// slot 2b takes (x1:x0), slot 2b+1 takes (x3:x2)
static inline double keyed_uniform(uint64_t seed, uint32_t mol, uint32_t partner, uint64_t step, uint32_t slot) {
    Philox4 r = philox4x32_10(mol, partner, (uint32_t)step, (slot & ~1u) | ((uint32_t)(step >> 32) << 8),
                              (uint32_t)seed, (uint32_t)(seed >> 32));
    const int h = (slot & 1u) ? 2 : 0;
    uint64_t bits = ((uint64_t)r.v[h + 1] << 32) | r.v[h];
    return (double)(bits >> 11) * (1.0 / 9007199254740992.0);
}

static inline int keyed_rand31(uint64_t seed, uint32_t mol, uint32_t count, uint64_t step) {
    Philox4 r = philox4x32_10(mol, count, (uint32_t)step, (uint32_t)SLOT_SHUFFLE | ((uint32_t)(step >> 32) << 8),
                              (uint32_t)seed, (uint32_t)(seed >> 32));
    return (int)(r.v[0] >> 1);
}

}  // namespace kmco
