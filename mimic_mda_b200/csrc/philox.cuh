// philox.cuh -- Philox4x32-10 counter-based generator (Salmon, Moraes, Dror, Shaw,
// "Parallel random numbers: as easy as 1, 2, 3", SC'11), written out from the paper.
// Counter-based means walker w of bin b at half-step t always draws the same numbers,
// whatever the grid shape, tiling or GPU count -- which is what makes the device-resident
// sampler reproducible and checkable against a CPU restatement.
#pragma once
#include <stdint.h>

namespace mda {

struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(Philox4 ctr, uint32_t k0, uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;      // round multipliers
    constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;      // Weyl key increments
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mulhi32(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = mulhi32(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = Philox4{hi1 ^ ctr.y ^ k0, lo1, hi0 ^ ctr.w ^ k1, lo0};
        k0 += W0;
        k1 += W1;
    }
    return ctr;
}

// The same with the ten round keys (k + r * Weyl increment) made beforehand: rk[2 r] and rk[2 r + 1] are round r's.  When
// rk lies in a kernel's parameter block the keys are constant-bank operands of the round's XORs: a round is two wide
// multiplies and two three-input XORs, nothing else.
__host__ __device__ __forceinline__ void philox_round_keys(uint32_t k0, uint32_t k1, uint32_t (&rk)[20]) {
    for (int r = 0; r < 10; ++r) {
        rk[2 * r] = k0 + (uint32_t)r * 0x9E3779B9u;
        rk[2 * r + 1] = k1 + (uint32_t)r * 0xBB67AE85u;
    }
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10_keys(Philox4 ctr, const uint32_t (&rk)[20]) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mulhi32(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = mulhi32(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = Philox4{hi1 ^ ctr.y ^ rk[2 * r], lo1, hi0 ^ ctr.w ^ rk[2 * r + 1], lo0};
    }
    return ctr;
}

// 53-bit uniform in [0, 1) from two 32-bit words (the MT19937 genrand_res53 recipe).
__host__ __device__ __forceinline__ double uniform53(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

}  // namespace mda
