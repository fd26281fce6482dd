// Four-dimensional scalar decomposition for G2 (Galbraith–Lin–Scott with the untwist-Frobenius-twist
// endomorphism ψ).  On the order-r subgroup of the twist ψ acts as multiplication by p ≡ z (mod r),
// z = −0xd201000000010000 the curve parameter, so |z|·Q = −ψ(Q) and the base-|z| digits of a
// canonical scalar,  k = k0 + k1·|z| + k2·|z|² + k3·|z|³  (0 ≤ k_i < |z| < 2^64, |z|⁴ > r),  give
//     k·Q = k0·Q + k1·(−ψ(Q)) + k2·ψ²(Q) + k3·(−ψ³(Q)):
// four times the points, a quarter of the scalar bits — the same number of bucket additions, a
// quarter of the windows to reduce and of the dependent doublings in the Horner chain, which on
// G2 (11 µs per dependent doubling) are a fifth of a 2^20-point MSM and most of a 2^17-point shard.
// ψ(x, y) = (x̄·γx, ȳ·γy), γx = (1+u)^−(p−1)/3, γy = (1+u)^−(p−1)/2;  ψ²(x, y) = (β·x, −y).
// Checked against the big-int oracle in tests/test_glv_constants.py (constants, ψ(Q) = [z]Q, digits).
//
// Plain integer code: compiled by g++ for the CPU test and by nvcc for the device.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define GLS4_HD __host__ __device__ __forceinline__
#else
#define GLS4_HD inline
#endif

namespace b200msm {

constexpr uint64_t GLS4_Z = 0xd201000000010000ull;   // |z|: top bit set, so it is a normalised divisor

// (u1·2^64 + u0) / v for u1 < v and v ≥ 2^63: quotient, remainder in r (Knuth D with two 32-bit digits)
GLS4_HD uint64_t gls4_div128by64(uint64_t u1, uint64_t u0, uint64_t v, uint64_t &r) {
    const uint64_t b = 1ull << 32, vn1 = v >> 32, vn0 = v & 0xffffffffull;
    const uint64_t un1 = u0 >> 32, un0 = u0 & 0xffffffffull;
    uint64_t q1 = u1 / vn1, rhat = u1 - q1 * vn1;
    while (q1 >= b || q1 * vn0 > b * rhat + un1) {
        q1--;
        rhat += vn1;
        if (rhat >= b) break;
    }
    const uint64_t un21 = u1 * b + un1 - q1 * v;
    uint64_t q0 = un21 / vn1;
    rhat = un21 - q0 * vn1;
    while (q0 >= b || q0 * vn0 > b * rhat + un0) {
        q0--;
        rhat += vn1;
        if (rhat >= b) break;
    }
    r = un21 * b + un0 - q0 * v;
    return q1 * b + q0;
}

// k (8×u32, canonical, < r) → its four base-|z| digits, each as 8×u32 with the upper six words zero
GLS4_HD void gls4_decompose(const uint32_t k[8], uint32_t parts[4][8]) {
    uint64_t a[4];
#pragma unroll
    for (int i = 0; i < 4; i++) a[i] = (uint64_t)k[2 * i] | ((uint64_t)k[2 * i + 1] << 32);
#pragma unroll
    for (int d = 0; d < 4; d++) {
        uint64_t rem = 0;
#pragma unroll
        for (int j = 3 - d; j >= 0; j--) {            // the quotient of step d has 4 − d limbs (|z|^d ≥ 2^(63d))
            if (rem == 0 && a[j] < GLS4_Z) {          // quotient digit 0
                rem = a[j];
                a[j] = 0;
            } else a[j] = gls4_div128by64(rem, a[j], GLS4_Z, rem);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) parts[d][i] = 0;
        parts[d][0] = (uint32_t)rem;
        parts[d][1] = (uint32_t)(rem >> 32);
    }
}

}  // namespace b200msm
