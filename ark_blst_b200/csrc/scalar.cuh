// Scalar-side device helpers: Fr canonicalisation and Booth window digits.
#pragma once
#include <cstdint>

namespace b200msm {

// ---- Fr: canonicalise the scalar (reference src/scalar.rs:450-463 does this on the host with a
// BigUint allocation per element; here it is 8 limbs of Montgomery reduction on the device) ----
__device__ __constant__ const uint32_t FR_MOD[8] = {0x00000001, 0xffffffff, 0xfffe5bfe, 0x53bda402,
                                                    0x09a1d805, 0x3339d808, 0x299d7d48, 0x73eda753};

__device__ __forceinline__ bool fr_geq_r(const uint32_t *s) {
#pragma unroll
    for (int i = 7; i >= 0; i--) {
        if (s[i] > FR_MOD[i]) return true;
        if (s[i] < FR_MOD[i]) return false;
    }
    return true;
}
__device__ __forceinline__ void fr_sub_r(uint32_t *s) {
    uint64_t brw = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t d = (uint64_t)s[i] - FR_MOD[i] - brw;
        s[i] = (uint32_t)d;
        brw = (d >> 32) & 1;
    }
}
// s ← s·2^-256 mod r  (Montgomery form → canonical integer); -r^-1 mod 2^32 = 0xffffffff
__device__ __forceinline__ void fr_from_mont(uint32_t *s) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t m = 0u - s[0];  // s[0]·0xffffffff
        uint64_t c = ((uint64_t)m * FR_MOD[0] + s[0]) >> 32;
#pragma unroll
        for (int j = 1; j < 8; j++) {
            c += (uint64_t)m * FR_MOD[j] + s[j];
            s[j - 1] = (uint32_t)c;
            c >>= 32;
        }
        s[7] = (uint32_t)c;
    }
    if (fr_geq_r(s)) fr_sub_r(s);
}

// c+1 bits of s starting at bit `lo` (lo ≥ -1; bits outside [0,256) read as 0)
__device__ __forceinline__ uint32_t scalar_bits(const uint32_t *s, int lo, int cnt) {
    uint32_t mask = (1u << cnt) - 1;
    if (lo < 0) return (s[0] << 1) & mask;
    int w = lo >> 5, sh = lo & 31;
    uint64_t v = s[w];
    if (w + 1 < 8) v |= (uint64_t)s[w + 1] << 32;
    return (uint32_t)(v >> sh) & mask;
}
// Booth-recoded signed digit of window w: u + b[wc-1] − 2^c·b[wc+c-1] ∈ [−2^(c-1), 2^(c-1)].
// Each window is independent of the others (no sequential carry), Σ d_w·2^(wc) = s for s < 2^255.
__device__ __forceinline__ int booth_digit(const uint32_t *s, int w, int c) {
    uint32_t v = scalar_bits(s, w * c - 1, c + 1);
    int d = (int)((v >> 1) & ((1u << c) - 1)) + (int)(v & 1);
    if ((v >> c) & 1) d -= (1 << c);
    return d;
}

__device__ __forceinline__ void load_scalar(uint32_t *s, const uint32_t *scalars, size_t i, int mont) {
    const uint4 *p = reinterpret_cast<const uint4 *>(scalars + 8 * i);
    uint4 a = p[0], b = p[1];
    s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w;
    s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w;
    if (mont) fr_from_mont(s);
    else { if (fr_geq_r(s)) fr_sub_r(s); if (fr_geq_r(s)) fr_sub_r(s); }  // 2^256 < 3r
}

// ---- GLV decomposition for G1 -------------------------------------------------------------------
// BLS12-381 has r = λ² + λ + 1 with λ = z² − 1 (128 bits) and the endomorphism φ(x, y) = (βx, y) = λ·(x, y)
// on G1.  For a canonical k < r:  k = k1 + k2·λ with k2 = ⌊k/λ⌋ ≤ λ + 1 and k1 = k mod λ — both
// non-negative and below 2^128 — so k·P = k1·P + k2·φ(P): twice the points, half the scalar bits,
// i.e. half the windows for the Horner chain (the MSM's serial tail).  ⌊k/λ⌋ by Barrett with
// μ = ⌊2^256/λ⌋ (at most one correction step; checked against big-int arithmetic in the tests).
__device__ __constant__ const uint32_t GLV_LAMBDA[4] = {0xffffffff, 0x00000000, 0x0001a402, 0xac45a401};
__device__ __constant__ const uint32_t GLV_MU[5] = {0xf6cfee30, 0x63f6e522, 0xe01faadd, 0x7c6becf1, 0x00000001};

__device__ __forceinline__ void glv_decompose(const uint32_t *k, uint32_t *k1, uint32_t *k2) {
    // q = (k·μ) >> 256 : 8×5 limbs, only limbs 8..12 of the product are kept
    uint32_t prod[13];
#pragma unroll
    for (int i = 0; i < 13; i++) prod[i] = 0;
#pragma unroll
    for (int j = 0; j < 5; j++) {
        uint64_t c = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            c += (uint64_t)k[i] * GLV_MU[j] + prod[i + j];
            prod[i + j] = (uint32_t)c;
            c >>= 32;
        }
        prod[8 + j] = (uint32_t)c;
    }
    uint32_t q[5];
#pragma unroll
    for (int i = 0; i < 5; i++) q[i] = prod[8 + i];
    // r = k − q·λ on 5 limbs (the true remainder is < 3λ < 2^130)
    uint32_t ql[5];
#pragma unroll
    for (int i = 0; i < 5; i++) ql[i] = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        uint64_t c = 0;
#pragma unroll
        for (int i = 0; i + j < 5; i++) {
            c += (uint64_t)q[i] * GLV_LAMBDA[j] + ql[i + j];
            ql[i + j] = (uint32_t)c;
            c >>= 32;
        }
    }
    uint32_t r[5];
    uint64_t brw = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) {
        uint64_t d = (uint64_t)k[i] - ql[i] - brw;
        r[i] = (uint32_t)d;
        brw = (d >> 32) & 1;
    }
    for (int it = 0; it < 3; it++) {                   // r ≥ λ → r −= λ, q += 1 (≤ 2 times)
        bool ge = r[4] != 0;
        if (!ge) {
            ge = true;
#pragma unroll
            for (int i = 3; i >= 0; i--) {
                if (r[i] > GLV_LAMBDA[i]) break;
                if (r[i] < GLV_LAMBDA[i]) { ge = false; break; }
            }
        }
        if (!ge) break;
        brw = 0;
#pragma unroll
        for (int i = 0; i < 5; i++) {
            uint64_t d = (uint64_t)r[i] - (i < 4 ? GLV_LAMBDA[i] : 0u) - brw;
            r[i] = (uint32_t)d;
            brw = (d >> 32) & 1;
        }
        uint64_t c = 1;
#pragma unroll
        for (int i = 0; i < 5; i++) {
            c += q[i];
            q[i] = (uint32_t)c;
            c >>= 32;
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        k1[i] = i < 4 ? r[i] : 0;
        k2[i] = i < 4 ? q[i] : 0;
    }
}

}  // namespace b200msm
