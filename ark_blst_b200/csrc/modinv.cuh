// Modular inversion in Fp by divsteps ("safegcd", Bernstein–Yang 2019; the signed-30-bit-limb
// formulation published in bitcoin-core/secp256k1 as modinv32, restated for the 381-bit BLS12-381
// modulus with 13 limbs).  Used for the ONE inversion a batch of affine additions shares
// (batch_affine.cuh): ≈25 k instructions — 30 rounds of 30 branch-free divsteps on the low limbs
// plus a 2×2 matrix update of the 390-bit state — against ≈190 k for the Fermat power a^(p−2)
// (fp_inv), and almost all of it on the integer ALU pipe rather than the IMAD pipe.
// Branch-free with a fixed trip count: the 32 lanes of a warp run in lock step.
//
// Plain integer code (int32 / int64): compiled by g++ for the CPU unit test
// (tests/cpp/test_modinv.cpp) and by nvcc for the device.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define MODINV_HD __host__ __device__ __forceinline__
#else
#define MODINV_HD inline
#endif

namespace b200msm {

constexpr int MI_LIMBS = 13;                       // 13 × 30 = 390 bits ≥ 381 + sign/slack
struct mi_s30 { int32_t v[MI_LIMBS]; };            // Σ v[i]·2^(30i), limbs in (−2^30, 2^30), top limb carries the sign
struct mi_t2x2 { int32_t u, v, q, r; };            // transition matrix of 30 divsteps, scaled by 2^30

// p in 30-bit limbs, and p^-1 mod 2^30
MODINV_HD int32_t mi_p30(int i) {
    constexpr int32_t P30[MI_LIMBS] = {0x3fffaaab, 0x27fbffff, 0x153ffffb, 0x2affffac, 0x30f6241e, 0x034a83da, 0x112bf673,
                                       0x12e13ce1, 0x2cd76477, 0x1ed90d2e, 0x29a4b1ba, 0x3a8e5ff9, 0x001a0111};
    return P30[i];
}
constexpr uint32_t MI_PINV30 = 0x30003u;
// divsteps needed for inputs below 2^381 with the half-delta start: ⌊(45907·381 + 26313)/19929⌋ = 878 ≤ 30·30
constexpr int MI_ROUNDS = 30;

// 30 divsteps on the low limbs; zeta = −(delta + 1/2).  Returns the new zeta and the matrix t with
// t·[f, g] = 2^30·[f', g'].
MODINV_HD int32_t mi_divsteps_30(int32_t zeta, uint32_t f0, uint32_t g0, mi_t2x2 &t) {
    uint32_t u = 1, v = 0, q = 0, r = 1, f = f0, g = g0;
#pragma unroll 6
    for (int i = 0; i < 30; i++) {
        uint32_t m1 = (uint32_t)(zeta >> 31);      // all ones iff zeta < 0
        uint32_t m2 = 0u - (g & 1u);               // all ones iff g odd
        uint32_t x = (f ^ m1) - m1, y = (u ^ m1) - m1, z = (v ^ m1) - m1;   // conditionally negated f, u, v
        g += x & m2;
        q += y & m2;
        r += z & m2;
        m1 &= m2;                                  // swap-and-negate case: zeta < 0 and g odd
        zeta = (int32_t)((uint32_t)zeta ^ m1) - 1;
        f += g & m1;
        u += q & m1;
        v += r & m1;
        g >>= 1;
        u <<= 1;
        v <<= 1;
    }
    t.u = (int32_t)u; t.v = (int32_t)v; t.q = (int32_t)q; t.r = (int32_t)r;
    return zeta;
}

// [d, e] ← t·[d, e] / 2^30 mod p, d and e staying in (−2p, p)
MODINV_HD void mi_update_de(mi_s30 &d, mi_s30 &e, const mi_t2x2 &t) {
    const int32_t M30 = (int32_t)(0xffffffffu >> 2);
    const int32_t u = t.u, v = t.v, q = t.q, r = t.r;
    const int32_t sd = d.v[MI_LIMBS - 1] >> 31, se = e.v[MI_LIMBS - 1] >> 31;
    int32_t md = (u & sd) + (v & se), me = (q & sd) + (r & se);   // + p·[u,q] if d < 0, + p·[v,r] if e < 0
    int32_t di = d.v[0], ei = e.v[0];
    int64_t cd = (int64_t)u * di + (int64_t)v * ei, ce = (int64_t)q * di + (int64_t)r * ei;
    // the multiple of p that clears the low 30 bits
    md -= (int32_t)((MI_PINV30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
    me -= (int32_t)((MI_PINV30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
    cd += (int64_t)mi_p30(0) * md;
    ce += (int64_t)mi_p30(0) * me;
    cd >>= 30;
    ce >>= 30;
#pragma unroll
    for (int i = 1; i < MI_LIMBS; i++) {
        di = d.v[i];
        ei = e.v[i];
        cd += (int64_t)u * di + (int64_t)v * ei;
        ce += (int64_t)q * di + (int64_t)r * ei;
        cd += (int64_t)mi_p30(i) * md;
        ce += (int64_t)mi_p30(i) * me;
        d.v[i - 1] = (int32_t)cd & M30; cd >>= 30;
        e.v[i - 1] = (int32_t)ce & M30; ce >>= 30;
    }
    d.v[MI_LIMBS - 1] = (int32_t)cd;
    e.v[MI_LIMBS - 1] = (int32_t)ce;
}

// [f, g] ← t·[f, g] / 2^30 (exact: the low 30 bits are zero by construction)
MODINV_HD void mi_update_fg(mi_s30 &f, mi_s30 &g, const mi_t2x2 &t) {
    const int32_t M30 = (int32_t)(0xffffffffu >> 2);
    const int32_t u = t.u, v = t.v, q = t.q, r = t.r;
    int32_t fi = f.v[0], gi = g.v[0];
    int64_t cf = (int64_t)u * fi + (int64_t)v * gi, cg = (int64_t)q * fi + (int64_t)r * gi;
    cf >>= 30;
    cg >>= 30;
#pragma unroll
    for (int i = 1; i < MI_LIMBS; i++) {
        fi = f.v[i];
        gi = g.v[i];
        cf += (int64_t)u * fi + (int64_t)v * gi;
        cg += (int64_t)q * fi + (int64_t)r * gi;
        f.v[i - 1] = (int32_t)cf & M30; cf >>= 30;
        g.v[i - 1] = (int32_t)cg & M30; cg >>= 30;
    }
    f.v[MI_LIMBS - 1] = (int32_t)cf;
    g.v[MI_LIMBS - 1] = (int32_t)cg;
}

// r in (−2p, p) → [0, p), negated first when sign < 0 (f ends as ±1: d = ±x^-1)
MODINV_HD void mi_normalize(mi_s30 &r, int32_t sign) {
    const int32_t M30 = (int32_t)(0xffffffffu >> 2);
    int32_t add = r.v[MI_LIMBS - 1] >> 31;
#pragma unroll
    for (int i = 0; i < MI_LIMBS; i++) r.v[i] += mi_p30(i) & add;
    const int32_t neg = sign >> 31;
#pragma unroll
    for (int i = 0; i < MI_LIMBS; i++) r.v[i] = (r.v[i] ^ neg) - neg;
#pragma unroll
    for (int i = 0; i < MI_LIMBS - 1; i++) { r.v[i + 1] += r.v[i] >> 30; r.v[i] &= M30; }
    add = r.v[MI_LIMBS - 1] >> 31;
#pragma unroll
    for (int i = 0; i < MI_LIMBS; i++) r.v[i] += mi_p30(i) & add;
#pragma unroll
    for (int i = 0; i < MI_LIMBS - 1; i++) { r.v[i + 1] += r.v[i] >> 30; r.v[i] &= M30; }
}

// 12×u32 (value < p) → 13 × 30-bit limbs and back
MODINV_HD void mi_from_u32(mi_s30 &r, const uint32_t a[12]) {
#pragma unroll
    for (int i = 0; i < MI_LIMBS; i++) {
        const int bit = 30 * i, w = bit >> 5, s = bit & 31;
        uint32_t lo = a[w] >> s;
        if (s > 2 && w + 1 < 12) lo |= a[w + 1] << (32 - s);
        r.v[i] = (int32_t)(lo & 0x3fffffffu);
    }
}
MODINV_HD void mi_to_u32(uint32_t a[12], const mi_s30 &r) {
#pragma unroll
    for (int w = 0; w < 12; w++) {
        const int bit = 32 * w, i = bit / 30, s = bit % 30;   // bits [32w, 32w+32) start in limb i at offset s
        uint32_t x = (uint32_t)r.v[i] >> s;
        if (i + 1 < MI_LIMBS) x |= (uint32_t)r.v[i + 1] << (30 - s);
        if (s > 28 && i + 2 < MI_LIMBS) x |= (uint32_t)r.v[i + 2] << (60 - s);
        a[w] = x;
    }
}

// out = a^-1 mod p as plain integers (a < p; 0 ↦ 0).  No Montgomery factor is touched here.
MODINV_HD void mi_inverse_u32(uint32_t out[12], const uint32_t a[12]) {
    mi_s30 d, e, f, g;
#pragma unroll
    for (int i = 0; i < MI_LIMBS; i++) { d.v[i] = 0; e.v[i] = 0; f.v[i] = mi_p30(i); }
    e.v[0] = 1;
    mi_from_u32(g, a);
    int32_t zeta = -1;
#pragma unroll 1
    for (int it = 0; it < MI_ROUNDS; it++) {
        mi_t2x2 t;
        zeta = mi_divsteps_30(zeta, (uint32_t)f.v[0], (uint32_t)g.v[0], t);
        mi_update_de(d, e, t);
        mi_update_fg(f, g, t);
#if defined(__CUDA_ARCH__)
        // g = 0 is a fixed point (f = ±1 and d stay as they are): stop as soon as every lane of the warp got
        // there — typically after ≈24 of the 30 worst-case rounds
        int32_t any = 0;
#pragma unroll
        for (int i = 0; i < MI_LIMBS; i++) any |= g.v[i];
        if (__all_sync(__activemask(), any == 0)) break;
#endif
    }
    mi_normalize(d, f.v[MI_LIMBS - 1]);
    mi_to_u32(out, d);
}

}  // namespace b200msm
