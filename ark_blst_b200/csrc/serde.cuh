// Point (de)serialisation on the device: the ZCash / IETF encodings of BLS12-381
// points ↔ the blst limb layout.
// Reference: CanonicalSerialize / CanonicalDeserialize for G1Affine and G2Affine
// (src/g1.rs:358-431, src/g2.rs:338-411): blstrs to_compressed / to_uncompressed on the way out,
// from_compressed_unchecked / from_uncompressed_unchecked plus Valid::check (on curve ∧ torsion
// free, src/g1.rs:386-396) on the way in; pinned cross-implementation by src/tests.rs:70-96.
//
// Encoding: big-endian coordinates (48 bytes per Fp; an Fp2 is written c1 then c0), flag bits in
// byte 0: 0x80 compressed, 0x40 infinity (all other bits zero), 0x20 "y is the lexicographically
// larger of (y, −y)" (compressed only).  Decompression takes y = sqrt(x³ + b): Fp by a^((p+1)/4)
// (p ≡ 3 mod 4), Fp2 by Adj / Rodríguez-Henríquez Algorithm 9; the subgroup test is [r]·P = O.
// One thread per point; these kernels are throughput-trivial next to an MSM, so the point
// routines are the out-of-line ones.
#pragma once
#include "ec.cuh"
#include "scalar.cuh"

namespace b200msm {

__device__ __constant__ const uint32_t EXP_P_PLUS1_DIV4[12] = {0xffffeaab, 0xee7fbfff, 0xac54ffff, 0x07aaffff, 0x3dac3d89, 0xd9cc34a8,
                                                               0x3ce144af, 0xd91dd2e1, 0x90d2eb35, 0x92c6e9ed, 0x8e5ff9a6, 0x0680447a};
__device__ __constant__ const uint32_t EXP_P_MINUS3_DIV4[12] = {0xffffeaaa, 0xee7fbfff, 0xac54ffff, 0x07aaffff, 0x3dac3d89, 0xd9cc34a8,
                                                                0x3ce144af, 0xd91dd2e1, 0x90d2eb35, 0x92c6e9ed, 0x8e5ff9a6, 0x0680447a};
__device__ __constant__ const uint32_t EXP_P_MINUS1_DIV2[12] = {0xffffd555, 0xdcff7fff, 0x58a9ffff, 0x0f55ffff, 0x7b587b12, 0xb3986950,
                                                                0x79c2895f, 0xb23ba5c2, 0x21a5d66b, 0x258dd3db, 0x1cbff34d, 0x0d0088f5};
__device__ __constant__ const uint32_t FP_FOUR[12] = {0x000cfff3, 0xaa270000, 0xfc34000a, 0x53cc0032, 0x6b0a807f, 0x478fe97a,
                                                      0xe6ba24d7, 0xb1d37ebe, 0xbf78ab2f, 0x8ec9733b, 0x3d83de7e, 0x09d64551};

// r = a^e, e a 384-bit constant (MSB-first square-and-multiply; variable time is fine here)
template <class F> __device__ __noinline__ void f_pow(F &r, const F &a, const uint32_t *e) {
    F acc;
    f_set_one(acc);
    bool started = false;
    for (int w = 11; w >= 0; w--) {
        uint32_t limb = e[w];
        for (int b = 31; b >= 0; b--) {
            if (started) f_sqr(acc, acc);
            if ((limb >> b) & 1) {
                if (started) f_mul(acc, acc, a);
                else { acc = a; started = true; }
            }
        }
    }
    r = acc;
}
__device__ __forceinline__ bool f_eq(const fp &a, const fp &b) { return fp_eq(a, b); }
__device__ __forceinline__ bool f_eq(const fp2 &a, const fp2 &b) { return fp_eq(a.c0, b.c0) && fp_eq(a.c1, b.c1); }

// canonical (non-Montgomery) limbs ↔ Montgomery
__device__ __forceinline__ void fp_to_mont(fp &r, const fp &a) {
    fp r2;
    fp_load(r2, FP_R2);
    fp_mul(r, a, r2);
}
__device__ __forceinline__ void fp_from_mont(fp &r, const fp &a) {
    fp one;
    fp_set_zero(one);
    one.l[0] = 1;
    fp_mul(r, a, one);
}
__device__ __forceinline__ bool fp_canon_lt_p(const fp &a) {     // canonical limbs < p
#pragma unroll
    for (int i = 11; i >= 0; i--) {
        if (a.l[i] < FP_P[i]) return true;
        if (a.l[i] > FP_P[i]) return false;
    }
    return false;
}
__device__ __forceinline__ bool fp_canon_gt_half(const fp &a) {  // canonical limbs > (p−1)/2
#pragma unroll
    for (int i = 11; i >= 0; i--) {
        if (a.l[i] > EXP_P_MINUS1_DIV2[i]) return true;
        if (a.l[i] < EXP_P_MINUS1_DIV2[i]) return false;
    }
    return false;
}
// 48 big-endian bytes ↔ canonical little-endian limbs
__device__ __forceinline__ void fp_from_be(fp &r, const uint8_t *b) {
#pragma unroll
    for (int j = 0; j < 12; j++) {
        const uint8_t *q = b + 44 - 4 * j;
        r.l[j] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
    }
}
__device__ __forceinline__ void fp_to_be(uint8_t *b, const fp &a) {
#pragma unroll
    for (int j = 0; j < 12; j++) {
        uint8_t *q = b + 44 - 4 * j;
        q[0] = (uint8_t)(a.l[j] >> 24); q[1] = (uint8_t)(a.l[j] >> 16); q[2] = (uint8_t)(a.l[j] >> 8); q[3] = (uint8_t)a.l[j];
    }
}

// ---- per-field pieces ----
// coordinate from bytes (flags already cleared by the caller in byte 0 via `first`): false if ≥ p
__device__ __forceinline__ bool coord_read(fp &v, const uint8_t *b, uint8_t first) {
    fp c;
    fp_from_be(c, b);
    c.l[11] = (c.l[11] & 0x00ffffffu) | ((uint32_t)first << 24);
    if (!fp_canon_lt_p(c)) return false;
    fp_to_mont(v, c);
    return true;
}
__device__ __forceinline__ bool coord_read(fp2 &v, const uint8_t *b, uint8_t first) {  // c1 then c0
    return coord_read(v.c1, b, first) && coord_read(v.c0, b + 48, b[48]);
}
__device__ __forceinline__ void coord_write(uint8_t *b, const fp &v) {
    fp c;
    fp_from_mont(c, v);
    fp_to_be(b, c);
}
__device__ __forceinline__ void coord_write(uint8_t *b, const fp2 &v) {
    coord_write(b, v.c1);
    coord_write(b + 48, v.c0);
}
__device__ __forceinline__ bool lex_largest(const fp &y) {
    fp c;
    fp_from_mont(c, y);
    return fp_canon_gt_half(c);
}
__device__ __forceinline__ bool lex_largest(const fp2 &y) {
    fp c1, c0;
    fp_from_mont(c1, y.c1);
    if (fp_canon_gt_half(c1)) return true;
    if (!fp_is_zero(c1)) return false;
    fp_from_mont(c0, y.c0);
    return fp_canon_gt_half(c0);
}
__device__ __forceinline__ void curve_b(fp &b) { fp_load(b, FP_FOUR); }                    // y² = x³ + 4
__device__ __forceinline__ void curve_b(fp2 &b) { fp_load(b.c0, FP_FOUR); fp_load(b.c1, FP_FOUR); }  // 4(1+u)

__device__ __forceinline__ bool f_sqrt(fp &r, const fp &a) {
    f_pow(r, a, EXP_P_PLUS1_DIV4);
    fp t;
    fp_sqr(t, r);
    return fp_eq(t, a);
}
__device__ __forceinline__ bool f_sqrt(fp2 &r, const fp2 &a) {
    if (f_is_zero(a)) { f_set_zero(r); return true; }
    fp2 a1, alpha, x0, t, minus_one;
    f_set_one(minus_one);
    f_cneg(minus_one, minus_one, true);
    f_pow(a1, a, EXP_P_MINUS3_DIV4);
    f_sqr(alpha, a1);
    f_mul(alpha, alpha, a);
    t = alpha;
    fp_neg(t.c1, t.c1);                        // α^p (Frobenius = conjugation)
    f_mul(t, t, alpha);
    if (f_eq(t, minus_one)) return false;
    f_mul(x0, a1, a);
    if (f_eq(alpha, minus_one)) {              // x = u·x0
        r.c0 = x0.c1;
        fp_neg(r.c0, r.c0);
        r.c1 = x0.c0;
    } else {
        fp2 one, b;
        f_set_one(one);
        f_add(t, one, alpha);
        f_pow(b, t, EXP_P_MINUS1_DIV2);
        f_mul(r, b, x0);
    }
    f_sqr(t, r);
    return f_eq(t, a);
}
template <class F> __device__ __forceinline__ bool on_curve(const F &x, const F &y) {
    F l, r, b;
    f_sqr(l, y);
    f_sqr(r, x);
    f_mul(r, r, x);
    curve_b(b);
    f_add(r, r, b);
    return f_eq(l, r);
}
// [r]·(x, y) == O ?
template <class F> __device__ __noinline__ bool torsion_free(const F &x, const F &y) {
    xyzz<F> acc;
    xyzz_set_inf(acc);
    for (int bit = 254; bit >= 0; bit--) {
        xyzz_dbl_ni(acc);
        if ((FR_MOD[bit >> 5] >> (bit & 31)) & 1) xyzz_madd_ni(acc, x, y);
    }
    return xyzz_is_inf(acc);
}

// status: 0 ok, 1 malformed (what blstrs from_*_unchecked rejects), 2 fails Valid::check
template <class F>
__global__ void __launch_bounds__(64)
k_deserialize(const uint8_t *__restrict__ in, size_t n, int compressed, int validate, uint32_t *__restrict__ aff,
              uint8_t *__restrict__ status) {
    constexpr int W = field_words<F>::value;
    constexpr int CB = W * 4;                   // bytes per coordinate
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t *b = in + i * (compressed ? CB : 2 * CB);
    uint32_t *o = aff + i * 2 * W;
    const uint8_t flags = b[0] & 0xE0, first = b[0] & 0x1F;
    F x, y;
    f_set_zero(x);
    f_set_zero(y);
    uint8_t st = 0;
    if (((flags & 0x80) != 0) != (compressed != 0)) st = 1;
    else if (flags & 0x40) {                    // infinity: every other bit must be clear
        uint8_t any = first | (flags & 0x20);
        for (int k = 1; k < (compressed ? CB : 2 * CB); k++) any |= b[k];
        if (any) st = 1;
    } else if (!compressed && (flags & 0x20)) st = 1;
    else if (!coord_read(x, b, first)) st = 1;
    else {
        if (compressed) {
            F rhs, cb_;
            f_sqr(rhs, x);
            f_mul(rhs, rhs, x);
            curve_b(cb_);
            f_add(rhs, rhs, cb_);
            if (!f_sqrt(y, rhs)) st = 1;
            else if (lex_largest(y) != ((flags & 0x20) != 0)) f_cneg(y, y, true);
        } else if (!coord_read(y, b + CB, b[CB])) st = 1;
        if (st == 0 && validate) {
            if ((!compressed && !on_curve(x, y)) || !torsion_free(x, y)) st = 2;
        }
    }
    if (st == 1) { f_set_zero(x); f_set_zero(y); }
    f_store(o, x);
    f_store(o + W, y);
    status[i] = st;
}

template <class F>
__global__ void __launch_bounds__(64)
k_serialize(const uint32_t *__restrict__ aff, size_t n, int compressed, uint8_t *__restrict__ out) {
    constexpr int W = field_words<F>::value;
    constexpr int CB = W * 4;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t *b = out + i * (compressed ? CB : 2 * CB);
    F x, y;
    f_load(x, aff + i * 2 * W);
    f_load(y, aff + i * 2 * W + W);
    if (f_is_zero(x) && f_is_zero(y)) {         // identity
        for (int k = 0; k < (compressed ? CB : 2 * CB); k++) b[k] = 0;
        b[0] = compressed ? 0xC0 : 0x40;
        return;
    }
    coord_write(b, x);
    if (compressed) b[0] |= 0x80 | (lex_largest(y) ? 0x20 : 0);
    else coord_write(b + CB, y);
}

}  // namespace b200msm
