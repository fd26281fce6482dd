// Fp2 = Fp[u]/(u²+1) as (c0, c1), the layout of blst_fp2 / reference src/fp2.rs:450-454,
// plus the field-generic overload set (f_*) that the curve code in ec.cuh is written against.
#pragma once
#include "fp.cuh"

namespace b200msm {

struct fp2 {
    fp c0, c1;
};

// ---- Fp2 ----
__device__ __forceinline__ void fp2_add(fp2 &r, const fp2 &a, const fp2 &b) {
    fp_add(r.c0, a.c0, b.c0);
    fp_add(r.c1, a.c1, b.c1);
}
__device__ __forceinline__ void fp2_sub(fp2 &r, const fp2 &a, const fp2 &b) {
    fp_sub(r.c0, a.c0, b.c0);
    fp_sub(r.c1, a.c1, b.c1);
}
// Karatsuba: 3 Fp products (work model: Fp2 mul = 3 Fp mul)
__device__ __forceinline__ void fp2_mul(fp2 &r, const fp2 &a, const fp2 &b) {
    fp aa, bb, sa, sb;
    fp_mul(aa, a.c0, b.c0);
    fp_mul(bb, a.c1, b.c1);
    fp_add(sa, a.c0, a.c1);
    fp_add(sb, b.c0, b.c1);
    fp_mul(sa, sa, sb);
    fp_sub(sa, sa, aa);
    fp_sub(r.c1, sa, bb);
    fp_sub(r.c0, aa, bb);
}
// (a0+a1)(a0-a1) + 2·a0·a1·u : 2 Fp products
__device__ __forceinline__ void fp2_sqr(fp2 &r, const fp2 &a) {
    fp s, d, m;
    fp_add(s, a.c0, a.c1);
    fp_sub(d, a.c0, a.c1);
    fp_mul(m, a.c0, a.c1);
    fp_mul(r.c0, s, d);
    fp_add(r.c1, m, m);
}

// ---- field-generic overloads ----
__device__ __forceinline__ void f_mul(fp &r, const fp &a, const fp &b) { fp_mul(r, a, b); }
__device__ __forceinline__ void f_sqr(fp &r, const fp &a) { fp_sqr(r, a); }
__device__ __forceinline__ void f_add(fp &r, const fp &a, const fp &b) { fp_add(r, a, b); }
__device__ __forceinline__ void f_sub(fp &r, const fp &a, const fp &b) { fp_sub(r, a, b); }
__device__ __forceinline__ void f_dbl(fp &r, const fp &a) { fp_add(r, a, a); }
__device__ __forceinline__ void f_cneg(fp &r, const fp &a, bool n) { fp_cneg(r, a, n); }
__device__ __forceinline__ bool f_is_zero(const fp &a) { return fp_is_zero(a); }
__device__ __forceinline__ void f_set_zero(fp &r) { fp_set_zero(r); }
__device__ __forceinline__ void f_set_one(fp &r) { fp_set_one(r); }
__device__ __forceinline__ void f_load(fp &r, const uint32_t *p) { fp_load(r, p); }
__device__ __forceinline__ void f_store(uint32_t *p, const fp &a) { fp_store(p, a); }
__device__ __forceinline__ void f_inv(fp &r, const fp &a) { fp_inv(r, a); }

__device__ __forceinline__ void f_mul(fp2 &r, const fp2 &a, const fp2 &b) { fp2_mul(r, a, b); }
__device__ __forceinline__ void f_sqr(fp2 &r, const fp2 &a) { fp2_sqr(r, a); }
__device__ __forceinline__ void f_add(fp2 &r, const fp2 &a, const fp2 &b) { fp2_add(r, a, b); }
__device__ __forceinline__ void f_sub(fp2 &r, const fp2 &a, const fp2 &b) { fp2_sub(r, a, b); }
__device__ __forceinline__ void f_dbl(fp2 &r, const fp2 &a) { fp2_add(r, a, a); }
__device__ __forceinline__ void f_cneg(fp2 &r, const fp2 &a, bool n) {
    fp_cneg(r.c0, a.c0, n);
    fp_cneg(r.c1, a.c1, n);
}
__device__ __forceinline__ bool f_is_zero(const fp2 &a) { return fp_is_zero(a.c0) && fp_is_zero(a.c1); }
__device__ __forceinline__ void f_set_zero(fp2 &r) { fp_set_zero(r.c0); fp_set_zero(r.c1); }
__device__ __forceinline__ void f_set_one(fp2 &r) { fp_set_one(r.c0); fp_set_zero(r.c1); }
__device__ __forceinline__ void f_load(fp2 &r, const uint32_t *p) { fp_load(r.c0, p); fp_load(r.c1, p + 12); }
__device__ __forceinline__ void f_store(uint32_t *p, const fp2 &a) { fp_store(p, a.c0); fp_store(p + 12, a.c1); }
// 1/(c0 + c1·u) = (c0 - c1·u)/(c0² + c1²)
__device__ __forceinline__ void f_inv(fp2 &r, const fp2 &a) {
    fp n, t;
    fp_sqr(n, a.c0);
    fp_sqr(t, a.c1);
    fp_add(n, n, t);
    fp_inv(n, n);
    fp_mul(r.c0, a.c0, n);
    fp_mul(t, a.c1, n);
    fp_neg(r.c1, t);
}

// word accessors with compile-time indices (keep everything in registers under full unrolling)
__device__ __forceinline__ uint32_t f_word(const fp &a, int i) { return a.l[i]; }
__device__ __forceinline__ void f_set_word(fp &a, int i, uint32_t v) { a.l[i] = v; }
__device__ __forceinline__ uint32_t f_word(const fp2 &a, int i) { return i < 12 ? a.c0.l[i] : a.c1.l[i - 12]; }
__device__ __forceinline__ void f_set_word(fp2 &a, int i, uint32_t v) {
    if (i < 12) a.c0.l[i] = v;
    else a.c1.l[i - 12] = v;
}

template <class F> struct field_words;
template <> struct field_words<fp> { static constexpr int value = 12; };   // u32 words per element
template <> struct field_words<fp2> { static constexpr int value = 24; };

}  // namespace b200msm
