/* T1 oracle — field layer.  TEST INFRASTRUCTURE ONLY (see oracle/README.md): nothing under
 * ark_blst_b200/ links or loads this; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do.
 *
 * Restates, in portable C (unsigned __int128), the arithmetic the reference delegates to the
 * un-vendored blst =0.3.10 (Cargo.toml:22): 384-bit Montgomery Fp on 6×u64 little-endian limbs
 * (layout proven by src/fp.rs:482-491; Montgomery one = blstrs::fp::R, src/fp.rs:532), Fp2 =
 * Fp[u]/(u²+1) as (c0,c1) (src/fp2.rs:450-454), and Fr on 4×u64 limbs with R = 2^256
 * (src/scalar.rs:23-25, 476-481).  PARITY UNPINNED by reference vectors (there are none); pinned
 * instead on the reference's constants (fp.rs:25-32, fp.rs:714-721, scalar.rs:476-481) and on the
 * big-int oracle oracle/bls12381.py, limb for limb (tests/test_oracle.py).
 */
#ifndef B200MSM_ORACLE_FIELD_H
#define B200MSM_ORACLE_FIELD_H
#include <stdint.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[6]; } fp_t;
typedef struct { fp_t c0, c1; } fp2_t;
typedef struct { uint64_t l[4]; } fr_t;

/* src/fp.rs:25-32 */
static const fp_t FP_P = {{0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                           0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL}};
/* 2^384 mod p (Montgomery one) and 2^768 mod p; asserted against big-int in tests */
static const fp_t FP_ONE = {{0x760900000002fffdULL, 0xebf4000bc40c0002ULL, 0x5f48985753c758baULL,
                             0x77ce585370525745ULL, 0x5c071a97a256ec6dULL, 0x15f65ec3fa80e493ULL}};
static const fp_t FP_R2 = {{0xf4df1f341c341746ULL, 0x0a76e6a609d104f1ULL, 0x8de5476c4c95b6d5ULL,
                            0x67eb88a9939d83c0ULL, 0x9a793e85b519952dULL, 0x11988fe592cae3aaULL}};
#define FP_PINV 0x89f3fffcfffcfffdULL /* -p^-1 mod 2^64 */

/* src/scalar.rs:476-481 */
static const fr_t FR_R = {{0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL,
                           0x73eda753299d7d48ULL}};
#define FR_RINV 0xfffffffeffffffffULL /* -r^-1 mod 2^64 */

/* ------------------------------------------------------------------ Fp */
static inline int fp_is_zero(const fp_t *a) {
    uint64_t t = 0;
    for (int i = 0; i < 6; i++) t |= a->l[i];
    return t == 0;
}
static inline int fp_eq(const fp_t *a, const fp_t *b) {
    uint64_t t = 0;
    for (int i = 0; i < 6; i++) t |= a->l[i] ^ b->l[i];
    return t == 0;
}
/* r = a - p if a >= p  (hi = carry-out word of a).  add / sub / final subtraction go through the
 * compiler's carry-flag builtins on x86-64 (adc / sbb chains, as blst's add_mod_384 / sub_mod_384
 * assembly has them); the u128 forms are the portable fallback */
#if defined(__x86_64__) && defined(__GNUC__)
#include <x86intrin.h>
static inline void fp_cond_sub_p(fp_t *r, const uint64_t a[6], uint64_t hi) {
    unsigned long long t[6];
    unsigned char brw = 0;
    for (int i = 0; i < 6; i++) brw = _subborrow_u64(brw, a[i], FP_P.l[i], &t[i]);
    /* take t when no final borrow, or when the carry-out word `hi` absorbs it */
    uint64_t keep = (uint64_t)0 - (uint64_t)((hi == 0) & brw);
    for (int i = 0; i < 6; i++) r->l[i] = (a[i] & keep) | (t[i] & ~keep);
}
static inline void fp_add(fp_t *r, const fp_t *a, const fp_t *b) {
    unsigned long long t[6];
    unsigned char c = 0;
    for (int i = 0; i < 6; i++) c = _addcarry_u64(c, a->l[i], b->l[i], &t[i]);
    fp_cond_sub_p(r, (const uint64_t *)t, c);
}
static inline void fp_sub(fp_t *r, const fp_t *a, const fp_t *b) {
    unsigned long long t[6];
    unsigned char brw = 0;
    for (int i = 0; i < 6; i++) brw = _subborrow_u64(brw, a->l[i], b->l[i], &t[i]);
    uint64_t mask = (uint64_t)0 - (uint64_t)brw;
    unsigned char c = 0;
    for (int i = 0; i < 6; i++) c = _addcarry_u64(c, t[i], FP_P.l[i] & mask, &t[i]);
    for (int i = 0; i < 6; i++) r->l[i] = t[i];
}
#else
static inline void fp_cond_sub_p(fp_t *r, const uint64_t a[6], uint64_t hi) {
    uint64_t t[6];
    u128 brw = 0;
    for (int i = 0; i < 6; i++) {
        u128 d = (u128)a[i] - FP_P.l[i] - (uint64_t)brw;
        t[i] = (uint64_t)d;
        brw = (d >> 64) & 1;
    }
    int use_t = (hi != 0) || (brw == 0);
    for (int i = 0; i < 6; i++) r->l[i] = use_t ? t[i] : a[i];
}
static inline void fp_add(fp_t *r, const fp_t *a, const fp_t *b) {
    uint64_t t[6];
    u128 c = 0;
    for (int i = 0; i < 6; i++) {
        c += (u128)a->l[i] + b->l[i];
        t[i] = (uint64_t)c;
        c >>= 64;
    }
    fp_cond_sub_p(r, t, (uint64_t)c);
}
static inline void fp_sub(fp_t *r, const fp_t *a, const fp_t *b) {
    uint64_t t[6];
    u128 brw = 0;
    for (int i = 0; i < 6; i++) {
        u128 d = (u128)a->l[i] - b->l[i] - (uint64_t)brw;
        t[i] = (uint64_t)d;
        brw = (d >> 64) & 1;
    }
    if (brw) {
        u128 c = 0;
        for (int i = 0; i < 6; i++) {
            c += (u128)t[i] + FP_P.l[i];
            t[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    memcpy(r->l, t, sizeof t);
}
#endif
static inline void fp_neg(fp_t *r, const fp_t *a) {
    if (fp_is_zero(a)) { *r = *a; return; }
    fp_t z = {{0}};
    fp_sub(r, &z, a);
}
static inline void fp_dbl(fp_t *r, const fp_t *a) { fp_add(r, a, a); }

/* Montgomery product a·b·2^-384 mod p, CIOS over 64-bit limbs — portable form */
static inline void fp_mul_portable(fp_t *r, const fp_t *a, const fp_t *b) {
    uint64_t t[8] = {0};
    for (int i = 0; i < 6; i++) {
        u128 c = 0;
        for (int j = 0; j < 6; j++) {
            c += (u128)a->l[j] * b->l[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[6];
        t[6] = (uint64_t)c;
        t[7] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * FP_PINV;
        c = ((u128)m * FP_P.l[0] + t[0]) >> 64;
        for (int j = 1; j < 6; j++) {
            c += (u128)m * FP_P.l[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[6];
        t[5] = (uint64_t)c;
        t[6] = t[7] + (uint64_t)(c >> 64);
    }
    fp_cond_sub_p(r, t, t[6]);
}

/* The same product with MULX and the two independent ADCX / ADOX carry chains — the instruction
 * mix of blst's mulx_mont_384 (the x86-64 path blst =0.3.10 takes on a CPU with ADX + BMI2;
 * restated, not copied: blst is not vendored in the reference).  One row = t += a·b_i on the two
 * chains, then m = t0·(−p⁻¹), t += m·p, and the window slides down one limb by renaming registers.
 * p < 2^381 leaves the top limb's high bits clear, so a row never carries out of t6.
 * Chosen at run time (fp_mul below); limb-for-limb equal to fp_mul_portable (tests/test_oracle.py). */
#if defined(__x86_64__) && defined(__GNUC__)
#define FP_HAVE_ADX_PATH 1
#define FP_ADX_ROW(T0, T1, T2, T3, T4, T5, T6, SRC, MULT)                                   \
    __asm__("xorl %%eax, %%eax\n\t"              /* rax = 0, CF = OF = 0 */                  \
            "mulx 0(%[s]), %[lo], %[ha]\n\t"                                                 \
            "adcx %[lo], %[t0]\n\t"                                                          \
            "mulx 8(%[s]), %[lo], %[hb]\n\t"                                                 \
            "adox %[ha], %[t1]\n\t"                                                          \
            "adcx %[lo], %[t1]\n\t"                                                          \
            "mulx 16(%[s]), %[lo], %[ha]\n\t"                                                \
            "adox %[hb], %[t2]\n\t"                                                          \
            "adcx %[lo], %[t2]\n\t"                                                          \
            "mulx 24(%[s]), %[lo], %[hb]\n\t"                                                \
            "adox %[ha], %[t3]\n\t"                                                          \
            "adcx %[lo], %[t3]\n\t"                                                          \
            "mulx 32(%[s]), %[lo], %[ha]\n\t"                                                \
            "adox %[hb], %[t4]\n\t"                                                          \
            "adcx %[lo], %[t4]\n\t"                                                          \
            "mulx 40(%[s]), %[lo], %[hb]\n\t"                                                \
            "adox %[ha], %[t5]\n\t"                                                          \
            "adcx %[lo], %[t5]\n\t"                                                          \
            "adox %[hb], %[t6]\n\t"                                                          \
            "adcx %%rax, %[t6]\n\t"                                                          \
            : [t0] "+r"(T0), [t1] "+r"(T1), [t2] "+r"(T2), [t3] "+r"(T3), [t4] "+r"(T4),     \
              [t5] "+r"(T5), [t6] "+r"(T6), [lo] "=&r"(lo_), [ha] "=&r"(ha_), [hb] "=&r"(hb_) \
            : [s] "r"(SRC), "d"(MULT), "m"(*(const uint64_t(*)[6])(SRC))                    \
            : "rax", "cc")
static inline void fp_mul_adx(fp_t *r, const fp_t *a, const fp_t *b) {
    uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0, t6 = 0, t7, lo_, ha_, hb_;
    const uint64_t *pa = a->l, *pp = FP_P.l;
#define FP_ADX_STEP(A0, A1, A2, A3, A4, A5, A6, A7, BI)           \
    FP_ADX_ROW(A0, A1, A2, A3, A4, A5, A6, pa, BI);               \
    A7 = 0;                                                       \
    FP_ADX_ROW(A0, A1, A2, A3, A4, A5, A6, pp, A0 * FP_PINV)      /* A0 is now 0: the window is A1..A6, A7 = 0 */
    FP_ADX_STEP(t0, t1, t2, t3, t4, t5, t6, t7, b->l[0]);
    FP_ADX_STEP(t1, t2, t3, t4, t5, t6, t7, t0, b->l[1]);
    FP_ADX_STEP(t2, t3, t4, t5, t6, t7, t0, t1, b->l[2]);
    FP_ADX_STEP(t3, t4, t5, t6, t7, t0, t1, t2, b->l[3]);
    FP_ADX_STEP(t4, t5, t6, t7, t0, t1, t2, t3, b->l[4]);
    FP_ADX_STEP(t5, t6, t7, t0, t1, t2, t3, t4, b->l[5]);
#undef FP_ADX_STEP
    uint64_t t[6] = {t6, t7, t0, t1, t2, t3};
    fp_cond_sub_p(r, t, t4);
}
static int fp_adx_state = -1; /* -1 unknown, 0 portable, 1 ADX */
static inline int fp_use_adx(void) {
    if (__builtin_expect(fp_adx_state < 0, 0)) {
        __builtin_cpu_init();
        fp_adx_state = __builtin_cpu_supports("adx") && __builtin_cpu_supports("bmi2");
    }
    return fp_adx_state;
}
#else
#define FP_HAVE_ADX_PATH 0
static int fp_adx_state = 0;
static inline int fp_use_adx(void) { return 0; }
static inline void fp_mul_adx(fp_t *r, const fp_t *a, const fp_t *b) { fp_mul_portable(r, a, b); }
#endif
static inline void fp_mul(fp_t *r, const fp_t *a, const fp_t *b) {
    if (fp_use_adx()) fp_mul_adx(r, a, b);
    else fp_mul_portable(r, a, b);
}
static inline void fp_sqr(fp_t *r, const fp_t *a) { fp_mul(r, a, a); }
static inline void fp_to_mont(fp_t *r, const fp_t *a) { fp_mul(r, a, &FP_R2); }
static inline void fp_from_mont(fp_t *r, const fp_t *a) {
    fp_t one = {{1, 0, 0, 0, 0, 0}};
    fp_mul(r, a, &one);
}
/* a^(p-2) */
static inline void fp_inv(fp_t *r, const fp_t *a) {
    uint64_t e[6];
    memcpy(e, FP_P.l, sizeof e);
    e[0] -= 2; /* low limb ...aaab, no borrow */
    fp_t acc = FP_ONE, base = *a;
    for (int i = 0; i < 384; i++) {
        if ((e[i >> 6] >> (i & 63)) & 1) fp_mul(&acc, &acc, &base);
        fp_sqr(&base, &base);
    }
    *r = acc;
}

/* ------------------------------------------------------------------ Fp2 */
static inline int fp2_is_zero(const fp2_t *a) { return fp_is_zero(&a->c0) && fp_is_zero(&a->c1); }
static inline int fp2_eq(const fp2_t *a, const fp2_t *b) {
    return fp_eq(&a->c0, &b->c0) && fp_eq(&a->c1, &b->c1);
}
static inline void fp2_add(fp2_t *r, const fp2_t *a, const fp2_t *b) {
    fp_add(&r->c0, &a->c0, &b->c0);
    fp_add(&r->c1, &a->c1, &b->c1);
}
static inline void fp2_sub(fp2_t *r, const fp2_t *a, const fp2_t *b) {
    fp_sub(&r->c0, &a->c0, &b->c0);
    fp_sub(&r->c1, &a->c1, &b->c1);
}
static inline void fp2_neg(fp2_t *r, const fp2_t *a) {
    fp_neg(&r->c0, &a->c0);
    fp_neg(&r->c1, &a->c1);
}
static inline void fp2_dbl(fp2_t *r, const fp2_t *a) { fp2_add(r, a, a); }
/* Karatsuba, 3 Fp products */
static inline void fp2_mul(fp2_t *r, const fp2_t *a, const fp2_t *b) {
    fp_t aa, bb, sa, sb, m;
    fp_mul(&aa, &a->c0, &b->c0);
    fp_mul(&bb, &a->c1, &b->c1);
    fp_add(&sa, &a->c0, &a->c1);
    fp_add(&sb, &b->c0, &b->c1);
    fp_mul(&m, &sa, &sb);
    fp_sub(&m, &m, &aa);
    fp_sub(&r->c1, &m, &bb);
    fp_sub(&r->c0, &aa, &bb);
}
/* (a0+a1)(a0-a1), 2·a0·a1 : 2 Fp products */
static inline void fp2_sqr(fp2_t *r, const fp2_t *a) {
    fp_t s, d, m;
    fp_add(&s, &a->c0, &a->c1);
    fp_sub(&d, &a->c0, &a->c1);
    fp_mul(&m, &a->c0, &a->c1);
    fp_mul(&r->c0, &s, &d);
    fp_add(&r->c1, &m, &m);
}
static inline void fp2_inv(fp2_t *r, const fp2_t *a) {
    fp_t n, t;
    fp_sqr(&n, &a->c0);
    fp_sqr(&t, &a->c1);
    fp_add(&n, &n, &t);
    fp_inv(&n, &n);
    fp_mul(&r->c0, &a->c0, &n);
    fp_mul(&t, &a->c1, &n);
    fp_neg(&r->c1, &t);
}

/* ------------------------------------------------------------------ Fr (only what MSM needs) */
static inline int fr_geq_r(const uint64_t a[4]) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] > FR_R.l[i]) return 1;
        if (a[i] < FR_R.l[i]) return 0;
    }
    return 1;
}
static inline void fr_sub_r(uint64_t a[4]) {
    u128 brw = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a[i] - FR_R.l[i] - (uint64_t)brw;
        a[i] = (uint64_t)d;
        brw = (d >> 64) & 1;
    }
}
/* Montgomery product a·b·2^-256 mod r */
static inline void fr_mul(fr_t *r, const fr_t *a, const fr_t *b) {
    uint64_t t[6] = {0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a->l[j] * b->l[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * FR_RINV;
        c = ((u128)m * FR_R.l[0] + t[0]) >> 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)m * FR_R.l[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    if (t[4] || fr_geq_r(t)) fr_sub_r(t);
    memcpy(r->l, t, 32);
}
/* Montgomery Fr -> canonical integer (what blstrs Scalar::to_bytes_le / src/scalar.rs:450-463 do) */
static inline void fr_from_mont(fr_t *r, const fr_t *a) {
    fr_t one = {{1, 0, 0, 0}};
    fr_mul(r, a, &one);
}
/* any 256-bit integer -> [0, r): at most two subtractions since 2^256 < 3r */
static inline void fr_canon(fr_t *a) {
    while (fr_geq_r(a->l)) fr_sub_r(a->l);
}
static inline void fr_add(fr_t *r, const fr_t *a, const fr_t *b) {
    uint64_t t[4];
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
        c += (u128)a->l[i] + b->l[i];
        t[i] = (uint64_t)c;
        c >>= 64;
    }
    if (c || fr_geq_r(t)) fr_sub_r(t);
    memcpy(r->l, t, 32);
}
#endif
