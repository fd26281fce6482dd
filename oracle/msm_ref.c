/* T1 oracle — CPU restatement of the reference's MSM path, and the timed CPU baseline ("port").
 * TEST INFRASTRUCTURE ONLY: loaded via ctypes by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs; never by ark_blst_b200/.
 *
 * What it follows in the reference:
 *   ref_g1_msm / ref_g2_msm   = <G?Projective as VariableBaseMSM>::msm, CPU arm
 *                               (src/g1.rs:602-619, src/g2.rs:582-599) → blstrs multi_exp → blst
 *                               p?s_mult_pippenger [un-vendored blst =0.3.10, Cargo.toml:22]:
 *                               window rule, Booth digits, XYZZ buckets, per-(window×slice) tiles
 *                               on a thread pool, Jacobian result.
 *   scalars_mont=1            = &[Scalar] as the trait passes them (Montgomery Fr, scalar.rs:23-25)
 *   scalars_mont=0            = &[BigInt<4>] as msm_bigint / the GPU arm passes them
 *                               (src/g1.rs:624-627, src/scalar.rs:458-463)
 *   identity bases            are accepted and skipped (the reference's blst arm mishandles them:
 *                               src/g1.rs:682-688; the arkworks arm src/g1.rs:690-693 defines the
 *                               intended result).
 * PARITY UNPINNED by reference vectors (none exist); pinned on oracle/bls12381.py (T0) and on
 * the reference's constants.  Build: oracle/Makefile → oracle/libmsm_ref.so
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "field.h"

/* ---- instantiate the curve layer for G1 (Fp) and G2 (Fp2) ---- */
static inline int32_t booth_digit(const uint64_t s[4], unsigned w, unsigned c);
#define BOOTH_DECLARED

#define F fp_t
#define FN(x) fp_##x
#define EC(x) g1_##x
#define F_ONE FP_ONE
#include "ec_tmpl.h"
#undef F
#undef FN
#undef EC
#undef F_ONE

static const fp2_t FP2_ONE = {{{0x760900000002fffdULL, 0xebf4000bc40c0002ULL, 0x5f48985753c758baULL,
                                0x77ce585370525745ULL, 0x5c071a97a256ec6dULL,
                                0x15f65ec3fa80e493ULL}},
                              {{0, 0, 0, 0, 0, 0}}};
#define F fp2_t
#define FN(x) fp2_##x
#define EC(x) g2_##x
#define F_ONE FP2_ONE
#include "ec_tmpl.h"
#undef F
#undef FN
#undef EC
#undef F_ONE

/* Booth digit of window `w` (width c) of canonical scalar s:  u + b[wc-1] - 2^c·b[wc+c-1],
 * in [-2^(c-1), 2^(c-1)]; Σ d_w·2^(wc) = s whenever bit W·c-1 of s is clear. */
static inline int32_t booth_digit(const uint64_t s[4], unsigned w, unsigned c) {
    unsigned lo = w * c;
    uint64_t v = 0; /* bits [lo-1, lo+c) → c+1 bits */
    for (unsigned k = 0; k <= c; k++) {
        int bit = (int)lo - 1 + (int)k;
        if (bit < 0 || bit >= 256) continue;
        v |= ((s[bit >> 6] >> (bit & 63)) & 1ULL) << k;
    }
    int32_t d = (int32_t)((v >> 1) & ((1u << c) - 1)) + (int32_t)(v & 1);
    if ((v >> c) & 1) d -= (int32_t)(1u << c);
    return d;
}

/* blst's pippenger_window_size (published rule): of log2(npoints) */
static unsigned window_rule(size_t n) {
    unsigned wbits = 0;
    while (n >>= 1) wbits++;
    return wbits > 12 ? wbits - 3 : wbits > 4 ? wbits - 2 : wbits ? 2 : 1;
}

/* blst's `breakdown(nbits, window, ncpus) -> (nx, ny, wnd)` [UPSTREAM-KNOWLEDGE: blst 0.3.10
 * bindings/rust/src/pippenger.rs, restated from its published logic; blst is not vendored in the
 * reference].  Few CPUs: one slice, the window nudged so the windows divide over the CPUs; many
 * CPUs: nx slices with the window narrowed by num_bits(3·nx/2). */
static unsigned num_bits(size_t l) {
    unsigned b = 0;
    while (l) { b++; l >>= 1; }
    return b;
}
static void breakdown(unsigned nbits, unsigned window, unsigned ncpus, unsigned *pnx, unsigned *pny,
                      unsigned *pwnd) {
    unsigned nx, wnd;
    if (nbits > window * ncpus) {
        nx = 1;
        wnd = num_bits(ncpus / 4);
        if (window + wnd > 18) wnd = window - wnd;
        else {
            wnd = (nbits / window + ncpus - 1) / ncpus;
            wnd = (nbits / (window + 1) + ncpus - 1) / ncpus < wnd ? window + 1 : window;
        }
    } else {
        nx = 2;
        wnd = window - 2;
        while ((nbits / wnd + 1) * nx < ncpus) {
            nx += 1;
            wnd = window - num_bits(3 * nx / 2);
        }
        nx -= 1;
        wnd = window - num_bits(3 * nx / 2);
    }
    unsigned ny = nbits / wnd + 1;
    wnd = nbits / ny + 1;
    *pnx = nx; *pny = ny; *pwnd = wnd;
}

/* ---- shared job description ------------------------------------------------------------- */
typedef struct {
    int g2;
    const void *bases;
    const fr_t *scalars; /* canonical */
    size_t n;
    unsigned c, nwin, nslice;
    void *tile_out; /* nwin × nslice xyzz */
    atomic_size_t next;
} job_t;

static void *worker(void *arg) {
    job_t *J = (job_t *)arg;
    size_t nb = (size_t)1 << (J->c - 1);
    size_t ntile = (size_t)J->nwin * J->nslice;
    void *buckets = malloc(nb * (J->g2 ? sizeof(g2_xyzz_t) : sizeof(g1_xyzz_t)));
    for (;;) {
        size_t t = atomic_fetch_add(&J->next, 1);
        if (t >= ntile) break;
        /* top windows first: they are the cheapest to finish late */
        unsigned w = (unsigned)(t / J->nslice), s = (unsigned)(t % J->nslice);
        size_t lo = J->n * s / J->nslice, hi = J->n * (s + 1) / J->nslice;
        if (J->g2)
            g2_tile(&((g2_xyzz_t *)J->tile_out)[t], (g2_xyzz_t *)buckets,
                    (const g2_aff_t *)J->bases, J->scalars, lo, hi, w, J->c);
        else
            g1_tile(&((g1_xyzz_t *)J->tile_out)[t], (g1_xyzz_t *)buckets,
                    (const g1_aff_t *)J->bases, J->scalars, lo, hi, w, J->c);
    }
    free(buckets);
    return NULL;
}

typedef struct {
    const fr_t *in;
    fr_t *out;
    size_t lo, hi;
    int mont;
} canon_t;
static void *canon_worker(void *arg) {
    canon_t *C = (canon_t *)arg;
    for (size_t i = C->lo; i < C->hi; i++) {
        if (C->mont) fr_from_mont(&C->out[i], &C->in[i]);
        else { C->out[i] = C->in[i]; fr_canon(&C->out[i]); }
    }
    return NULL;
}

static int msm_common(int g2, const void *bases, const uint64_t *scalars, size_t n, int mont,
                      uint64_t *out, int nthreads, int window) {
    size_t outw = g2 ? 36 : 18;
    memset(out, 0, outw * 8);
    if (n == 0) return 0;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];

    fr_t *canon = (fr_t *)malloc(n * sizeof(fr_t));
    if (!canon) return -1;
    {
        canon_t cj[256];
        for (int t = 0; t < nthreads; t++) {
            cj[t] = (canon_t){(const fr_t *)scalars, canon, n * t / nthreads,
                              n * (t + 1) / nthreads, mont};
            pthread_create(&th[t], NULL, canon_worker, &cj[t]);
        }
        for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    }

    job_t J;
    J.g2 = g2;
    J.bases = bases;
    J.scalars = canon;
    J.n = n;
    J.c = window > 0 ? (unsigned)window : window_rule(n);
    if (J.c > 24) J.c = 24;
    J.nslice = 1;
    /* the tile grid of blst's multi-threaded driver (bindings/rust/src/pippenger.rs, `breakdown`):
     * nx point slices × ny windows, the window narrowed as slices are added so that the bucket
     * reduction each tile repeats stays in proportion */
    if (nthreads > 1 && n >= 32 && window <= 0) {
        unsigned nx, ny, wnd;
        breakdown(255, J.c, (unsigned)nthreads, &nx, &ny, &wnd);
        if (wnd < 2) wnd = 2;
        J.c = wnd;
        J.nslice = nx;
    }
    J.nwin = (256 + J.c - 1) / J.c; /* W·c ≥ 256 > 255 = |r| keeps the top Booth carry inside */
    size_t ntile = (size_t)J.nwin * J.nslice;
    size_t xs = g2 ? sizeof(g2_xyzz_t) : sizeof(g1_xyzz_t);
    J.tile_out = malloc(ntile * xs);
    atomic_init(&J.next, 0);
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, worker, &J);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);

    /* Horner over windows, MSB first: acc = acc·2^c + Σ_slices tile */
    if (g2) {
        g2_xyzz_t acc;
        g2_xyzz_set_inf(&acc);
        for (unsigned w = J.nwin; w-- > 0;) {
            for (unsigned k = 0; k < J.c; k++) g2_xyzz_dbl(&acc, &acc);
            for (unsigned s = 0; s < J.nslice; s++)
                g2_xyzz_add(&acc, &((g2_xyzz_t *)J.tile_out)[(size_t)w * J.nslice + s]);
        }
        g2_jac_t r;
        g2_xyzz_to_jac(&r, &acc);
        memcpy(out, &r, sizeof r);
    } else {
        g1_xyzz_t acc;
        g1_xyzz_set_inf(&acc);
        for (unsigned w = J.nwin; w-- > 0;) {
            for (unsigned k = 0; k < J.c; k++) g1_xyzz_dbl(&acc, &acc);
            for (unsigned s = 0; s < J.nslice; s++)
                g1_xyzz_add(&acc, &((g1_xyzz_t *)J.tile_out)[(size_t)w * J.nslice + s]);
        }
        g1_jac_t r;
        g1_xyzz_to_jac(&r, &acc);
        memcpy(out, &r, sizeof r);
    }
    free(J.tile_out);
    free(canon);
    return 0;
}

/* ---- exported API (ctypes) -------------------------------------------------------------- */
#define API __attribute__((visibility("default")))

API int ref_g1_msm(const uint64_t *bases, const uint64_t *scalars, size_t n, int scalars_mont,
                   uint64_t out[18], int nthreads, int window) {
    return msm_common(0, bases, scalars, n, scalars_mont, out, nthreads, window);
}
API int ref_g2_msm(const uint64_t *bases, const uint64_t *scalars, size_t n, int scalars_mont,
                   uint64_t out[36], int nthreads, int window) {
    return msm_common(1, bases, scalars, n, scalars_mont, out, nthreads, window);
}

/* naive Σ sᵢ·Pᵢ by double-and-add — the literal shape of src/tests.rs:58-61 */
API void ref_g1_msm_naive(const uint64_t *bases, const uint64_t *scalars, size_t n, int mont,
                          uint64_t out[18]) {
    g1_xyzz_t acc, t;
    g1_xyzz_set_inf(&acc);
    for (size_t i = 0; i < n; i++) {
        fr_t s;
        memcpy(&s, scalars + 4 * i, 32);
        if (mont) fr_from_mont(&s, &s); else fr_canon(&s);
        g1_mul(&t, (const g1_aff_t *)bases + i, s.l);
        g1_xyzz_add(&acc, &t);
    }
    g1_jac_t r;
    g1_xyzz_to_jac(&r, &acc);
    memcpy(out, &r, sizeof r);
}
API void ref_g2_msm_naive(const uint64_t *bases, const uint64_t *scalars, size_t n, int mont,
                          uint64_t out[36]) {
    g2_xyzz_t acc, t;
    g2_xyzz_set_inf(&acc);
    for (size_t i = 0; i < n; i++) {
        fr_t s;
        memcpy(&s, scalars + 4 * i, 32);
        if (mont) fr_from_mont(&s, &s); else fr_canon(&s);
        g2_mul(&t, (const g2_aff_t *)bases + i, s.l);
        g2_xyzz_add(&acc, &t);
    }
    g2_jac_t r;
    g2_xyzz_to_jac(&r, &acc);
    memcpy(out, &r, sizeof r);
}

API void ref_g1_to_affine(const uint64_t jac[18], uint64_t aff[12]) {
    g1_jac_to_aff((g1_aff_t *)aff, (const g1_jac_t *)jac);
}
API void ref_g2_to_affine(const uint64_t jac[36], uint64_t aff[24]) {
    g2_jac_to_aff((g2_aff_t *)aff, (const g2_jac_t *)jac);
}
/* r = a + b on Jacobian inputs (final combine of per-GPU partials on the host, for tests) */
API void ref_g1_add(const uint64_t a[18], const uint64_t b[18], uint64_t r[18]) {
    g1_xyzz_t x, y;
    g1_jac_to_xyzz(&x, (const g1_jac_t *)a);
    g1_jac_to_xyzz(&y, (const g1_jac_t *)b);
    g1_xyzz_add(&x, &y);
    g1_xyzz_to_jac((g1_jac_t *)r, &x);
}
API void ref_g2_add(const uint64_t a[36], const uint64_t b[36], uint64_t r[36]) {
    g2_xyzz_t x, y;
    g2_jac_to_xyzz(&x, (const g2_jac_t *)a);
    g2_jac_to_xyzz(&y, (const g2_jac_t *)b);
    g2_xyzz_add(&x, &y);
    g2_xyzz_to_jac((g2_jac_t *)r, &x);
}
/* k·P, canonical k */
API void ref_g1_mul(const uint64_t aff[12], const uint64_t k[4], uint64_t out[18]) {
    g1_xyzz_t t;
    g1_mul(&t, (const g1_aff_t *)aff, k);
    g1_xyzz_to_jac((g1_jac_t *)out, &t);
}
API void ref_g2_mul(const uint64_t aff[24], const uint64_t k[4], uint64_t out[36]) {
    g2_xyzz_t t;
    g2_mul(&t, (const g2_aff_t *)aff, k);
    g2_xyzz_to_jac((g2_jac_t *)out, &t);
}

/* field unit entry points for limb-for-limb checks against the big-int oracle */
API void ref_fp_mul(const uint64_t a[6], const uint64_t b[6], uint64_t r[6]) {
    fp_mul((fp_t *)r, (const fp_t *)a, (const fp_t *)b);
}
API void ref_fp_add(const uint64_t a[6], const uint64_t b[6], uint64_t r[6]) {
    fp_add((fp_t *)r, (const fp_t *)a, (const fp_t *)b);
}
API void ref_fp_sub(const uint64_t a[6], const uint64_t b[6], uint64_t r[6]) {
    fp_sub((fp_t *)r, (const fp_t *)a, (const fp_t *)b);
}
API void ref_fp_inv(const uint64_t a[6], uint64_t r[6]) { fp_inv((fp_t *)r, (const fp_t *)a); }
API void ref_fp2_mul(const uint64_t a[12], const uint64_t b[12], uint64_t r[12]) {
    fp2_mul((fp2_t *)r, (const fp2_t *)a, (const fp2_t *)b);
}
API void ref_fp2_sqr(const uint64_t a[12], uint64_t r[12]) {
    fp2_sqr((fp2_t *)r, (const fp2_t *)a);
}
API void ref_fr_from_mont(const uint64_t a[4], uint64_t r[4]) {
    fr_from_mont((fr_t *)r, (const fr_t *)a);
}
API int ref_booth_digit(const uint64_t s[4], unsigned w, unsigned c) { return booth_digit(s, w, c); }
API unsigned ref_window_rule(size_t n) { return window_rule(n); }
API void ref_breakdown(unsigned nbits, unsigned window, unsigned ncpus, unsigned out[3]) {
    breakdown(nbits, window, ncpus, &out[0], &out[1], &out[2]);
}
/* which Montgomery product this build runs on this CPU: 1 = MULX/ADCX/ADOX, 0 = portable u128 */
API int ref_mul_impl(void) { return fp_use_adx(); }
API void ref_fp_mul_portable(const uint64_t a[6], const uint64_t b[6], uint64_t r[6]) {
    fp_mul_portable((fp_t *)r, (const fp_t *)a, (const fp_t *)b);
}
API void ref_fp_mul_adx(const uint64_t a[6], const uint64_t b[6], uint64_t r[6]) {
    fp_mul_adx((fp_t *)r, (const fp_t *)a, (const fp_t *)b);
}

/* ---- deterministic synthetic inputs (must match oracle/bls12381.py and csrc/synth.cuh) ---- */
static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static void synth_scalar(uint64_t seed, uint64_t i, fr_t *s) {
    for (int j = 0; j < 4; j++) s->l[j] = splitmix64(seed + 4 * i + (uint64_t)j);
    s->l[3] &= 0x7fffffffffffffffULL;
    if (fr_geq_r(s->l)) fr_sub_r(s->l);
}
static const fr_t FR_R2 = {{0xc999e990f3f29c6dULL, 0x2b6cedcb87925c23ULL, 0x05d314967254398fULL,
                            0x0748d9d99f59ff11ULL}}; /* 2^512 mod r, asserted in tests */
API void ref_synth_scalars(uint64_t seed, size_t n, int mont, uint64_t *out) {
    for (size_t i = 0; i < n; i++) {
        fr_t s;
        synth_scalar(seed, i, &s);
        if (mont) fr_mul(&s, &s, &FR_R2);
        memcpy(out + 4 * i, s.l, 32);
    }
}
API void ref_synth_dlogs(uint64_t seed, size_t n, uint64_t *out) {
    for (size_t i = 0; i < n; i++) {
        fr_t s;
        synth_scalar(seed ^ 0x5EEDBA5E5EEDBA5EULL, i, &s);
        if (!(s.l[0] | s.l[1] | s.l[2] | s.l[3])) s.l[0] = 1;
        memcpy(out + 4 * i, s.l, 32);
    }
}

typedef struct {
    int g2;
    size_t lo, hi;
    uint64_t *out;
    const void *table;
    const uint64_t *dlogs;
} synth_t;
static void *synth_worker(void *arg) {
    synth_t *S = (synth_t *)arg;
    if (S->g2)
        g2_fixed_base_range((g2_aff_t *)(S->out + 24 * S->lo), (const g2_aff_t *)S->table, S->dlogs, S->lo, S->hi);
    else
        g1_fixed_base_range((g1_aff_t *)(S->out + 12 * S->lo), (const g1_aff_t *)S->table, S->dlogs, S->lo, S->hi);
    return NULL;
}
/* bases Pᵢ = kᵢ·G (gen passed in as Montgomery affine limbs by the caller); kᵢ = ref_synth_dlogs */
API void ref_synth_bases(int g2, uint64_t seed, size_t n, const uint64_t *gen, uint64_t *out,
                         int nthreads) {
    if (n == 0) return;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    uint64_t *dlogs = (uint64_t *)malloc(n * 32);
    ref_synth_dlogs(seed, n, dlogs);
    void *table = malloc(32 * 255 * (g2 ? sizeof(g2_aff_t) : sizeof(g1_aff_t)));
    if (g2) g2_build_table((g2_aff_t *)table, (const g2_aff_t *)gen);
    else g1_build_table((g1_aff_t *)table, (const g1_aff_t *)gen);
    pthread_t th[256];
    synth_t sj[256];
    for (int t = 0; t < nthreads; t++) {
        sj[t] = (synth_t){g2, n * t / nthreads, n * (t + 1) / nthreads, out, table, dlogs};
        pthread_create(&th[t], NULL, synth_worker, &sj[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(table);
    free(dlogs);
}

/* T2 oracle helper: Σ sᵢ·kᵢ mod r over the synthetic streams (canonical in, canonical out) */
API void ref_fr_dot_synth(uint64_t seed_bases, const uint64_t *scalars_canon, size_t n,
                          uint64_t out[4]) {
    fr_t acc = {{0, 0, 0, 0}};
    for (size_t i = 0; i < n; i++) {
        fr_t s, k, km, prod;
        memcpy(&s, scalars_canon + 4 * i, 32);
        synth_scalar(seed_bases ^ 0x5EEDBA5E5EEDBA5EULL, i, &k);
        if (!(k.l[0] | k.l[1] | k.l[2] | k.l[3])) k.l[0] = 1;
        fr_mul(&km, &k, &FR_R2);  /* k·R */
        fr_mul(&prod, &s, &km);   /* s·k·R·R^-1 = s·k */
        fr_add(&acc, &acc, &prod);
    }
    memcpy(out, acc.l, 32);
}
