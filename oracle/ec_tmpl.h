/* T1 oracle — curve layer, textually instantiated twice (Fp → G1, Fp2 → G2) by msm_ref.c.
 * TEST INFRASTRUCTURE ONLY.  Restates the published algorithm shape of blst 0.3.10's Pippenger
 * (un-vendored, Cargo.toml:22; reached from the reference at src/g1.rs:614-617 and
 * src/g2.rs:594-597): Booth-signed c-bit digits, XYZZ buckets fed by mixed additions of affine
 * points, running-sum integration of the buckets, Horner combination of the windows, Jacobian
 * result.  Formulas are the EFD "xyzz" set for a = 0 (madd-2008-s, add-2008-s, dbl-2008-s-1).
 * The value computed is the reference's definition  Σ sᵢ·Pᵢ  (src/tests.rs:58-61).
 *
 * Required macros:  F (field type)  FN(x) (field fn prefix)  EC(x) (curve symbol prefix)
 *                   F_ONE (Montgomery one initialiser expr)
 */

typedef struct { F x, y; } EC(aff_t);          /* blst_p?_affine: infinity = all-zero  */
typedef struct { F x, y, z; } EC(jac_t);       /* blst_p?: infinity iff z == 0         */
typedef struct { F x, y, zz, zzz; } EC(xyzz_t); /* bucket form: infinity iff zz == 0   */

static inline int EC(aff_is_inf)(const EC(aff_t) * p) {
    return FN(is_zero)(&p->x) && FN(is_zero)(&p->y);
}
static inline void EC(xyzz_set_inf)(EC(xyzz_t) * p) { memset(p, 0, sizeof *p); }
static inline int EC(xyzz_is_inf)(const EC(xyzz_t) * p) { return FN(is_zero)(&p->zz); }

/* p = 2·(x,y) */
static void EC(xyzz_mdbl)(EC(xyzz_t) * r, const F *x, const F *y) {
    F U, V, W, S, M, t;
    FN(dbl)(&U, y);
    FN(sqr)(&V, &U);
    FN(mul)(&W, &U, &V);
    FN(mul)(&S, x, &V);
    FN(sqr)(&t, x);
    FN(dbl)(&M, &t);
    FN(add)(&M, &M, &t);
    FN(sqr)(&r->x, &M);
    FN(sub)(&r->x, &r->x, &S);
    FN(sub)(&r->x, &r->x, &S);
    FN(sub)(&t, &S, &r->x);
    FN(mul)(&t, &M, &t);
    FN(mul)(&U, &W, y);
    FN(sub)(&r->y, &t, &U);
    r->zz = V;
    r->zzz = W;
}
/* r = 2·p */
static void EC(xyzz_dbl)(EC(xyzz_t) * r, const EC(xyzz_t) * p) {
    if (EC(xyzz_is_inf)(p) || FN(is_zero)(&p->y)) { EC(xyzz_set_inf)(r); return; }
    F U, V, W, S, M, t, X3;
    FN(dbl)(&U, &p->y);
    FN(sqr)(&V, &U);
    FN(mul)(&W, &U, &V);
    FN(mul)(&S, &p->x, &V);
    FN(sqr)(&t, &p->x);
    FN(dbl)(&M, &t);
    FN(add)(&M, &M, &t);
    FN(sqr)(&X3, &M);
    FN(sub)(&X3, &X3, &S);
    FN(sub)(&X3, &X3, &S);
    FN(sub)(&t, &S, &X3);
    FN(mul)(&t, &M, &t);
    FN(mul)(&U, &W, &p->y);
    FN(sub)(&r->y, &t, &U);
    FN(mul)(&r->zz, &V, &p->zz);
    FN(mul)(&r->zzz, &W, &p->zzz);
    r->x = X3;
}
/* acc += (x, ±y) ; neg selects -y.  Handles acc = inf, equal and opposite points. */
static void EC(xyzz_madd)(EC(xyzz_t) * acc, const EC(aff_t) * q, int neg) {
    if (EC(aff_is_inf)(q)) return;
    F qy = q->y;
    if (neg) FN(neg)(&qy, &qy);
    if (EC(xyzz_is_inf)(acc)) {
        acc->x = q->x;
        acc->y = qy;
        acc->zz = F_ONE;
        acc->zzz = F_ONE;
        return;
    }
    F U2, S2, Pp, R, PP, PPP, Q, t, X3;
    FN(mul)(&U2, &q->x, &acc->zz);
    FN(mul)(&S2, &qy, &acc->zzz);
    FN(sub)(&Pp, &U2, &acc->x);
    FN(sub)(&R, &S2, &acc->y);
    if (FN(is_zero)(&Pp)) {
        if (FN(is_zero)(&R)) EC(xyzz_mdbl)(acc, &q->x, &qy);
        else EC(xyzz_set_inf)(acc);
        return;
    }
    FN(sqr)(&PP, &Pp);
    FN(mul)(&PPP, &Pp, &PP);
    FN(mul)(&Q, &acc->x, &PP);
    FN(sqr)(&X3, &R);
    FN(sub)(&X3, &X3, &PPP);
    FN(sub)(&X3, &X3, &Q);
    FN(sub)(&X3, &X3, &Q);
    FN(sub)(&t, &Q, &X3);
    FN(mul)(&t, &R, &t);
    FN(mul)(&Q, &acc->y, &PPP);
    FN(sub)(&acc->y, &t, &Q);
    FN(mul)(&acc->zz, &acc->zz, &PP);
    FN(mul)(&acc->zzz, &acc->zzz, &PPP);
    acc->x = X3;
}
/* acc += b */
static void EC(xyzz_add)(EC(xyzz_t) * acc, const EC(xyzz_t) * b) {
    if (EC(xyzz_is_inf)(b)) return;
    if (EC(xyzz_is_inf)(acc)) { *acc = *b; return; }
    F U1, U2, S1, S2, Pp, R, PP, PPP, Q, t, X3;
    FN(mul)(&U1, &acc->x, &b->zz);
    FN(mul)(&U2, &b->x, &acc->zz);
    FN(mul)(&S1, &acc->y, &b->zzz);
    FN(mul)(&S2, &b->y, &acc->zzz);
    FN(sub)(&Pp, &U2, &U1);
    FN(sub)(&R, &S2, &S1);
    if (FN(is_zero)(&Pp)) {
        if (FN(is_zero)(&R)) EC(xyzz_dbl)(acc, acc);
        else EC(xyzz_set_inf)(acc);
        return;
    }
    FN(sqr)(&PP, &Pp);
    FN(mul)(&PPP, &Pp, &PP);
    FN(mul)(&Q, &U1, &PP);
    FN(sqr)(&X3, &R);
    FN(sub)(&X3, &X3, &PPP);
    FN(sub)(&X3, &X3, &Q);
    FN(sub)(&X3, &X3, &Q);
    FN(sub)(&t, &Q, &X3);
    FN(mul)(&t, &R, &t);
    FN(mul)(&Q, &S1, &PPP);
    FN(sub)(&acc->y, &t, &Q);
    FN(mul)(&t, &acc->zz, &b->zz);
    FN(mul)(&acc->zz, &t, &PP);
    FN(mul)(&t, &acc->zzz, &b->zzz);
    FN(mul)(&acc->zzz, &t, &PPP);
    acc->x = X3;
}
/* (X·ZZ, Y·ZZZ, ZZ) is the same point in Jacobian coordinates */
static void EC(xyzz_to_jac)(EC(jac_t) * r, const EC(xyzz_t) * p) {
    if (EC(xyzz_is_inf)(p)) { memset(r, 0, sizeof *r); return; }
    FN(mul)(&r->x, &p->x, &p->zz);
    FN(mul)(&r->y, &p->y, &p->zzz);
    r->z = p->zz;
}
static void EC(jac_to_xyzz)(EC(xyzz_t) * r, const EC(jac_t) * p) {
    if (FN(is_zero)(&p->z)) { EC(xyzz_set_inf)(r); return; }
    r->x = p->x;
    r->y = p->y;
    FN(sqr)(&r->zz, &p->z);
    FN(mul)(&r->zzz, &r->zz, &p->z);
}
static void EC(jac_to_aff)(EC(aff_t) * r, const EC(jac_t) * p) {
    if (FN(is_zero)(&p->z)) { memset(r, 0, sizeof *r); return; }
    F zi, zi2;
    FN(inv)(&zi, &p->z);
    FN(sqr)(&zi2, &zi);
    FN(mul)(&r->x, &p->x, &zi2);
    FN(mul)(&zi2, &zi2, &zi);
    FN(mul)(&r->y, &p->y, &zi2);
}
/* k·q for a canonical 256-bit k, MSB-first double-and-add (shape of src/g1.rs:331-341) */
static void EC(mul)(EC(xyzz_t) * r, const EC(aff_t) * q, const uint64_t k[4]) {
    EC(xyzz_t) acc;
    EC(xyzz_set_inf)(&acc);
    for (int i = 255; i >= 0; i--) {
        EC(xyzz_dbl)(&acc, &acc);
        if ((k[i >> 6] >> (i & 63)) & 1) EC(xyzz_madd)(&acc, q, 0);
    }
    *r = acc;
}

/* ---- Pippenger over one (window, slice) tile -------------------------------------------- */
static void EC(tile)(EC(xyzz_t) * out, EC(xyzz_t) * buckets, const EC(aff_t) * bases,
                     const fr_t *scalars, size_t lo, size_t hi, unsigned w, unsigned c) {
    size_t nb = (size_t)1 << (c - 1);
    memset(buckets, 0, nb * sizeof *buckets);
    for (size_t i = lo; i < hi; i++) {
        int32_t d = booth_digit(scalars[i].l, w, c);
        if (d > 0) EC(xyzz_madd)(&buckets[d - 1], &bases[i], 0);
        else if (d < 0) EC(xyzz_madd)(&buckets[-d - 1], &bases[i], 1);
    }
    /* Σ_b b·bucket[b] by running sum from the top bucket down */
    EC(xyzz_t) run, acc;
    EC(xyzz_set_inf)(&run);
    EC(xyzz_set_inf)(&acc);
    for (size_t b = nb; b-- > 0;) {
        EC(xyzz_add)(&run, &buckets[b]);
        EC(xyzz_add)(&acc, &run);
    }
    *out = acc;
}

/* ---- fast synthetic bases: k·G by an 8-bit fixed-base table + batched normalisation -------- */
/* out[i] = affine(in[i]) for XYZZ inputs, one field inversion per call (Montgomery's trick).
 * x = X/ZZ, y = Y/ZZZ with 1/ZZ = ZZ²·(1/ZZZ)² since ZZ³ = ZZZ². */
static void EC(batch_to_aff)(EC(aff_t) * out, const EC(xyzz_t) * in, size_t n, F *scratch) {
    F acc = F_ONE;
    for (size_t i = 0; i < n; i++) {
        scratch[i] = acc;
        if (!EC(xyzz_is_inf)(&in[i])) FN(mul)(&acc, &acc, &in[i].zzz);
    }
    F inv;
    FN(inv)(&inv, &acc);
    for (size_t i = n; i-- > 0;) {
        if (EC(xyzz_is_inf)(&in[i])) { memset(&out[i], 0, sizeof out[i]); continue; }
        F zi, t;
        FN(mul)(&zi, &inv, &scratch[i]);      /* 1/ZZZ_i */
        FN(mul)(&inv, &inv, &in[i].zzz);
        FN(mul)(&out[i].y, &in[i].y, &zi);
        FN(sqr)(&t, &zi);
        FN(mul)(&t, &t, &in[i].zz);
        FN(mul)(&t, &t, &in[i].zz);           /* 1/ZZ_i */
        FN(mul)(&out[i].x, &in[i].x, &t);
    }
}
/* table[j*255 + d-1] = d·2^(8j)·G, j < 32, 1 ≤ d ≤ 255 */
static void EC(build_table)(EC(aff_t) * table, const EC(aff_t) * gen) {
    size_t n = 32 * 255;
    EC(xyzz_t) *t = (EC(xyzz_t) *)malloc(n * sizeof *t);
    F *scratch = (F *)malloc(n * sizeof(F));
    EC(xyzz_t) base;
    EC(xyzz_set_inf)(&base);
    EC(xyzz_madd)(&base, gen, 0);
    for (int j = 0; j < 32; j++) {
        EC(xyzz_t) acc = base;
        for (int d = 1; d <= 255; d++) {
            t[j * 255 + d - 1] = acc;
            EC(xyzz_add)(&acc, &base);
        }
        for (int k = 0; k < 8; k++) EC(xyzz_dbl)(&base, &base);
    }
    EC(batch_to_aff)(table, t, n, scratch);
    free(t);
    free(scratch);
}
/* out[i - lo] for i in [lo, hi): k_i·G with k_i canonical little-endian bytes from `dlogs` */
static void EC(fixed_base_range)(EC(aff_t) * out, const EC(aff_t) * table, const uint64_t *dlogs, size_t lo,
                                 size_t hi) {
    enum { CH = 512 };
    EC(xyzz_t) buf[CH];
    F scratch[CH];
    for (size_t s = lo; s < hi; s += CH) {
        size_t m = hi - s < CH ? hi - s : CH;
        for (size_t i = 0; i < m; i++) {
            const uint8_t *kb = (const uint8_t *)(dlogs + 4 * (s + i));
            EC(xyzz_set_inf)(&buf[i]);
            for (int j = 0; j < 32; j++)
                if (kb[j]) EC(xyzz_madd)(&buf[i], &table[j * 255 + kb[j] - 1], 0);
        }
        EC(batch_to_aff)(out + (s - lo), buf, m, scratch);
    }
}

