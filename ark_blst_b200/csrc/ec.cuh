// Short-Weierstrass a = 0 curve arithmetic (BLS12-381 G1 over Fp, G2 over Fp2), field-generic.
//
//   affine<F>  = blst_p{1,2}_affine  (x, y; infinity = all-zero)     reference src/g1.rs:54-56
//   jac<F>     = blst_p{1,2}         (X, Y, Z; infinity iff Z = 0)   reference src/g1.rs:435-437
//   xyzz<F>    = bucket form (X, Y, ZZ, ZZZ) with x = X/ZZ, y = Y/ZZZ, ZZ³ = ZZZ²; inf iff ZZ = 0
//
// Formulas: EFD "xyzz" madd-2008-s (8M+2S), add-2008-s (12M+2S), dbl-2008-s-1 (6M+3S) — the
// operation counts SURVEY §8(d)'s work model uses.  Every exceptional case (identity operand,
// P+P, P+(−P)) is handled, so results are exact for any input, including the identity bases the
// reference's blst arm mishandles (src/g1.rs:682-688).
#pragma once
#include "fp2.cuh"

namespace b200msm {

template <class F> struct affine { F x, y; };
template <class F> struct jac { F x, y, z; };
template <class F> struct xyzz { F x, y, zz, zzz; };

template <class F> __device__ __forceinline__ void xyzz_set_inf(xyzz<F> &p) {
    f_set_zero(p.x); f_set_zero(p.y); f_set_zero(p.zz); f_set_zero(p.zzz);
}
template <class F> __device__ __forceinline__ bool xyzz_is_inf(const xyzz<F> &p) { return f_is_zero(p.zz); }
template <class F> __device__ __forceinline__ bool affine_is_inf(const affine<F> &p) {
    return f_is_zero(p.x) && f_is_zero(p.y);
}

// r = 2·(x, y), affine input not at infinity
template <class F> __device__ __forceinline__ void xyzz_mdbl(xyzz<F> &r, const F &x, const F &y) {
    if (f_is_zero(y)) { xyzz_set_inf(r); return; }  // order-2 point (none in the r-torsion)
    F U, S, M, t;
    f_dbl(U, y);
    f_sqr(r.zz, U);            // V
    f_mul(r.zzz, U, r.zz);     // W
    f_mul(S, x, r.zz);
    f_sqr(t, x);
    f_dbl(M, t);
    f_add(M, M, t);
    f_sqr(r.x, M);
    f_sub(r.x, r.x, S);
    f_sub(r.x, r.x, S);
    f_sub(t, S, r.x);
    f_mul(t, M, t);
    f_mul(U, r.zzz, y);
    f_sub(r.y, t, U);
}

// p = 2·p
template <class F> __device__ __forceinline__ void xyzz_dbl(xyzz<F> &p) {
    if (xyzz_is_inf(p) || f_is_zero(p.y)) { xyzz_set_inf(p); return; }
    F U, V, W, S, M, t;
    f_dbl(U, p.y);
    f_sqr(V, U);
    f_mul(W, U, V);
    f_mul(S, p.x, V);
    f_sqr(t, p.x);
    f_dbl(M, t);
    f_add(M, M, t);
    f_sqr(p.x, M);
    f_sub(p.x, p.x, S);
    f_sub(p.x, p.x, S);
    f_sub(t, S, p.x);
    f_mul(t, M, t);
    f_mul(U, W, p.y);
    f_sub(p.y, t, U);
    f_mul(p.zz, V, p.zz);
    f_mul(p.zzz, W, p.zzz);
}

// rare path of madd, out of line and by value so the accumulator never has its address taken
template <class F> __device__ __noinline__ xyzz<F> xyzz_mdbl_val(F x, F y) {
    xyzz<F> r;
    xyzz_mdbl(r, x, y);
    return r;
}

// acc += (qx, qy); q affine, not at infinity (callers filter identity bases)
template <class F> __device__ __forceinline__ void xyzz_madd(xyzz<F> &acc, const F &qx, const F &qy) {
    if (xyzz_is_inf(acc)) {
        acc.x = qx; acc.y = qy;
        f_set_one(acc.zz); f_set_one(acc.zzz);
        return;
    }
    F P, R, PP, PPP, Q;
    f_mul(P, qx, acc.zz);      // U2
    f_mul(R, qy, acc.zzz);     // S2
    f_sub(P, P, acc.x);
    f_sub(R, R, acc.y);
    if (f_is_zero(P)) {        // same x: doubling or cancellation (rare, data-dependent)
        if (f_is_zero(R)) acc = xyzz_mdbl_val(qx, qy);
        else xyzz_set_inf(acc);
        return;
    }
    f_sqr(PP, P);
    f_mul(PPP, P, PP);
    f_mul(Q, acc.x, PP);
    f_sqr(acc.x, R);
    f_sub(acc.x, acc.x, PPP);
    f_sub(acc.x, acc.x, Q);
    f_sub(acc.x, acc.x, Q);
    f_sub(Q, Q, acc.x);
    f_mul(Q, R, Q);
    f_mul(acc.y, acc.y, PPP);
    f_sub(acc.y, Q, acc.y);
    f_mul(acc.zz, acc.zz, PP);
    f_mul(acc.zzz, acc.zzz, PPP);
}

template <class F> __device__ __noinline__ xyzz<F> xyzz_dbl_val(xyzz<F> p) {
    xyzz_dbl(p);
    return p;
}

// acc += b
template <class F> __device__ __forceinline__ void xyzz_add(xyzz<F> &acc, const xyzz<F> &b) {
    if (xyzz_is_inf(b)) return;
    if (xyzz_is_inf(acc)) { acc = b; return; }
    F U1, S1, P, R, PP, PPP, Q;
    f_mul(U1, acc.x, b.zz);
    f_mul(P, b.x, acc.zz);     // U2
    f_mul(S1, acc.y, b.zzz);
    f_mul(R, b.y, acc.zzz);    // S2
    f_sub(P, P, U1);
    f_sub(R, R, S1);
    if (f_is_zero(P)) {
        if (f_is_zero(R)) acc = xyzz_dbl_val(acc);
        else xyzz_set_inf(acc);
        return;
    }
    f_sqr(PP, P);
    f_mul(PPP, P, PP);
    f_mul(Q, U1, PP);
    f_sqr(acc.x, R);
    f_sub(acc.x, acc.x, PPP);
    f_sub(acc.x, acc.x, Q);
    f_sub(acc.x, acc.x, Q);
    f_sub(Q, Q, acc.x);
    f_mul(Q, R, Q);
    f_mul(S1, S1, PPP);
    f_sub(acc.y, Q, S1);
    f_mul(acc.zz, acc.zz, b.zz);
    f_mul(acc.zz, acc.zz, PP);
    f_mul(acc.zzz, acc.zzz, b.zzz);
    f_mul(acc.zzz, acc.zzz, PPP);
}

// (X·ZZ, Y·ZZZ, ZZ) is the same point in Jacobian coordinates (Z = ZZ)
template <class F> __device__ __forceinline__ void xyzz_to_jac(jac<F> &r, const xyzz<F> &p) {
    if (xyzz_is_inf(p)) { f_set_zero(r.x); f_set_zero(r.y); f_set_zero(r.z); return; }
    f_mul(r.x, p.x, p.zz);
    f_mul(r.y, p.y, p.zzz);
    r.z = p.zz;
}
template <class F> __device__ __forceinline__ void jac_to_xyzz(xyzz<F> &r, const jac<F> &p) {
    if (f_is_zero(p.z)) { xyzz_set_inf(r); return; }
    r.x = p.x; r.y = p.y;
    f_sqr(r.zz, p.z);
    f_mul(r.zzz, r.zz, p.z);
}
template <class F> __device__ __forceinline__ void xyzz_to_affine(affine<F> &r, const xyzz<F> &p) {
    if (xyzz_is_inf(p)) { f_set_zero(r.x); f_set_zero(r.y); return; }
    F zi, t;                   // 1/ZZZ → 1/ZZ = ZZ²/ZZZ² · ... use: x = X·ZZ²/ZZZ² (ZZ³ = ZZZ²)
    f_inv(zi, p.zzz);
    f_mul(r.y, p.y, zi);
    f_sqr(t, zi);              // 1/ZZZ² = 1/ZZ³
    f_mul(t, t, p.zz);
    f_mul(t, t, p.zz);         // 1/ZZ
    f_mul(r.x, p.x, t);
}

// ---- out-of-line copies for the kernels where point operations are not the bottleneck ----
template <class F> __device__ __noinline__ void xyzz_add_ni(xyzz<F> &acc, const xyzz<F> &b) { xyzz_add(acc, b); }
template <class F> __device__ __noinline__ void xyzz_dbl_ni(xyzz<F> &p) { xyzz_dbl(p); }
template <class F> __device__ __noinline__ void xyzz_madd_ni(xyzz<F> &acc, const F &qx, const F &qy) { xyzz_madd(acc, qx, qy); }

// ---- global-memory I/O in the reference's layouts (u32 view of the u64 limb arrays) ----
template <class F> __device__ __forceinline__ void xyzz_load(xyzz<F> &r, const uint32_t *p) {
    constexpr int W = field_words<F>::value;
    f_load(r.x, p); f_load(r.y, p + W); f_load(r.zz, p + 2 * W); f_load(r.zzz, p + 3 * W);
}
template <class F> __device__ __forceinline__ void xyzz_store(uint32_t *p, const xyzz<F> &a) {
    constexpr int W = field_words<F>::value;
    f_store(p, a.x); f_store(p + W, a.y); f_store(p + 2 * W, a.zz); f_store(p + 3 * W, a.zzz);
}
template <class F> __device__ __forceinline__ void jac_store(uint32_t *p, const jac<F> &a) {
    constexpr int W = field_words<F>::value;
    f_store(p, a.x); f_store(p + W, a.y); f_store(p + 2 * W, a.z);
}
template <class F> __device__ __forceinline__ void jac_load(jac<F> &r, const uint32_t *p) {
    constexpr int W = field_words<F>::value;
    f_load(r.x, p); f_load(r.y, p + W); f_load(r.z, p + 2 * W);
}

}  // namespace b200msm
