// Quad-cooperative point arithmetic for the latency-bound tail of the MSM (bucket reduction levels,
// window combination, partial sums).
//
// There the work is a strictly serial chain of point operations (a running sum; 255 doublings of
// Horner's rule), so what matters is the latency of ONE operation, not throughput.  A point
// operation contains up to four independent field multiplications per dependency level
// (XYZZ add: 14 products in 4 levels, dbl: 9 products in 3 levels).  Four adjacent lanes of a
// warp (a "quad") therefore execute one operation together: every lane holds a replica of the
// operands, each lane computes one product of the current level, and the products are exchanged
// with width-4 shuffles.  The fma pipe issues per warp instruction regardless of how many lanes
// are live, so the three extra lanes cost nothing while the chain gets ~3× shorter.
//
// The main path is branch-free so that all quads of a warp stay converged and the shuffles are
// plain full-mask SHFL; exceptional operands (identity, P+P, P−P) are resolved afterwards by
// selects, the doubling case by the generic single-lane routine.
#pragma once
#include "ec.cuh"

namespace b200msm {

template <class F> struct quad_ops {
    static constexpr int W = field_words<F>::value;
    int q;  // lane index inside the quad
    __device__ __forceinline__ quad_ops() { q = threadIdx.x & 3; }
    // operand of lane q
    __device__ __forceinline__ F sel(const F &a0, const F &a1, const F &a2, const F &a3) const {
        F r;
#pragma unroll
        for (int i = 0; i < W; i++) {
            uint32_t lo = (q & 1) ? f_word(a1, i) : f_word(a0, i);
            uint32_t hi = (q & 1) ? f_word(a3, i) : f_word(a2, i);
            f_set_word(r, i, (q & 2) ? hi : lo);
        }
        return r;
    }
    // every lane multiplies its own (a, b); all four products come back to every lane
    __device__ __forceinline__ void mul4(F &r0, F &r1, F &r2, F &r3, const F &a, const F &b) const {
        F p;
        f_mul(p, a, b);
#pragma unroll
        for (int i = 0; i < W; i++) {
            uint32_t v = f_word(p, i);
            f_set_word(r0, i, __shfl_sync(0xffffffffu, v, 0, 4));
            f_set_word(r1, i, __shfl_sync(0xffffffffu, v, 1, 4));
            f_set_word(r2, i, __shfl_sync(0xffffffffu, v, 2, 4));
            f_set_word(r3, i, __shfl_sync(0xffffffffu, v, 3, 4));
        }
    }
    __device__ __forceinline__ void mul3(F &r0, F &r1, F &r2, const F &a, const F &b) const {
        F p;
        f_mul(p, a, b);
#pragma unroll
        for (int i = 0; i < W; i++) {
            uint32_t v = f_word(p, i);
            f_set_word(r0, i, __shfl_sync(0xffffffffu, v, 0, 4));
            f_set_word(r1, i, __shfl_sync(0xffffffffu, v, 1, 4));
            f_set_word(r2, i, __shfl_sync(0xffffffffu, v, 2, 4));
        }
    }
    __device__ __forceinline__ void mul2(F &r0, F &r1, const F &a, const F &b) const {
        F p;
        f_mul(p, a, b);
#pragma unroll
        for (int i = 0; i < W; i++) {
            uint32_t v = f_word(p, i);
            f_set_word(r0, i, __shfl_sync(0xffffffffu, v, 0, 4));
            f_set_word(r1, i, __shfl_sync(0xffffffffu, v, 1, 4));
        }
    }
};

template <class F> __device__ __forceinline__ void xyzz_select(xyzz<F> &dst, const xyzz<F> &src, bool take) {
    constexpr int W = field_words<F>::value;
#pragma unroll
    for (int i = 0; i < W; i++) {
        f_set_word(dst.x, i, take ? f_word(src.x, i) : f_word(dst.x, i));
        f_set_word(dst.y, i, take ? f_word(src.y, i) : f_word(dst.y, i));
        f_set_word(dst.zz, i, take ? f_word(src.zz, i) : f_word(dst.zz, i));
        f_set_word(dst.zzz, i, take ? f_word(src.zzz, i) : f_word(dst.zzz, i));
    }
}

// p = 2·p; the four lanes of the quad hold (and leave with) the same p. 9 products, 3 levels.
// Must be called by whole warps in convergence.
template <class F> __device__ __forceinline__ void xyzz_dbl_quad(xyzz<F> &p) {
    const bool to_inf = xyzz_is_inf(p) || f_is_zero(p.y);
    quad_ops<F> Q;
    F U, V, T, M, Wv, S, MM, ZZ3, t, d0, d1, d2;
    f_dbl(U, p.y);
    // level 1: V = U², T = X²
    Q.mul2(V, T, Q.sel(U, p.x, U, p.x), Q.sel(U, p.x, U, p.x));
    f_dbl(M, T);
    f_add(M, M, T);
    // level 2: W = U·V, S = X·V, MM = M², ZZ3 = V·ZZ
    Q.mul4(Wv, S, MM, ZZ3, Q.sel(U, p.x, M, V), Q.sel(V, V, M, p.zz));
    f_sub(t, MM, S);
    f_sub(t, t, S);            // X3
    f_sub(S, S, t);            // S - X3
    // level 3: M·(S-X3), W·Y, W·ZZZ
    Q.mul3(d0, d1, d2, Q.sel(M, Wv, Wv, Wv), Q.sel(S, p.y, p.zzz, p.zzz));
    p.x = t;
    f_sub(p.y, d0, d1);
    p.zz = ZZ3;
    p.zzz = d2;
    if (to_inf) xyzz_set_inf(p);
}

// acc += b, replicated across the quad. 14 products, 4 levels. Whole warps, converged.
template <class F> __device__ __forceinline__ void xyzz_add_quad(xyzz<F> &acc, const xyzz<F> &b) {
    const bool b_inf = xyzz_is_inf(b), a_inf = xyzz_is_inf(acc);
    quad_ops<F> Q;
    xyzz<F> r;
    F U1, U2, S1, S2, P, R, PP, RR, ZZm, ZZZm, PPP, Qv, d0, d1;
    // level 1: U1 = X1·ZZ2, U2 = X2·ZZ1, S1 = Y1·ZZZ2, S2 = Y2·ZZZ1
    Q.mul4(U1, U2, S1, S2, Q.sel(acc.x, b.x, acc.y, b.y), Q.sel(b.zz, acc.zz, b.zzz, acc.zzz));
    f_sub(P, U2, U1);
    f_sub(R, S2, S1);
    // level 2: PP = P², RR = R², ZZ1·ZZ2, ZZZ1·ZZZ2
    Q.mul4(PP, RR, ZZm, ZZZm, Q.sel(P, R, acc.zz, acc.zzz), Q.sel(P, R, b.zz, b.zzz));
    // level 3: PPP = P·PP, Q = U1·PP, ZZ3 = ZZm·PP
    Q.mul3(PPP, Qv, r.zz, Q.sel(P, U1, ZZm, ZZm), Q.sel(PP, PP, PP, PP));
    f_sub(r.x, RR, PPP);
    f_sub(r.x, r.x, Qv);
    f_sub(r.x, r.x, Qv);
    f_sub(Qv, Qv, r.x);
    // level 4: R·(Q-X3), S1·PPP, ZZZ3 = ZZZm·PPP
    Q.mul3(d0, d1, r.zzz, Q.sel(R, S1, ZZZm, ZZZm), Q.sel(Qv, PPP, PPP, PPP));
    f_sub(r.y, d0, d1);
    // exceptional operands
    const bool same_x = f_is_zero(P), same_y = f_is_zero(R);
    if (!a_inf && !b_inf && same_x) {          // quad-uniform, rare
        if (same_y) r = xyzz_dbl_val(acc);
        else xyzz_set_inf(r);
    }
    xyzz_select(r, b, a_inf);                  // ∞ + b = b
    xyzz_select(acc, r, !b_inf);               // acc + ∞ = acc
}

}  // namespace b200msm
