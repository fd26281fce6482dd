// Quad-distributed point arithmetic for the latency-bound tail of the MSM (bucket reduction levels,
// reduction tree, window combination, partial sums).
//
// There the work is a chain of dependent point operations (a running sum; 255 doublings of
// Horner's rule), so what matters is the latency of ONE operation.  A point operation has up to
// four independent field products per dependency level (XYZZ add: 14 products in 4 levels, dbl:
// 9 products in 3 levels).  Four adjacent lanes of a warp (a "quad") therefore execute one
// operation together, and each lane OWNS one coordinate of every point variable:
//        lane 0: X      lane 1: Y      lane 2: ZZ      lane 3: ZZZ
// Per level every lane computes one product; operands that live on another lane arrive by
// width-4 shuffles (6 field-element exchanges per add, 5 per dbl).  Compared with replicating
// the whole point on all four lanes this keeps a quarter of the state per thread (no spills for
// G1, far fewer for G2) and loads/stores become naturally coalesced (lane q moves coordinate q).
// The fma pipe issues per warp instruction regardless of how many lanes do useful work, so the
// extra lanes are free while the chain gets ~3× shorter.
//
// The main path is branch-free, so all quads of a warp stay converged and the shuffles are plain
// full-mask SHFL; exceptional operands (identity, P+P, P−P) are resolved afterwards by selects,
// the doubling case by the generic single-lane routine on a gathered copy.  Every function here
// must be called by whole, converged warps.
#pragma once
#include "ec.cuh"

namespace b200msm {

constexpr unsigned QFULL = 0xffffffffu;

template <class F> __device__ __forceinline__ F q_shfl(const F &v, int src) {   // from lane `src` of the quad
    F r;
#pragma unroll
    for (int i = 0; i < field_words<F>::value; i++) f_set_word(r, i, __shfl_sync(QFULL, f_word(v, i), src, 4));
    return r;
}
template <class F> __device__ __forceinline__ F q_shfl_xor(const F &v, int mask) {
    F r;
#pragma unroll
    for (int i = 0; i < field_words<F>::value; i++) f_set_word(r, i, __shfl_xor_sync(QFULL, f_word(v, i), mask, 4));
    return r;
}
template <class F> __device__ __forceinline__ F q_sel(bool c, const F &a, const F &b) {  // c ? a : b
    F r;
#pragma unroll
    for (int i = 0; i < field_words<F>::value; i++) f_set_word(r, i, c ? f_word(a, i) : f_word(b, i));
    return r;
}
__device__ __forceinline__ bool q_flag(bool f, int src) { return __shfl_sync(QFULL, (int)f, src, 4) != 0; }

// lane q's coordinate of a point stored as XYZZ in memory
template <class F> __device__ __forceinline__ void q_load(F &c, const uint32_t *p) {
    f_load(c, p + (threadIdx.x & 3) * field_words<F>::value);
}
template <class F> __device__ __forceinline__ void q_store(uint32_t *p, const F &c) {
    f_store(p + (threadIdx.x & 3) * field_words<F>::value, c);
}
// gather the four coordinates onto every lane of THIS quad / take this lane's coordinate back.
// Only for the rare exceptional paths, which are quad-uniform but diverge from the rest of the
// warp: the shuffles therefore name just the four lanes of the quad in their mask.
template <class F> __device__ __noinline__ void q_gather(xyzz<F> &p, const F &c) {
    const unsigned mask = 0xFu << (threadIdx.x & 28);
    F *dst[4] = {&p.x, &p.y, &p.zz, &p.zzz};
#pragma unroll
    for (int k = 0; k < 4; k++) {
#pragma unroll
        for (int i = 0; i < field_words<F>::value; i++) f_set_word(*dst[k], i, __shfl_sync(mask, f_word(c, i), k, 4));
    }
}
template <class F> __device__ __forceinline__ F q_own(const xyzz<F> &p) {
    const int q = threadIdx.x & 3;
    return q_sel(q < 2, q_sel(q == 0, p.x, p.y), q_sel(q == 2, p.zz, p.zzz));
}
template <class F> __device__ __forceinline__ void q_set_inf(F &c) { f_set_zero(c); }

// A = 2·A.  9 products in 3 levels.
template <class F> __device__ __forceinline__ void q_dbl(F &A) {
    const int q = threadIdx.x & 3;
    // ZZ = 0 or Y = 0. Both shuffles must run on every lane: no short-circuit between them.
    const bool own_zero = f_is_zero(A);
    const bool zz_zero = q_flag(own_zero, 2), y_zero = q_flag(own_zero, 1);
    const bool to_inf = zz_zero | y_zero;
    F U, M1, M2, M3, t;
    f_dbl(U, A);                                   // meaningful on lane 1: U = 2Y
    // level 1: lane0 T = X², lane1 V = U²
    t = q_sel(q == 1, U, A);
    f_mul(M1, t, t);
    const F Vb = q_shfl(M1, 1);
    f_dbl(t, M1);
    f_add(t, t, M1);                               // lane 0: M = 3X²
    const F Mb = q_shfl(t, 0);
    // level 2: lane0 S = X·V, lane1 W = U·V, lane2 ZZ3 = ZZ·V, lane3 MM = M²
    f_mul(M2, q_sel(q == 1, U, q_sel(q == 3, Mb, A)), q_sel(q == 3, Mb, Vb));
    const F MMb = q_shfl(M2, 3), Wb = q_shfl(M2, 1);
    F X3, d;
    f_sub(X3, MMb, M2);
    f_sub(X3, X3, M2);                             // lane 0: X3 = MM − 2S
    f_sub(d, M2, X3);                              // lane 0: S − X3
    // level 3: lane0 M·(S−X3), lane1 W·Y, lane3 ZZZ3 = W·ZZZ
    f_mul(M3, q_sel(q == 0, Mb, q_sel(q == 1, M2, Wb)), q_sel(q == 0, d, A));
    t = q_shfl_xor(M3, 1);                         // lane 1 receives lane 0's product
    F Y3;
    f_sub(Y3, t, M3);
    A = q_sel(q < 2, q_sel(q == 0, X3, Y3), q_sel(q == 2, M2, M3));
    if (to_inf) q_set_inf(A);
}

// A += B.  14 products in 4 levels.
template <class F> __device__ __forceinline__ void q_add(F &A, const F &B) {
    const int q = threadIdx.x & 3;
    const bool a_zero = f_is_zero(A), b_zero = f_is_zero(B);
    const bool a_inf = q_flag(a_zero, 2), b_inf = q_flag(b_zero, 2);
    F M1, M2, M3, M4, D, t;
    // level 1: lane0 U1 = X1·ZZ2, lane1 S1 = Y1·ZZZ2, lane2 U2 = ZZ1·X2, lane3 S2 = ZZZ1·Y2
    f_mul(M1, A, q_shfl_xor(B, 2));
    t = q_shfl_xor(M1, 2);
    {
        F d0, d1;
        f_sub(d0, t, M1);
        f_sub(d1, M1, t);
        D = q_sel(q < 2, d0, d1);                  // lanes 0,2: P = U2 − U1; lanes 1,3: R = S2 − S1
    }
    const bool d_zero = f_is_zero(D);
    const bool same_x = q_flag(d_zero, 0), same_y = q_flag(d_zero, 1);
    // level 2: lane0 PP = P², lane1 RR = R², lane2 ZZ1·ZZ2, lane3 ZZZ1·ZZZ2
    f_mul(M2, q_sel(q < 2, D, A), q_sel(q < 2, D, B));
    const F PPb = q_shfl(M2, 0);
    const F Pn = q_shfl_xor(D, 1);                 // lanes 1,3 receive P; lanes 0,2 receive R
    // level 3: lane0 Q = U1·PP, lane1 PPP = P·PP, lane2 ZZ3 = ZZm·PP, lane3 ZZZm·P
    f_mul(M3, q_sel(q == 0, M1, q_sel(q == 1, Pn, M2)), q_sel(q == 3, Pn, PPb));
    f_sub(t, M2, M3);                              // lane 1: RR − PPP
    const F E = q_shfl_xor(q_sel(q == 1, t, M3), 1);  // lane 0 receives RR − PPP, lane 1 receives Q
    F X3, Qv, d;
    Qv = q_sel(q == 0, M3, E);                     // Q on lanes 0 and 1
    f_sub(X3, q_sel(q == 0, E, t), Qv);
    f_sub(X3, X3, Qv);                             // lanes 0,1: X3 = RR − PPP − 2Q
    f_sub(d, Qv, X3);                              // Q − X3
    // level 4: lane0 R·(Q−X3), lane1 S1·PPP, lane3 ZZZ3 = (ZZZm·P)·PP
    f_mul(M4, q_sel(q == 0, Pn, q_sel(q == 1, M1, M3)), q_sel(q == 0, d, q_sel(q == 1, M3, PPb)));
    t = q_shfl_xor(M4, 1);                         // lane 1 receives lane 0's product
    F Y3;
    f_sub(Y3, t, M4);
    F Rres = q_sel(q < 2, q_sel(q == 0, X3, Y3), q_sel(q == 2, M3, M4));
    if (!a_inf && !b_inf && same_x) {              // quad-uniform, rare: P + P or P − P
        if (same_y) {
            xyzz<F> full;
            q_gather(full, A);
            full = xyzz_dbl_val(full);
            Rres = q_own(full);
        } else {
            q_set_inf(Rres);
        }
    }
    Rres = q_sel(a_inf, B, Rres);                  // ∞ + B = B
    A = q_sel(b_inf, A, Rres);                     // A + ∞ = A
}

// this lane's coordinate of the Jacobian image (X·ZZ, Y·ZZZ, ZZ): lanes 0..2 hold X', Y', Z
template <class F> __device__ __forceinline__ F q_to_jac(const F &A) {
    const int q = threadIdx.x & 3;
    const bool inf = q_flag(f_is_zero(A), 2);
    F prod;
    f_mul(prod, A, q_shfl_xor(A, 2));              // lane0: X·ZZ, lane1: Y·ZZZ
    F r = q_sel(q < 2, prod, A);                   // lane2: Z = ZZ
    if (inf) f_set_zero(r);
    return r;
}
// from a Jacobian record in memory (X, Y, Z): lane0 X, lane1 Y, lane2 Z², lane3 Z³
template <class F> __device__ __forceinline__ F q_load_jac(const uint32_t *p) {
    constexpr int W = field_words<F>::value;
    const int q = threadIdx.x & 3;
    F c, z, z2, z3;
    f_load(c, p + (q < 2 ? q : 2) * W);
    z = q_shfl(c, 2);
    f_mul(z2, z, z);
    f_mul(z3, z2, z);
    F r = q_sel(q < 2, c, q_sel(q == 2, z2, z3));
    if (f_is_zero(z)) f_set_zero(r);
    return r;
}

}  // namespace b200msm
