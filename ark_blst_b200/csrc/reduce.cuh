// Bucket reduction, window combination and partial-sum kernels (template over the field).
// All of them are chains of dependent point operations, so they run on quads (quad.cuh): four
// lanes per chain, every lane holding a replica of the chain state.
#pragma once
#include "quad.cuh"

namespace b200msm {

// ---- bucket reduction -----------------------------------------------------------------------
// Weighted sum with 0-based weights, Wsum0(X) = Σ i·X[i], by segments of m:
//     Wsum0(X) = Σ_s A[s] + m·Wsum0(R),   A[s] = Wsum0(segment s),  R[s] = Sum(segment s)
// Level k consumes X_k = R_{k-1} and carries C_k[s] = Σ_{seg s} C_{k-1} + M_k·A_k[s] with
// M_k = Π_{j<k} m_j (a power of two → doublings), so that at the last level (one segment)
// C = Wsum0(X_0) and R = Sum(X_0); the window value Σ (b+1)·bucket[b] is C + R.
// Quad (w, s): window w, segment s. Arrays are window-major with the given per-window lengths.
template <class F>
__global__ void __launch_bounds__(128)
k_wsum_level(const uint32_t *__restrict__ X, const uint32_t *__restrict__ Cin, uint32_t len, uint32_t m,
             int log2M, uint32_t nwin, uint32_t *__restrict__ Rout, uint32_t *__restrict__ Cout) {
    constexpr int PW = 4 * field_words<F>::value;  // words per XYZZ point
    uint32_t nseg = len / m;
    uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const bool live = t < nseg * nwin;             // quads past the end idle along (warp stays whole)
    if (!live) t = 0;
    const bool writer = live && (threadIdx.x & 3) == 0;
    uint32_t w = t / nseg, s = t % nseg;
    const uint32_t *x = X + ((size_t)w * len + (size_t)s * m) * PW;
    xyzz<F> run, acc, tmp;
    xyzz_set_inf(run);
    xyzz_set_inf(acc);
    for (uint32_t j = m; j-- > 0;) {
        xyzz_load(tmp, x + (size_t)j * PW);
        xyzz_add_quad(run, tmp);
        if (j) xyzz_add_quad(acc, run);            // warp-uniform: element 0 has weight 0
    }
    if (writer) xyzz_store(Rout + ((size_t)w * nseg + s) * PW, run);
    for (int k = 0; k < log2M; k++) xyzz_dbl_quad(acc);
    if (Cin) {
        const uint32_t *ci = Cin + ((size_t)w * len + (size_t)s * m) * PW;
        for (uint32_t j = 0; j < m; j++) {
            xyzz_load(tmp, ci + (size_t)j * PW);
            xyzz_add_quad(acc, tmp);
        }
    }
    if (writer) xyzz_store(Cout + ((size_t)w * nseg + s) * PW, acc);
}

// window value_w = C[w] + R[w]; result = Σ_w 2^(c·w)·value_w by Horner from the top window,
// written as a Jacobian point (blst_p1 / blst_p2 layout). One quad.
template <class F>
__global__ void k_combine(const uint32_t *__restrict__ C, const uint32_t *__restrict__ R, int nwin, int c,
                          uint32_t *__restrict__ out) {
    constexpr int PW = 4 * field_words<F>::value;
    if (blockIdx.x) return;                        // one warp; quad 0 writes
    xyzz<F> acc, a;
    xyzz_set_inf(acc);
    for (int w = nwin - 1; w >= 0; w--) {
        for (int k = 0; k < c; k++) xyzz_dbl_quad(acc);
        xyzz_load(a, C + (size_t)w * PW);
        xyzz_add_quad(acc, a);
        xyzz_load(a, R + (size_t)w * PW);
        xyzz_add_quad(acc, a);
    }
    if (threadIdx.x == 0) {
        jac<F> r;
        xyzz_to_jac(r, acc);
        jac_store(out, r);
    }
}

// out = Σ partial_i (Jacobian in, Jacobian out): the final addition after the gather of per-GPU
// partial sums (or of the per-shard partials of a chunked single-GPU run). One quad.
template <class F>
__global__ void k_sum_partials(const uint32_t *__restrict__ partials, int count, uint32_t *__restrict__ out) {
    constexpr int JW = 3 * field_words<F>::value;
    if (blockIdx.x) return;                        // one warp; quad 0 writes
    xyzz<F> acc, a;
    jac<F> j;
    xyzz_set_inf(acc);
    for (int i = 0; i < count; i++) {
        jac_load(j, partials + (size_t)i * JW);
        jac_to_xyzz(a, j);
        xyzz_add_quad(acc, a);
    }
    if (threadIdx.x == 0) {
        xyzz_to_jac(j, acc);
        jac_store(out, j);
    }
}

}  // namespace b200msm
