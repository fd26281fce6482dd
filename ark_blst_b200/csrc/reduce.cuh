// Bucket reduction, window combination and partial-sum kernels (template over the field).
// All of them are chains of dependent point operations, so they run on quads (quad.cuh): four
// lanes per chain, lane q owning coordinate q (X, Y, ZZ, ZZZ) of every point in the chain.
#pragma once
#include "quad.cuh"

namespace b200msm {

// out-of-line copies of the quad operations: k_combine is a chain of dependent point operations of ≈7 µs (G1) / ≈20 µs
// (G2) each, so a call costs nothing, while the fully inlined Fp2 version took ptxas five minutes — and ran SLOWER
// (G2 combine 1.31 → 1.18 ms out of line: the inlined body does not fit the instruction cache; G1 0.60 → 0.61).
// (The same for the G2 reduction levels was measured and is slower — reduce 1.48 → 1.60 ms — they stay inlined.)
template <class F> __device__ __noinline__ void q_dbl_ni(F &A) { q_dbl(A); }
template <class F> __device__ __noinline__ void q_add_ni(F &A, const F &B) { q_add(A, B); }

// ---- bucket reduction, lower part: running sums ----------------------------------------------
// Weighted sum with 0-based weights, Wsum0(X) = Σ i·X[i], by segments of m:
//     Wsum0(X) = Σ_s A[s] + m·Wsum0(R),   A[s] = Wsum0(segment s),  R[s] = Sum(segment s)
// Level k consumes X_k = R_{k-1} and carries C_k[s] = Σ_{seg s} C_{k-1} + M_k·A_k[s] with
// M_k = Π_{j<k} m_j (a power of two → doublings), so that Wsum0(X_0) = Σ C_k + M_{k+1}·Wsum0(R_k)
// and Sum(X_0) = Sum(R_k) hold after every level.
// Quad (w, s): window w, segment s. Arrays are window-major with the given per-window lengths.
template <class F>
__global__ void __launch_bounds__(128, (field_words<F>::value == 12 ? 3 : 2))  // G1: 168 regs, no spills
k_wsum_level(const uint32_t *__restrict__ X, const uint32_t *__restrict__ Cin, uint32_t len, uint32_t m,
             int log2M, uint32_t nwin, uint32_t *__restrict__ Rout, uint32_t *__restrict__ Cout) {
    constexpr int PW = 4 * field_words<F>::value;  // words per XYZZ point
    uint32_t nseg = len / m;
    uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const bool live = t < nseg * nwin;             // quads past the end idle along (warp stays whole)
    if (!live) t = 0;
    uint32_t w = t / nseg, s = t % nseg;
    const uint32_t *x = X + ((size_t)w * len + (size_t)s * m) * PW;
    F run, acc, tmp, nxt;
    q_set_inf(run);
    q_set_inf(acc);
    q_load(nxt, x + (size_t)(m - 1) * PW);
    for (uint32_t j = m; j-- > 0;) {
        tmp = nxt;
        if (j) q_load(nxt, x + (size_t)(j - 1) * PW);  // prefetch: the global load leaves the chain
        q_add(run, tmp);
        if (j) q_add(acc, run);                    // warp-uniform: element 0 has weight 0
    }
    if (live) q_store(Rout + ((size_t)w * nseg + s) * PW, run);
    for (int k = 0; k < log2M; k++) q_dbl(acc);
    if (Cin) {
        const uint32_t *ci = Cin + ((size_t)w * len + (size_t)s * m) * PW;
        q_load(nxt, ci);
        for (uint32_t j = 0; j < m; j++) {
            tmp = nxt;
            if (j + 1 < m) q_load(nxt, ci + (size_t)(j + 1) * PW);
            q_add(acc, tmp);
        }
    }
    if (live) q_store(Cout + ((size_t)w * nseg + s) * PW, acc);
}

// The same level with ONE THREAD per (window, segment) instead of a quad: with ≥ 10^5 chains in flight the level is
// throughput-bound, and the thread form (inlined XYZZ additions, the accumulation kernel's shape) runs closer to
// the pipe than the quad form, whose 14 products per addition occupy 16 lane-slots and pay 6 shuffles of 12 words.
// Used by run_pass from 131 072 chains up (G1 2^24: 213 k chains, 9.8 → see profiles/r02_experiments.md).
template <class F>
__global__ void __launch_bounds__(128, 2)
k_wsum_level_thread(const uint32_t *__restrict__ X, const uint32_t *__restrict__ Cin, uint32_t len, uint32_t m,
                    int log2M, uint32_t nwin, uint32_t *__restrict__ Rout, uint32_t *__restrict__ Cout) {
    constexpr int PW = 4 * field_words<F>::value;
    const uint32_t nseg = len / m;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nseg * nwin) return;
    const uint32_t w = t / nseg, s = t % nseg;
    const uint32_t *x = X + ((size_t)w * len + (size_t)s * m) * PW;
    xyzz<F> run, acc, tmp;
    xyzz_set_inf(run);
    xyzz_set_inf(acc);
    for (uint32_t j = m; j-- > 0;) {
        xyzz_load(tmp, x + (size_t)j * PW);
        xyzz_add(run, tmp);                        // inlined: the operands stay in registers
        if (j) xyzz_add(acc, run);                 // element 0 has weight 0
    }
    xyzz_store(Rout + ((size_t)w * nseg + s) * PW, run);
    for (int k = 0; k < log2M; k++) xyzz_dbl_ni(acc);
    if (Cin) {
        const uint32_t *ci = Cin + ((size_t)w * len + (size_t)s * m) * PW;
        for (uint32_t j = 0; j < m; j++) {
            xyzz_load(tmp, ci + (size_t)j * PW);
            xyzz_add_ni(acc, tmp);
        }
    }
    xyzz_store(Cout + ((size_t)w * nseg + s) * PW, acc);
}

// ---- upper part of the reduction: a log-depth tree instead of more running-sum levels --------
// For X[0..S) per window, Wsum0(X) = Σ_k 2^k·V_k with V_k = Σ_{i: bit k of i set} X_i.  A node
// covering 2^j consecutive elements keeps its sum and its j partial V's; merging two siblings is
//     S' = S_l + S_r,   V'_k = V_l,k + V_r,k (k < j),   V'_j = S_r,   C' = C_l + C_r
// — independent additions, one kernel launch per tree level, every addition on its own quad.
// Depth log2(S) additions instead of ~3·m per running-sum level.
// Task layout at level j: per = j + 1 (+1 with a carried C array) additions per merged node.
template <class F>
__global__ void __launch_bounds__(128)
k_tree_level(const uint32_t *__restrict__ Sin, size_t sin_stride, const uint32_t *__restrict__ Vin,
             const uint32_t *__restrict__ Cin, size_t cin_stride, uint32_t *__restrict__ Sout, uint32_t *__restrict__ Vout,
             uint32_t *__restrict__ Cout, size_t out_stride, uint32_t S, int j, uint32_t nwin) {
    constexpr int PW = 4 * field_words<F>::value;
    const uint32_t n2 = S >> (j + 1);              // merged nodes per window
    const uint32_t per = (uint32_t)j + 1 + (Cin ? 1 : 0);
    const uint32_t ntask = n2 * per * nwin;
    uint32_t idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const bool live = idx < ntask;
    if (!live) idx = 0;                            // idle quads tag along so the warp stays whole
    const uint32_t w = idx / (n2 * per), rem = idx % (n2 * per);
    const uint32_t t = rem / per, r = rem % per;
    const uint32_t *pa, *pb;
    uint32_t *pd;
    if (r < (uint32_t)j) {                         // V_k of the two children (packed, stride j per node)
        const uint32_t *v = Vin + (size_t)w * out_stride * PW;
        pa = v + ((size_t)(2 * t) * j + r) * PW;
        pb = v + ((size_t)(2 * t + 1) * j + r) * PW;
        pd = Vout + ((size_t)w * out_stride + (size_t)t * (j + 1) + r) * PW;
    } else if (r == (uint32_t)j) {                 // node sums
        const uint32_t *sp = Sin + (size_t)w * sin_stride * PW;
        pa = sp + (size_t)(2 * t) * PW;
        pb = sp + (size_t)(2 * t + 1) * PW;
        pd = Sout + ((size_t)w * out_stride + t) * PW;
    } else {                                       // carried C sums
        const uint32_t *cp = Cin + (size_t)w * cin_stride * PW;
        pa = cp + (size_t)(2 * t) * PW;
        pb = cp + (size_t)(2 * t + 1) * PW;
        pd = Cout + ((size_t)w * out_stride + t) * PW;
    }
    F a, b;
    q_load(a, pa);
    q_load(b, pb);
    if (live && r == (uint32_t)j)                  // V'_j = S_r
        q_store(Vout + ((size_t)w * out_stride + (size_t)t * (j + 1) + j) * PW, b);
    q_add(a, b);
    if (live) q_store(pd, a);
}

// Per window: value_w = C_root + 2^log2M·(Σ_k 2^k V_k) + S_root (one quad per window, Horner over
// k), then result = Σ_w 2^(c·w)·value_w by Horner from the top window (warp 0), written as a
// Jacobian point (blst_p1 / blst_p2 layout). One block of 128 threads = 32 quads.
template <class F>
__global__ void __launch_bounds__(128)
k_combine(const uint32_t *__restrict__ Sroot, const uint32_t *__restrict__ V, const uint32_t *__restrict__ Croot,
          size_t stride, int logS, int log2M, int nwin, int c, int split_top, uint32_t *__restrict__ wsum,
          uint32_t *__restrict__ out) {
    constexpr int W = field_words<F>::value;
    constexpr int PW = 4 * W;
    const int qd = threadIdx.x >> 2;
    F acc, a;
    // split_top: the upper half of the unsigned top digit adds 2^(c−1)·Σ buckets of the last window.  Its c − 1 dependent
    // doublings run on WARP 3 (idle whenever nwin ≤ 24 — and split_top means c | 128, c ≥ 8, nwin ≤ 17) next to the
    // window values instead of after them on everybody's critical path (−50 µs at c = 16); the result goes to
    // wsum[nwin] and is added when the Horner chain starts.
    const bool top_aside = split_top && nwin <= 24;
    if (top_aside && threadIdx.x >= 96) {          // warp-uniform: the whole warp runs the chain, quad 24 stores
        q_load(a, Sroot + (size_t)(nwin - 1) * stride * PW);
        for (int k = 0; k < c - 1; k++) q_dbl_ni(a);
        if (qd == 24) q_store(wsum + (size_t)nwin * PW, a);
    } else
    for (int base = 0; base < nwin; base += 32) {  // block-uniform trip count
        const int w = base + qd < nwin ? base + qd : 0;  // surplus quads recompute window 0
        q_set_inf(acc);
        const uint32_t *v = V + (size_t)w * stride * PW;
        for (int k = logS - 1; k >= 0; k--) {      // T = Σ_k 2^k V_k
            q_dbl_ni(acc);
            q_load(a, v + (size_t)k * PW);
            q_add_ni(acc, a);
        }
        for (int k = 0; k < log2M; k++) q_dbl_ni(acc);
        if (Croot) {
            q_load(a, Croot + (size_t)w * stride * PW);
            q_add_ni(acc, a);
        }
        q_load(a, Sroot + (size_t)w * stride * PW);
        q_add_ni(acc, a);
        if (split_top && !top_aside) {             // upper half of the unsigned top digit: + 2^(c−1)·Σ buckets
            const bool extra = w == nwin - 1;      // (every quad runs the chain: q_dbl needs whole warps)
            for (int k = 0; k < c - 1; k++) q_dbl_ni(a);
            F zero;
            q_set_inf(zero);
            a = q_sel(extra, a, zero);
            q_add_ni(acc, a);
        }
        if (base + qd < nwin) q_store(wsum + (size_t)(base + qd) * PW, acc);
    }
    __syncthreads();                               // window sums visible to warp 0
    if (threadIdx.x >= 32) return;                 // warp 0 finishes (its 8 quads in lock-step)
    q_set_inf(acc);
    int top = nwin - 1;
    if (split_top) {                               // the two halves of the top digit share one weight
        q_load(acc, wsum + (size_t)top * PW);
        if (top_aside) {
            q_load(a, wsum + (size_t)nwin * PW);
            q_add_ni(acc, a);
        }
        top--;
    }
    for (int ww = top; ww >= 0; ww--) {
        if (ww != top)                             // nothing to double before the top window
            for (int k = 0; k < c; k++) q_dbl_ni(acc);
        q_load(a, wsum + (size_t)ww * PW);
        q_add_ni(acc, a);
    }
    F r = q_to_jac(acc);
    if (threadIdx.x < 3) f_store(out + threadIdx.x * W, r);
}

// out = Σ partial_i (Jacobian in, Jacobian out): the final addition after the gather of per-GPU
// partial sums (or of the per-pass partials of a chunked run). One warp; quad 0 writes.
template <class F>
__global__ void k_sum_partials(const uint32_t *__restrict__ partials, int count, uint32_t *__restrict__ out) {
    constexpr int W = field_words<F>::value;
    if (blockIdx.x) return;
    F acc;
    q_set_inf(acc);
    for (int i = 0; i < count; i++) {
        F a = q_load_jac<F>(partials + (size_t)i * 3 * W);
        q_add(acc, a);
    }
    F r = q_to_jac(acc);
    if (threadIdx.x < 3) f_store(out + threadIdx.x * W, r);
}

}  // namespace b200msm
