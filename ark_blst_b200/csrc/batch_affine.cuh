// Batched-affine pairing rounds in front of the XYZZ bucket accumulation.
//
// An affine addition costs 1 inversion + 3 products; sharing the inversion over a batch
// (Montgomery's trick: 3 more products per element) makes it 6 products per addition against the
// 10 (G1) / 28→17 (G2) of the XYZZ mixed addition.  The grouping pads every bucket's segment of the
// sorted entry array to a multiple of 2^R (k_prep.cu; padding slots hold 0xffffffff), so for R
// rounds the pair (2s, 2s+1) of the current array always lies inside one bucket and a round is a
// FLAT map over output slots — no bucket lookups, no scans:
//     round 1: entries (point index | sign, gathered from the bases)  → affine array A1 (half as long)
//     round r: A(r−1)                                                 → A(r)
// after which k_accumulate_direct adds each bucket's 2^−R-times-shorter run of A(R) in XYZZ form.
// With R = 2, 3/4 of all additions are affine ones.
//
// One round = five launches over flat arrays (all sizes bounded on the host, the exact slot count is
// read on the device from the scan's grand total):
//     k_ba_fwd       thread t walks slots t, t+NT, …: d = x2 − x1 (1 for a slot with nothing to add),
//                    writes the running product before each slot and its total T[t]
//     k_ba_prod_fwd  the same one level up over T (K2 entries per thread) → U
//     k_ba_invert    U ← 1/U, one divsteps inversion per thread (modinv.cuh)
//     k_ba_prod_bwd  unwinds level 2: T[i] ← 1/T[i]
//     k_ba_bwd       unwinds level 1: 1/d per slot, λ = Δy/Δx, x3 = λ² − x1 − x2, y3 = λ(x1 − x3) − y1
// Every exceptional case is exact: an empty or single-point slot copies, P + P doubles
// (d = 2y, λ = 3x²/2y), P + (−P) yields "empty".  "Empty" in the intermediate arrays is an x whose
// top word is 0xffffffff (no reduced field element has it); in the inputs it is a padding entry or
// an identity base (all-zero bytes).
#pragma once
#include "ec.cuh"
#include "modinv.cuh"

namespace b200msm {

// ---- inversion of a Montgomery-form element by divsteps ----
// (aR)^-1 as a plain integer is a^-1·R^-1; one Montgomery product with R^3 gives a^-1·R.
static __device__ __noinline__ void fp_inv_sg(fp &r, const fp &a) {
    constexpr uint32_t R3[12] = {0xd94ca1e0, 0xed48ac6b, 0x03a7adf8, 0x315f831e, 0x615e29dd, 0x9a53352a,
                                 0x921e1761, 0x34c04e5e, 0x65724728, 0x2512d435, 0x91755d4d, 0x0aa63460};
    fp x, r3;
    mi_inverse_u32(x.l, a.l);
#pragma unroll
    for (int i = 0; i < 12; i++) r3.l[i] = R3[i];
    fp_mul(r, x, r3);
}
__device__ __forceinline__ void f_inv_sg(fp &r, const fp &a) { fp_inv_sg(r, a); }
// 1/(c0 + c1·u) = (c0 − c1·u)/(c0² + c1²)
__device__ __forceinline__ void f_inv_sg(fp2 &r, const fp2 &a) {
    fp n, t;
    fp_sqr(n, a.c0);
    fp_sqr(t, a.c1);
    fp_add(n, n, t);
    fp_inv_sg(n, n);
    fp_mul(r.c0, a.c0, n);
    fp_mul(t, a.c1, n);
    fp_neg(r.c1, t);
}

// ---- the "empty" marker of the intermediate arrays ----
__device__ __forceinline__ uint32_t f_top_word(const fp &a) { return a.l[11]; }
__device__ __forceinline__ uint32_t f_top_word(const fp2 &a) { return a.c0.l[11]; }
template <class F> __device__ __forceinline__ bool f_is_marked(const F &x) { return f_top_word(x) == 0xffffffffu; }
template <class F> __device__ __forceinline__ void f_set_marked(F &x) {
    constexpr int W = field_words<F>::value;
#pragma unroll
    for (int i = 0; i < W; i++) f_set_word(x, i, 0xffffffffu);
}
template <class F> __device__ __forceinline__ bool f_eq(const F &a, const F &b) {
    constexpr int W = field_words<F>::value;
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < W; i++) t |= f_word(a, i) ^ f_word(b, i);
    return t == 0;
}

// where a round reads its points from
struct BaSrc {
    const uint32_t *pts;     // first round: the bases (or the fixed-base table); later: the previous round's output
    const uint32_t *vals;    // first round: sorted entries (point index | sign ≪ 31; 0xffffffff = padding)
    const uint32_t *endo_x;  // first round, GLV: β·x table for the entries with index ≥ n_pts …
    uint32_t n_pts;
    int img_full;            // … or (four parts, G2) the full-point image tables −ψ, ψ², −ψ³ back to back (accumulate.cuh)
};

// x of input position `pos`; false when the position holds nothing.  FIRST: `v` returns the entry
// for ba_load_y.  An identity base (x = y = 0) counts as nothing.
template <class F, bool FIRST>
__device__ __forceinline__ bool ba_load_x(const BaSrc &s, size_t pos, F &x, uint32_t &v) {
    constexpr int W = field_words<F>::value;
    if (FIRST) {
        v = s.vals[pos];
        if (v == 0xffffffffu) return false;
        uint32_t idx = v & 0x7fffffffu;
        const bool endo = idx >= s.n_pts;
        if (endo) idx -= s.n_pts;
        const uint32_t *p = (endo && s.img_full ? s.endo_x : s.pts) + (size_t)idx * (2 * W);
        f_load(x, endo && !s.img_full ? s.endo_x + (size_t)idx * W : p);
        if (f_is_zero(x)) {                       // x = 0: the identity encoding iff y = 0 too (rare either way)
            F y;
            f_load(y, p + W);
            if (f_is_zero(y)) return false;
        }
        return true;
    }
    f_load(x, s.pts + pos * (2 * W));
    return !f_is_marked(x);
}
template <class F, bool FIRST>
__device__ __forceinline__ void ba_load_y(const BaSrc &s, size_t pos, uint32_t v, F &y) {
    constexpr int W = field_words<F>::value;
    if (FIRST) {
        uint32_t idx = v & 0x7fffffffu;
        const bool endo = idx >= s.n_pts;
        if (endo) idx -= s.n_pts;
        f_load(y, (endo && s.img_full ? s.endo_x : s.pts) + (size_t)idx * (2 * W) + W);
        f_cneg(y, y, v >> 31);
    } else f_load(y, s.pts + pos * (2 * W) + W);
}

// (A software prefetch of the next slot's pair — vals, then prefetch.global.L1 of both points — was measured on
// B200 and is SLOWER: G1 2^20 accumulate 5.02 → 5.10 ms, 2^24 61.3 → 78.0 ms; at 2^24 the prefetched lines are
// evicted before their use and the gather's DRAM traffic doubles.  profiles/r02_experiments.md)
enum { BA_NONE = 0, BA_FIRST = 1, BA_SECOND = 2, BA_ADD = 3, BA_DBL = 4, BA_CANCEL = 5 };

// slots of a thread: t, t + NT, t + 2·NT, …  A warp stops at the first j whose 32 slots all lie past the end.
__device__ __forceinline__ uint32_t ba_trip_count(uint32_t t, uint32_t NT, uint32_t K, uint32_t S_out) {
    const uint32_t first = t & ~31u;
    if (first >= S_out) return 0;
    const uint32_t j = (S_out - first + NT - 1) / NT;   // number of j with j·NT + first < S_out
    return j < K ? j : K;
}

// Slot range of one launch.  Unsplit (split_align_log = 0): all S = total ≫ shift output slots of the round.  Split in
// two parts (run_pass runs them as two interleaved pipelines on two streams, so that the latency-bound second
// level + inversion of one part hides under the other part's large kernels): the boundary is half the entry count
// rounded down to a multiple of 2^split_align_log entries — a multiple of 64·2^R, so it falls on a bucket-aligned,
// warp-aligned slot in every round, and part p of round r+1 reads exactly what part p of round r wrote.
__device__ __forceinline__ void ba_range(const uint32_t *total_ptr, int shift, int part, int split_align_log, uint32_t &lo,
                                         uint32_t &S_out) {
    const uint32_t total = *total_ptr, S_all = total >> shift;
    if (split_align_log == 0) { lo = 0; S_out = S_all; return; }
    const uint32_t b = ((total >> 1) & ~((1u << split_align_log) - 1)) >> shift;
    lo = part ? b : 0;
    S_out = (part ? S_all : b) - lo;
}

template <class F, bool FIRST>
__global__ void __launch_bounds__(128)
k_ba_fwd(BaSrc s, const uint32_t *__restrict__ total_ptr, int shift, int part, int split_align_log, uint32_t NT, uint32_t K,
         uint32_t *__restrict__ prefix, uint32_t *__restrict__ T) {
    constexpr int W = field_words<F>::value;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= NT) return;
    uint32_t lo, S_out;
    ba_range(total_ptr, shift, part, split_align_log, lo, S_out);
    const uint32_t trips = ba_trip_count(t, NT, K, S_out);
    F acc;
    f_set_one(acc);
    for (uint32_t j = 0; j < trips; j++) {
        const uint32_t slot = j * NT + t;
        F d;
        f_set_one(d);
        if (slot < S_out) {
            F x1, x2;
            uint32_t v1 = 0, v2 = 0;
            const size_t g = (size_t)lo + slot;   // slot of the whole round: positions 2g, 2g + 1 of the round's input
            const bool h1 = ba_load_x<F, FIRST>(s, 2 * g, x1, v1);
            const bool h2 = ba_load_x<F, FIRST>(s, 2 * g + 1, x2, v2);
            if (h1 && h2) {
                f_sub(d, x2, x1);
                if (f_is_zero(d)) {               // same x: doubling (d = 2y) or cancellation (d = 1); rare
                    F y1, y2;
                    ba_load_y<F, FIRST>(s, 2 * g, v1, y1);
                    ba_load_y<F, FIRST>(s, 2 * g + 1, v2, y2);
                    if (f_eq(y1, y2) && !f_is_zero(y1)) f_dbl(d, y1);
                    else f_set_one(d);
                }
            }
        }
        f_store(prefix + (size_t)slot * W, acc);
        f_mul(acc, acc, d);
    }
    f_store(T + (size_t)t * W, acc);
}

template <class F, bool FIRST>
__global__ void __launch_bounds__(128)
k_ba_bwd(BaSrc s, const uint32_t *__restrict__ total_ptr, int shift, int part, int split_align_log, uint32_t NT, uint32_t K,
         const uint32_t *__restrict__ prefix, const uint32_t *__restrict__ Tinv, uint32_t *__restrict__ out) {
    constexpr int W = field_words<F>::value;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= NT) return;
    uint32_t lo, S_out;
    ba_range(total_ptr, shift, part, split_align_log, lo, S_out);
    const uint32_t trips = ba_trip_count(t, NT, K, S_out);
    F inv;
    f_load(inv, Tinv + (size_t)t * W);
    for (uint32_t j = trips; j-- > 0;) {
        const uint32_t slot = j * NT + t;
        F x1, y1, x2, y2, d, num;
        int kind = BA_NONE;
        f_set_one(d);
        if (slot < S_out) {
            uint32_t v1 = 0, v2 = 0;
            const size_t g = (size_t)lo + slot;
            const bool h1 = ba_load_x<F, FIRST>(s, 2 * g, x1, v1);
            const bool h2 = ba_load_x<F, FIRST>(s, 2 * g + 1, x2, v2);
            if (h1) ba_load_y<F, FIRST>(s, 2 * g, v1, y1);
            if (h2) ba_load_y<F, FIRST>(s, 2 * g + 1, v2, y2);
            kind = h1 ? (h2 ? BA_ADD : BA_FIRST) : (h2 ? BA_SECOND : BA_NONE);
            if (kind == BA_ADD) {
                f_sub(d, x2, x1);
                f_sub(num, y2, y1);
                if (f_is_zero(d)) {               // rare
                    if (f_is_zero(num) && !f_is_zero(y1)) {
                        kind = BA_DBL;
                        f_dbl(d, y1);
                        f_sqr(num, x1);
                        F t3;
                        f_dbl(t3, num);
                        f_add(num, num, t3);      // 3·x²
                    } else {
                        kind = BA_CANCEL;
                        f_set_one(d);
                    }
                }
            }
        }
        F pre, dinv;
        f_load(pre, prefix + (size_t)slot * W);
        f_mul(dinv, inv, pre);                    // 1/d of this slot
        f_mul(inv, inv, d);                       // inverse of the product of the slots before it
        if (slot >= S_out) continue;
        uint32_t *o = out + ((size_t)lo + slot) * (2 * W);
        if (kind == BA_ADD || kind == BA_DBL) {
            F lam, x3;
            f_mul(lam, num, dinv);
            f_sqr(x3, lam);
            f_sub(x3, x3, x1);
            f_sub(x3, x3, x2);
            f_sub(x1, x1, x3);
            f_mul(x1, lam, x1);
            f_sub(x1, x1, y1);
            f_store(o, x3);
            f_store(o + W, x1);
        } else if (kind == BA_FIRST) {
            f_store(o, x1);
            f_store(o + W, y1);
        } else if (kind == BA_SECOND) {
            f_store(o, x2);
            f_store(o + W, y2);
        } else {
            F m;
            f_set_marked(m);
            f_store(o, m);
        }
    }
}

// ---- level 2 of Montgomery's trick over the per-thread totals T[0..n): U[u] = Π T[u + j·NU] ----
template <class F>
__global__ void __launch_bounds__(128)
k_ba_prod_fwd(const uint32_t *__restrict__ T, uint32_t n, uint32_t NU, uint32_t K2, uint32_t *__restrict__ prefix2,
              uint32_t *__restrict__ U) {
    constexpr int W = field_words<F>::value;
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= NU) return;
    F acc, a;
    f_set_one(acc);
    for (uint32_t j = 0; j < K2; j++) {
        const size_t i = (size_t)j * NU + u;
        if (i >= n) break;
        f_load(a, T + i * W);
        f_store(prefix2 + i * W, acc);
        f_mul(acc, acc, a);
    }
    f_store(U + (size_t)u * W, acc);
}
template <class F>
__global__ void __launch_bounds__(64)
k_ba_invert(uint32_t *__restrict__ U, uint32_t NU) {
    constexpr int W = field_words<F>::value;
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= NU) return;
    F a;
    f_load(a, U + (size_t)u * W);
    f_inv_sg(a, a);
    f_store(U + (size_t)u * W, a);
}
template <class F>
__global__ void __launch_bounds__(128)
k_ba_prod_bwd(uint32_t *__restrict__ T, uint32_t n, uint32_t NU, uint32_t K2, const uint32_t *__restrict__ prefix2,
              const uint32_t *__restrict__ Uinv) {
    constexpr int W = field_words<F>::value;
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= NU) return;
    uint32_t trips = u < n ? (uint32_t)(((size_t)n - u + NU - 1) / NU) : 0;
    if (trips > K2) trips = K2;
    F inv, a, pre;
    f_load(inv, Uinv + (size_t)u * W);
    for (uint32_t j = trips; j-- > 0;) {
        const size_t i = (size_t)j * NU + u;
        f_load(a, T + i * W);
        f_load(pre, prefix2 + i * W);
        f_mul(pre, inv, pre);
        f_mul(inv, inv, a);
        f_store(T + i * W, pre);
    }
}

// ---- XYZZ accumulation of what the rounds left: bucket b owns positions [start[b], start[b+1]) ≫ shift of A(R) ----
template <class F>
__global__ void __launch_bounds__(128)
k_accumulate_direct(const uint32_t *__restrict__ pts, const uint32_t *__restrict__ start, const uint32_t *__restrict__ order,
                    uint32_t nb, uint32_t heavy_thr, int shift, int into, uint32_t *__restrict__ buckets) {
    constexpr int W = field_words<F>::value;
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nb) return;
    uint32_t b = order[t];
    uint32_t s = start[b] >> shift, e = start[b + 1] >> shift;
    if (e - s > heavy_thr) return;  // written by k_heavy_final
    xyzz<F> acc;
    if (into) {
        if (s == e) return;
        xyzz_load(acc, buckets + (size_t)b * (4 * W));
    } else xyzz_set_inf(acc);
    for (uint32_t j = s; j < e; j++) {
        const uint32_t *p = pts + (size_t)j * (2 * W);
        F x, y;
        f_load(x, p);
        f_load(y, p + W);
        if (f_is_marked(x)) continue;
        xyzz_madd(acc, x, y);
    }
    xyzz_store(buckets + (size_t)b * (4 * W), acc);
}

}  // namespace b200msm
