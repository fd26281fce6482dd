// Bucket accumulation (template over the field; instantiated per group in its own TU).
//
// Light buckets (≤ `heavy_thr` entries — everything a uniform scalar distribution produces) take
// one thread each, in decreasing-size order so the lanes of a warp run the same trip count.
// Heavy buckets — the degenerate top window when c does not divide 255 evenly, or the digit-1
// bucket of witness-like scalars full of ones — are cut into block-sized tasks whose partial sums
// are tree-reduced in shared memory and then folded per bucket; task lists and counts live in
// device memory, so no host round trip is needed to size that work.
#pragma once
#include "ec.cuh"

namespace b200msm {

struct HeavyHeader {
    uint32_t n_heavy;  // heavy buckets
    uint32_t n_tasks;  // block tasks over all heavy buckets
};
struct HeavyBucket {
    uint32_t bucket, task_base, n_tasks;
};
struct HeavyTask {
    uint32_t offset, len;  // slice of the sorted value array
};

// ---- light buckets: the hot kernel -----------------------------------------------------------
// Accumulator in XYZZ, points read as 16-byte vectors in the reference's affine layout, sign
// applied to y on the fly.
template <class F>
__global__ void __launch_bounds__(128)
k_accumulate(const uint32_t *__restrict__ bases, const uint32_t *__restrict__ vals,
             const uint32_t *__restrict__ start, const uint32_t *__restrict__ order, uint32_t nb,
             uint32_t heavy_thr, const uint32_t *__restrict__ endo_x, uint32_t n_pts, int img_full, int into,
             uint32_t *__restrict__ buckets) {
    constexpr int W = field_words<F>::value;
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nb) return;
    uint32_t b = order[t];
    uint32_t s = start[b], e = start[b + 1];
    if (e - s > heavy_thr) return;  // written by k_heavy_final
    xyzz<F> acc;
    // `into`: a later slice of a streamed MSM (the bases arrive from the host in slices, every slice
    // is accumulated into the same buckets, one reduction at the end) continues from the bucket
    if (into) {
        if (s == e) return;
        xyzz_load(acc, buckets + (size_t)b * (4 * W));
    } else xyzz_set_inf(acc);
    uint32_t v = s < e ? vals[s] : 0;
    for (uint32_t j = s; j < e; j++) {
        // index ≥ n_pts (GLV): an endomorphism image of point index − n_pts.  Two parts: φ(P) = (β·x, y), x from
        // the precomputed β·x table, y from the base.  Four parts (img_full; G2): the full point from the image
        // tables −ψ(Q), ψ²(Q), −ψ³(Q) stored back to back (index − n_pts runs over 3·n_pts entries).
        uint32_t idx = v & 0x7fffffffu;
        const bool endo = idx >= n_pts;
        if (endo) idx -= n_pts;
        const uint32_t *p = (endo && img_full ? endo_x : bases) + (size_t)idx * (2 * W);
        const uint32_t *px = endo && !img_full ? endo_x + (size_t)idx * W : p;
        const uint32_t sign = v >> 31;
        F x, y;
        f_load(x, px);
        f_load(y, p + W);
        if (j + 1 < e) {  // pull the next point towards L1 while this one is being added
            v = vals[j + 1];
            uint32_t nidx = v & 0x7fffffffu;
            const bool nendo = nidx >= n_pts;
            if (nendo) nidx -= n_pts;
            const char *q = reinterpret_cast<const char *>((nendo && img_full ? endo_x : bases) + (size_t)nidx * (2 * W));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(q));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(q + 8 * W - 4));
        }
        if (f_is_zero(x) && f_is_zero(y)) continue;  // identity base
        f_cneg(y, y, sign);
        xyzz_madd(acc, x, y);
    }
    xyzz_store(buckets + (size_t)b * (4 * W), acc);
}

// ---- fixed-base window table (resident bases) -------------------------------------
// With table[w][i] = 2^(c·w)·P_i stored in affine form, Σ_w 2^(cw)·Σ_i d_{w,i}·P_i becomes
// Σ_{w,i} d_{w,i}·table[w][i]: ONE bucket set shared by all windows, so the bucket reduction runs
// over a single window and the 255 dependent doublings of the Horner chain disappear.  The
// grouping kernels then group by the digit alone and emit the table index w·stride + i, and the
// kernels above run unchanged with `bases` = the table (k_prep.cu: k_hist / k_scatter, tbl_stride).
// table[w] from table[w−1]: c doublings of every point, written as Jacobian for the batch
// normalisation that follows (one launch pair per window; runs once per uploaded key)
template <class F>
__global__ void __launch_bounds__(128)
k_table_shift(const uint32_t *__restrict__ prev, size_t n, int c, uint32_t *__restrict__ jac_out) {
    constexpr int W = field_words<F>::value;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F x, y;
    f_load(x, prev + i * (2 * W));
    f_load(y, prev + i * (2 * W) + W);
    xyzz<F> p;
    if (f_is_zero(x) && f_is_zero(y)) xyzz_set_inf(p);
    else {
        xyzz_mdbl(p, x, y);
        for (int k = 1; k < c; k++) xyzz_dbl_ni(p);
    }
    jac<F> r;
    xyzz_to_jac(r, p);
    jac_store(jac_out + i * (3 * W), r);
}

// β·x for every base (GLV): the x coordinates of the endomorphism images φ(P) = (β·x, y) = λ·P, written as a
// dense table of field elements. β is the cube root of unity in Fp that makes φ act as THE λ of
// scalar.cuh's decomposition: β on G1 and β² on G2 (checked against the big-int oracle: λ·Q = (β²·x, y)
// for Q in G2). Fp2 coordinates take β on both components. Montgomery limbs.
__device__ __constant__ const uint32_t GLV_BETA[12] = {0x8671f071, 0xcd03c9e4, 0x1fcda5d2, 0x5dab2246, 0xd3851b95, 0x587042af,
                                                       0x01bacb9e, 0x8eb60ebe, 0x83d050d2, 0x03f97d6e, 0x54638741, 0x18f02065};
__device__ __constant__ const uint32_t GLV_BETA_SQ[12] = {0x798a64e8, 0x30f1361b, 0x7ece5a2a, 0xf3b8ddab, 0xc61577f7, 0x16a8ca3a,
                                                          0x74fd029b, 0xc26a2ff8, 0x60701c6e, 0x3636b766, 0x241b6160, 0x051ba4ab};
__device__ __forceinline__ void endo_mul(fp &r, const fp &x) {
    fp beta;
    fp_load(beta, GLV_BETA);
    fp_mul(r, x, beta);
}
__device__ __forceinline__ void endo_mul(fp2 &r, const fp2 &x) {
    fp beta;
    fp_load(beta, GLV_BETA_SQ);
    fp_mul(r.c0, x.c0, beta);
    fp_mul(r.c1, x.c1, beta);
}
// Four-part decomposition on G2 (gls4.cuh): the three image tables −ψ(Q), ψ²(Q), −ψ³(Q) of every base, full points,
// stored back to back (n points each).  ψ(x, y) = (x̄·γx, ȳ·γy) with γx = (1+u)^−(p−1)/3 = (0, γx1) and
// γy = (1+u)^−(p−1)/2; ψ²(x, y) = (β·x, −y).  The identity (all-zero) maps to itself.  Montgomery limbs; checked
// against the big-int oracle (tests/test_glv_constants.py: constants, ψ(Q) = [z]·Q).
__device__ __constant__ const uint32_t PSI_GX_C1[12] = {0x867545c3, 0x890dc9e4, 0x3285a5d5, 0x2af32253, 0x309b7e2c, 0x50880866,
                                                        0x7e881024, 0xa20d1b8c, 0xe2db9068, 0x14e4f04f, 0x1564853a, 0x14e56d3f};
__device__ __constant__ const uint32_t PSI_GY_C0[12] = {0xa55c9ad1, 0x3e2f585d, 0x86c18183, 0x4294213d, 0x8b623732, 0x382844c8,
                                                        0x19103e18, 0x92ad2afd, 0xac7cf0b9, 0x1d794e4f, 0x7d825ec8, 0x0bd592fc};
__device__ __constant__ const uint32_t PSI_GY_C1[12] = {0x5aa30fda, 0x7bcfa7a2, 0x2a927e7c, 0xdc17dec1, 0x6b4ebef1, 0x2f088dd8,
                                                        0xda74d4a7, 0xd1ca2087, 0x96cebc1d, 0x2da25966, 0xbbfd87d2, 0x0e2b7eed};
// (x̄·γx, ȳ·γy) of an Fp2 point, y negated when `neg_y`
__device__ __forceinline__ void psi_apply(fp2 &rx, fp2 &ry, const fp2 &x, const fp2 &y, bool neg_y) {
    fp g, t;
    fp_load(g, PSI_GX_C1);
    // (x0 − x1·u)·(γ·u) = x1·γ + x0·γ·u
    fp_mul(t, x.c1, g);
    fp_mul(rx.c1, x.c0, g);
    rx.c0 = t;
    fp2 gy, yc;
    fp_load(gy.c0, PSI_GY_C0);
    fp_load(gy.c1, PSI_GY_C1);
    yc.c0 = y.c0;
    fp_neg(yc.c1, y.c1);
    fp2_mul(ry, yc, gy);
    if (neg_y) { fp_neg(ry.c0, ry.c0); fp_neg(ry.c1, ry.c1); }
}
static __global__ void __launch_bounds__(128)
k_psi_tables(const uint32_t *__restrict__ bases, size_t n, uint32_t *__restrict__ img) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fp2 x, y, x1, y1, x2, y2, x3, y3;
    f_load(x, bases + i * 48);
    f_load(y, bases + i * 48 + 24);
    psi_apply(x1, y1, x, y, true);                 // −ψ(Q)
    fp beta;
    fp_load(beta, GLV_BETA);
    fp_mul(x2.c0, x.c0, beta);                     // ψ²(Q) = (β·x, −y)
    fp_mul(x2.c1, x.c1, beta);
    fp_neg(y2.c0, y.c0);
    fp_neg(y2.c1, y.c1);
    psi_apply(x3, y3, x2, y2, true);               // −ψ³(Q) = −ψ(ψ²(Q))
    f_store(img + i * 48, x1);
    f_store(img + i * 48 + 24, y1);
    f_store(img + (n + i) * 48, x2);
    f_store(img + (n + i) * 48 + 24, y2);
    f_store(img + (2 * n + i) * 48, x3);
    f_store(img + (2 * n + i) * 48 + 24, y3);
}

template <class F>
static __global__ void __launch_bounds__(256)
k_endo_table(const uint32_t *__restrict__ bases, size_t n, uint32_t *__restrict__ endo_x) {
    constexpr int W = field_words<F>::value;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F x;
    f_load(x, bases + i * (2 * W));
    endo_mul(x, x);
    f_store(endo_x + i * W, x);
}

// ---- heavy buckets ---------------------------------------------------------------------------
// one thread per bucket rank: heavy ones claim a slot and a run of tasks of ≤ chunk entries
static __global__ void __launch_bounds__(256)
k_plan_heavy(const uint32_t *__restrict__ start, const uint32_t *__restrict__ order, uint32_t nb, uint32_t heavy_thr,
             uint32_t chunk, int shift, HeavyHeader *__restrict__ hdr, HeavyBucket *__restrict__ hb, HeavyTask *__restrict__ tasks) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nb) return;
    uint32_t b = order[t];
    // shift > 0: after batched-affine rounds (batch_affine.cuh) the bucket owns positions [start[b], start[b+1]) >> shift
    uint32_t s = start[b] >> shift, cnt = (start[b + 1] >> shift) - s;
    if (cnt <= heavy_thr) return;
    uint32_t nt = (cnt + chunk - 1) / chunk;
    uint32_t slot = atomicAdd(&hdr->n_heavy, 1u);
    uint32_t base = atomicAdd(&hdr->n_tasks, nt);
    hb[slot] = HeavyBucket{b, base, nt};
    for (uint32_t k = 0; k < nt; k++) {
        uint32_t off = k * chunk;
        tasks[base + k] = HeavyTask{s + off, cnt - off < chunk ? cnt - off : chunk};
    }
}

// block tree-sum of one XYZZ value per thread through shared memory; result in thread 0's `acc`
template <class F, int THREADS>
__device__ __forceinline__ void block_tree_sum(xyzz<F> &acc, uint32_t *smem) {
    constexpr int PW = 4 * field_words<F>::value;
    for (int stride = THREADS / 2; stride >= 1; stride >>= 1) {
        if ((int)threadIdx.x >= stride && (int)threadIdx.x < 2 * stride) xyzz_store(smem + (size_t)(threadIdx.x - stride) * PW, acc);
        __syncthreads();
        if ((int)threadIdx.x < stride) {
            xyzz<F> o;
            xyzz_load(o, smem + (size_t)threadIdx.x * PW);
            if (field_words<F>::value == 12) xyzz_add(acc, o);   // G1: inlined, both operands stay in registers (witness-like 2^20: 5.34 → 4.81 ms)
            else xyzz_add_ni(acc, o);                            // G2: out of line (the inlined Fp2 tree took ptxas five minutes for ≈1 % of a rare path)
        }
        __syncthreads();
    }
}

// grid-stride over tasks: each block sums ≤ chunk points into one partial
template <class F, int THREADS>
__global__ void __launch_bounds__(THREADS)
k_heavy_tasks(const uint32_t *__restrict__ bases, const uint32_t *__restrict__ vals,
              const HeavyHeader *__restrict__ hdr, const HeavyTask *__restrict__ tasks, const uint32_t *__restrict__ endo_x,
              uint32_t n_pts, int img_full, uint32_t *__restrict__ partials) {
    constexpr int W = field_words<F>::value;
    constexpr int PW = 4 * W;
    __shared__ __align__(16) uint32_t smem[(THREADS / 2) * PW];
    const uint32_t nt = hdr->n_tasks;
    for (uint32_t t = blockIdx.x; t < nt; t += gridDim.x) {
        HeavyTask tk = tasks[t];
        xyzz<F> acc;
        xyzz_set_inf(acc);
        for (uint32_t j = threadIdx.x; j < tk.len; j += THREADS) {
            if (!vals) {                               // direct: the points themselves, as the batched-affine rounds left them
                const uint32_t *p = bases + (size_t)(tk.offset + j) * (2 * W);
                F x, y;
                f_load(x, p);
                if (f_word(x, 11) == 0xffffffffu) continue;   // the rounds' "empty" marker
                f_load(y, p + W);
                xyzz_madd(acc, x, y);
                continue;
            }
            uint32_t v = vals[tk.offset + j];
            uint32_t idx = v & 0x7fffffffu;
            const bool endo = idx >= n_pts;
            if (endo) idx -= n_pts;
            const uint32_t *p = (endo && img_full ? endo_x : bases) + (size_t)idx * (2 * W);
            F x, y;
            f_load(x, endo && !img_full ? endo_x + (size_t)idx * W : p);
            f_load(y, p + W);
            if (f_is_zero(x) && f_is_zero(y)) continue;
            f_cneg(y, y, v >> 31);
            xyzz_madd(acc, x, y);                      // inlined: the accumulator stays in registers
        }
        block_tree_sum<F, THREADS>(acc, smem);
        if (threadIdx.x == 0) xyzz_store(partials + (size_t)t * PW, acc);
    }
}

// grid-stride over heavy buckets: each block folds the bucket's task partials into the bucket
template <class F, int THREADS>
__global__ void __launch_bounds__(THREADS)
k_heavy_final(const HeavyHeader *__restrict__ hdr, const HeavyBucket *__restrict__ hb,
              const uint32_t *__restrict__ partials, int into, uint32_t *__restrict__ buckets) {
    constexpr int PW = 4 * field_words<F>::value;
    __shared__ __align__(16) uint32_t smem[(THREADS / 2) * PW];
    const uint32_t nh = hdr->n_heavy;
    for (uint32_t h = blockIdx.x; h < nh; h += gridDim.x) {
        HeavyBucket B = hb[h];
        xyzz<F> acc, o;
        xyzz_set_inf(acc);
        for (uint32_t k = threadIdx.x; k < B.n_tasks; k += THREADS) {
            xyzz_load(o, partials + (size_t)(B.task_base + k) * PW);
            xyzz_add_ni(acc, o);
        }
        if (B.n_tasks > 1) block_tree_sum<F, THREADS>(acc, smem);  // block-uniform condition
        if (threadIdx.x == 0) {
            if (into) {                            // streamed MSM: earlier slices already fed this bucket
                xyzz_load(o, buckets + (size_t)B.bucket * PW);
                xyzz_add_ni(acc, o);
            }
            xyzz_store(buckets + (size_t)B.bucket * PW, acc);
        }
    }
}

}  // namespace b200msm
