// Bucket accumulation kernel (template over the field; instantiated per group in its own TU).
#pragma once
#include "ec.cuh"

namespace b200msm {

// ---- bucket accumulation: the hot kernel ---------------------------------------------------
// One thread per bucket, buckets taken in decreasing-size order so the lanes of a warp run the
// same trip count.  Accumulator in XYZZ, points read as 16-byte vectors in the reference's affine
// layout, sign applied to y on the fly.
template <class F>
__global__ void __launch_bounds__(128)
k_accumulate(const uint32_t *__restrict__ bases, const uint32_t *__restrict__ vals,
             const uint32_t *__restrict__ start, const uint32_t *__restrict__ order, uint32_t nb,
             uint32_t *__restrict__ buckets) {
    constexpr int W = field_words<F>::value;
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nb) return;
    uint32_t b = order[t];
    uint32_t s = start[b], e = start[b + 1];
    xyzz<F> acc;
    xyzz_set_inf(acc);
    for (uint32_t j = s; j < e; j++) {
        uint32_t v = vals[j];
        const uint32_t *p = bases + (size_t)(v & 0x7fffffffu) * (2 * W);
        F x, y;
        f_load(x, p);
        f_load(y, p + W);
        if (f_is_zero(x) && f_is_zero(y)) continue;  // identity base
        f_cneg(y, y, v >> 31);
        xyzz_madd(acc, x, y);
    }
    xyzz_store(buckets + (size_t)b * (4 * W), acc);
}

}  // namespace b200msm
