// Batch normalisation of Jacobian points to affine on the device.
// Reference: `CurveGroup::normalize_batch` → `blstrs::G?Projective::batch_normalize`
// (src/g1.rs:536-543, src/g2.rs:516-523) — the step every caller runs right before `msm`
// (src/tests.rs:63, ScalarMul::batch_convert_to_mul_base src/g1.rs:597-599).
//
// Montgomery's trick per thread over K points taken with a grid-sized stride (so neighbouring
// threads touch neighbouring 144/288-byte records): forward pass stores the running product of the
// Z's in the x slot of the output record, one Fermat inversion per thread, backward pass peels
// one Z at a time.  (x, y) = (X/Z², Y/Z³); identity (Z = 0) → all-zero affine, as blst encodes it.
// Cost per point: 7 field products + one inversion (≈570 products) per K points.
#pragma once
#include "ec.cuh"

namespace b200msm {

template <class F>
__global__ void __launch_bounds__(128)
k_normalize_batch(const uint32_t *__restrict__ proj, size_t n, uint32_t *__restrict__ aff) {
    constexpr int W = field_words<F>::value;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    F acc, z;
    f_set_one(acc);
    for (size_t i = t; i < n; i += stride) {            // forward: prefix products of the non-zero Z's
        f_store(aff + i * 2 * W, acc);
        f_load(z, proj + i * 3 * W + 2 * W);
        if (!f_is_zero(z)) f_mul(acc, acc, z);
    }
    F inv;
    f_inv(inv, acc);
    size_t last = t + ((n - 1 - t) / stride) * stride;  // largest index of this thread
    for (size_t i = last;; i -= stride) {               // backward
        f_load(z, proj + i * 3 * W + 2 * W);
        F x, y;
        if (f_is_zero(z)) {
            f_set_zero(x);
            f_set_zero(y);
        } else {
            F pre, zi, zi2;
            f_load(pre, aff + i * 2 * W);
            f_mul(zi, inv, pre);                         // 1/Z_i
            f_mul(inv, inv, z);
            f_sqr(zi2, zi);
            f_load(x, proj + i * 3 * W);
            f_load(y, proj + i * 3 * W + W);
            f_mul(x, x, zi2);
            f_mul(zi2, zi2, zi);
            f_mul(y, y, zi2);
        }
        f_store(aff + i * 2 * W, x);
        f_store(aff + i * 2 * W + W, y);
        if (i == t) break;
    }
}

}  // namespace b200msm
