// Synthetic inputs generated on the device (bench / tests only — not on the reference's path).
// Streams are counter-based so any index range can be produced anywhere, and bit-identical to
// oracle/bls12381.py::synth_scalar / synth_dlog and oracle/msm_ref.c::synth_scalar:
//   scalar_i = (splitmix64(seed+4i+j))_{j<4} masked to 255 bits, minus r once if ≥ r
//   base_i   = k_i·G,  k_i = scalar stream of (seed ^ 0x5EEDBA5E5EEDBA5E), 0 → 1
// matching the reference benches' "random affine bases × random Fr" shape (benches/group.rs:18-26)
// while keeping the discrete logs known, so Σ sᵢPᵢ = (Σ sᵢkᵢ)·G can be checked at any n.
#pragma once
#include "quad.cuh"
#include "scalar.cuh"

namespace b200msm {

// standard generators, affine Montgomery limbs (values asserted against the oracle in tests)
__device__ __constant__ const uint32_t G1_GEN_AFF[24] = {
    0xfd530c16, 0x5cb38790, 0x9976fff5, 0x7817fc67, 0x143ba1c1, 0x154f95c7,
    0xf3d0e747, 0xf0ae6acd, 0x21dbf440, 0xedce6ecc, 0x9e0bfb75, 0x12017741,
    0x0ce72271, 0xbaac93d5, 0x7918fd8e, 0x8c22631a, 0x570725ce, 0xdd595f13,
    0x50405194, 0x51ac5829, 0xad0059c0, 0x0e1c8c3f, 0x5008a26a, 0x0bbc3efc};
__device__ __constant__ const uint32_t G2_GEN_AFF[48] = {
    0x02940a10, 0xf5f28fa2, 0x87b4961a, 0xb3f5fb26, 0x3e2ae580, 0xa1a893b5,
    0x1a3caee9, 0x9894999d, 0x1863366b, 0x6f67b763, 0x4350bcd7, 0x05819192,
    0x9e23f606, 0xa5a9c075, 0xbccd60c3, 0xaaa0c59d, 0xe2867806, 0x3bb17e18,
    0x8541b367, 0x1b1ab6cc, 0xf2158547, 0xc2b6ed0e, 0x7360edf3, 0x11922a09,
    0x60494c4a, 0x4c730af8, 0x5e369c5a, 0x597cfa1f, 0xaa0a635a, 0xe7e6856c,
    0x6e0d495f, 0xbbefb5e9, 0xf0ef25a2, 0x07d3a975, 0x7e80dae5, 0x0083fd8e,
    0xdf64b05d, 0xadc0fc92, 0x2b1461dc, 0x18aa270a, 0x3be4eba0, 0x86adac6a,
    0xc93da33a, 0x79495c4e, 0xa43ccaed, 0xe7175850, 0x63de1bf2, 0x0b2bc2a1};
// 2^512 mod r: multiplying by it (Montgomery) takes a canonical scalar into Montgomery form
__device__ __constant__ const uint32_t FR_R2[8] = {0xf3f29c6d, 0xc999e990, 0x87925c23, 0x2b6cedcb,
                                                   0x7254398f, 0x05d31496, 0x9f59ff11, 0x0748d9d9};

template <class F> __device__ __forceinline__ const uint32_t *gen_aff();
template <> __device__ __forceinline__ const uint32_t *gen_aff<fp>() { return G1_GEN_AFF; }
template <> __device__ __forceinline__ const uint32_t *gen_aff<fp2>() { return G2_GEN_AFF; }

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ void synth_scalar(uint32_t *s, uint64_t seed, uint64_t i) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
        uint64_t v = splitmix64(seed + 4 * i + j);
        s[2 * j] = (uint32_t)v;
        s[2 * j + 1] = (uint32_t)(v >> 32);
    }
    s[7] &= 0x7fffffffu;
    if (fr_geq_r(s)) fr_sub_r(s);
}
// s ← s·b·2^-256 mod r
__device__ __forceinline__ void fr_mont_mul(uint32_t *s, const uint32_t *b) {
    uint32_t t[10];
#pragma unroll
    for (int i = 0; i < 10; i++) t[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            c += (uint64_t)s[j] * b[i] + t[j];
            t[j] = (uint32_t)c;
            c >>= 32;
        }
        c += t[8];
        t[8] = (uint32_t)c;
        t[9] = (uint32_t)(c >> 32);
        uint32_t m = 0u - t[0];
        c = ((uint64_t)m * FR_MOD[0] + t[0]) >> 32;
#pragma unroll
        for (int j = 1; j < 8; j++) {
            c += (uint64_t)m * FR_MOD[j] + t[j];
            t[j - 1] = (uint32_t)c;
            c >>= 32;
        }
        c += t[8];
        t[7] = (uint32_t)c;
        t[8] = t[9] + (uint32_t)(c >> 32);
    }
    if (t[8] || fr_geq_r(t)) fr_sub_r(t);
#pragma unroll
    for (int i = 0; i < 8; i++) s[i] = t[i];
}

static __global__ void __launch_bounds__(256)
k_synth_scalars(uint64_t seed, size_t n, int mont, uint32_t *__restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s[8];
    synth_scalar(s, seed, i);
    if (mont) fr_mont_mul(s, FR_R2);
    uint4 *p = reinterpret_cast<uint4 *>(out + 8 * i);
    p[0] = make_uint4(s[0], s[1], s[2], s[3]);
    p[1] = make_uint4(s[4], s[5], s[6], s[7]);
}

// base_i = k_i·G by MSB-first double-and-add, normalised to affine
template <class F>
__global__ void __launch_bounds__(128)
k_synth_bases(uint64_t seed, size_t n, uint32_t *__restrict__ out) {
    constexpr int W = field_words<F>::value;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t k[8];
    synth_scalar(k, seed ^ 0x5EEDBA5E5EEDBA5EULL, i);
    uint32_t any = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) any |= k[j];
    if (!any) k[0] = 1;
    F gx, gy;
    f_load(gx, gen_aff<F>());
    f_load(gy, gen_aff<F>() + W);
    xyzz<F> acc;
    xyzz_set_inf(acc);
    for (int bit = 254; bit >= 0; bit--) {
        xyzz_dbl_ni(acc);
        if ((k[bit >> 5] >> (bit & 31)) & 1) xyzz_madd_ni(acc, gx, gy);
    }
    affine<F> a;
    xyzz_to_affine(a, acc);
    f_store(out + i * 2 * W, a.x);
    f_store(out + i * 2 * W + W, a.y);
}

// ---- integer-pipe peak (roofline denominator), same kernels as tools/imad_peak.cu ----
template <int MODE>
__global__ void __launch_bounds__(256) k_imad_peak(uint32_t *out, uint32_t seed, int iters) {
    uint32_t a = seed + threadIdx.x, b = seed * 3 + 1 + threadIdx.x * 2;
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 16; i++) r[i] = seed + i + threadIdx.x * 7;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 64; k++) {
            if (MODE == 0) {
#pragma unroll
                for (int c = 0; c < 8; c++)
                    asm volatile("mad.lo.u32 %0, %0, %2, %1;" : "+r"(r[c]) : "r"(a), "r"(b));
            } else {
                asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(r[0]) : "r"(a), "r"(r[15]));
                asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(r[1]) : "r"(a), "r"(b));
#pragma unroll
                for (int c = 1; c < 8; c++) {
                    asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(r[2 * c]) : "r"(a), "r"(b));
                    asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(r[2 * c + 1]) : "r"(a), "r"(b));
                }
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s ^= r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- unit hooks (parity tests) ---------------------------------------------------------------
template <class F>
__global__ void k_dbg_field_op(int op, const uint32_t *a, const uint32_t *b, uint32_t *out, size_t n) {
    constexpr int W = field_words<F>::value;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F x, y, r;
    f_load(x, a + i * W);
    if (op <= 2) f_load(y, b + i * W);
    switch (op) {
        case 0: f_mul(r, x, y); break;
        case 1: f_add(r, x, y); break;
        case 2: f_sub(r, x, y); break;
        case 3: f_sqr(r, x); break;
        case 4: f_cneg(r, x, true); break;
        default: f_inv(r, x); break;
    }
    f_store(out + i * W, r);
}
template <class F>
__global__ void k_dbg_point_op(int op, const uint32_t *acc_in, const uint32_t *q, uint32_t *out, size_t n) {
    constexpr int W = field_words<F>::value;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    xyzz<F> acc;
    xyzz_load(acc, acc_in + i * 4 * W);
    if (op == 0) {
        F x, y;
        f_load(x, q + i * 2 * W);
        f_load(y, q + i * 2 * W + W);
        if (!(f_is_zero(x) && f_is_zero(y))) xyzz_madd_ni(acc, x, y);
    } else if (op == 1) {
        xyzz<F> b;
        xyzz_load(b, q + i * 4 * W);
        xyzz_add_ni(acc, b);
    } else {
        xyzz_dbl_ni(acc);
    }
    jac<F> r;
    xyzz_to_jac(r, acc);
    jac_store(out + i * 3 * W, r);
}


// quad-distributed variants of the same hooks: op 3 = q_add (XYZZ + XYZZ), op 4 = q_dbl
template <class F>
__global__ void __launch_bounds__(128)
k_dbg_point_op_quad(int op, const uint32_t *acc_in, const uint32_t *qin, uint32_t *out, size_t n) {
    constexpr int W = field_words<F>::value;
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const bool live = i < n;
    if (!live) i = 0;
    F a, b;
    q_load(a, acc_in + i * 4 * W);
    if (op == 3) {
        q_load(b, qin + i * 4 * W);
        q_add(a, b);
    } else {
        q_dbl(a);
    }
    F r = q_to_jac(a);
    if (live && (threadIdx.x & 3) < 3) f_store(out + i * 3 * W + (threadIdx.x & 3) * W, r);
}

}  // namespace b200msm
