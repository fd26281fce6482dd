// 384-bit Montgomery arithmetic over the BLS12-381 base field for sm_100a.
//
// Layout: a field element is read exactly as blst stores it — 6×u64 little-endian Montgomery
// limbs (R = 2^384; reference src/fp.rs:482-491, 532) — and viewed as 12×u32 little-endian limbs,
// which is the same byte image.  Values are kept fully reduced in [0, p) between operations so
// equality / zero tests are limb comparisons and results are bit-identical to the CPU oracle.
//
// Multiplication is an interleaved (CIOS) Montgomery product written as PTX mad.lo.cc / madc.hi.cc
// carry chains over two column-staggered accumulators ("aligned" columns 0..11 and "shifted"
// columns 1..12).  ptxas fuses each lo/hi pair into one IMAD.WIDE.U32.X with predicate carry
// in/out, so one row costs 12 fma-pipe issues instead of 24.  Work model: 588
// 32-bit IMAD per product = 12 rows × (2·24 + 1).
#pragma once
#include <cstdint>

namespace b200msm {

struct fp {
    uint32_t l[12];
};

// p, little-endian u32 (reference src/fp.rs:25-32)
__device__ __constant__ const uint32_t FP_P[12] = {
    0xffffaaab, 0xb9feffff, 0xb153ffff, 0x1eabfffe, 0xf6b0f624, 0x6730d2a0,
    0xf38512bf, 0x64774b84, 0x434bacd7, 0x4b1ba7b6, 0x397fe69a, 0x1a0111ea};
// 2^384 mod p: Montgomery one (blstrs::fp::R, reference src/fp.rs:532)
__device__ __constant__ const uint32_t FP_ONE[12] = {
    0x0002fffd, 0x76090000, 0xc40c0002, 0xebf4000b, 0x53c758ba, 0x5f489857,
    0x70525745, 0x77ce5853, 0xa256ec6d, 0x5c071a97, 0xfa80e493, 0x15f65ec3};
// 2^768 mod p
__device__ __constant__ const uint32_t FP_R2[12] = {
    0x1c341746, 0xf4df1f34, 0x09d104f1, 0x0a76e6a6, 0x4c95b6d5, 0x8de5476c,
    0x939d83c0, 0x67eb88a9, 0xb519952d, 0x9a793e85, 0x92cae3aa, 0x11988fe5};
#define FP_M0 0xfffcfffdu  // -p^-1 mod 2^32

// compile-time copy of p so fully unrolled code can take limbs as immediates
__device__ __forceinline__ constexpr uint32_t fp_p(int i) {
    constexpr uint32_t P[12] = {0xffffaaab, 0xb9feffff, 0xb153ffff, 0x1eabfffe,
                                0xf6b0f624, 0x6730d2a0, 0xf38512bf, 0x64774b84,
                                0x434bacd7, 0x4b1ba7b6, 0x397fe69a, 0x1a0111ea};
    return P[i];
}

// ---- PTX carry-chain primitives (CC lives across consecutive volatile asm statements) ----
#define PTX3(name, ins)                                                              \
    __device__ __forceinline__ uint32_t name(uint32_t a, uint32_t b, uint32_t c) {   \
        uint32_t r;                                                                  \
        asm volatile(ins " %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));     \
        return r;                                                                    \
    }
#define PTX2(name, ins)                                                  \
    __device__ __forceinline__ uint32_t name(uint32_t a, uint32_t b) {   \
        uint32_t r;                                                      \
        asm volatile(ins " %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));     \
        return r;                                                        \
    }
PTX3(ptx_mad_lo_cc, "mad.lo.cc.u32")
PTX3(ptx_madc_lo_cc, "madc.lo.cc.u32")
PTX3(ptx_mad_hi_cc, "mad.hi.cc.u32")
PTX3(ptx_madc_hi_cc, "madc.hi.cc.u32")
PTX3(ptx_madc_hi, "madc.hi.u32")
PTX2(ptx_add_cc, "add.cc.u32")
PTX2(ptx_addc_cc, "addc.cc.u32")
PTX2(ptx_addc, "addc.u32")
PTX2(ptx_sub_cc, "sub.cc.u32")
PTX2(ptx_subc_cc, "subc.cc.u32")
PTX2(ptx_subc, "subc.u32")
#undef PTX3
#undef PTX2

__device__ __forceinline__ uint32_t ptx_mul_lo(uint32_t a, uint32_t b) { return a * b; }
__device__ __forceinline__ uint32_t ptx_mul_hi(uint32_t a, uint32_t b) { return __umulhi(a, b); }

// ---------------------------------------------------------------------------------------------
// rows of the product.  `x` points at the first limb used; limbs x[0], x[2], ... x[10] are taken.
// ---------------------------------------------------------------------------------------------
// acc[j], acc[j+1] = lo, hi of x[j]·y  (j even) — no carries, columns are disjoint
__device__ __forceinline__ void row_mul(uint32_t *acc, const uint32_t *x, uint32_t y) {
#pragma unroll
    for (int j = 0; j < 12; j += 2) {
        acc[j] = ptx_mul_lo(x[j], y);
        acc[j + 1] = ptx_mul_hi(x[j], y);
    }
}
// acc += Σ_{j even} x[j]·y·2^(32j); carry-out left in CC
__device__ __forceinline__ void row_mad(uint32_t *acc, const uint32_t *x, uint32_t y) {
    acc[0] = ptx_mad_lo_cc(x[0], y, acc[0]);
    acc[1] = ptx_madc_hi_cc(x[0], y, acc[1]);
#pragma unroll
    for (int j = 2; j < 12; j += 2) {
        acc[j] = ptx_madc_lo_cc(x[j], y, acc[j]);
        acc[j + 1] = ptx_madc_hi_cc(x[j], y, acc[j + 1]);
    }
}
// same with the modulus as multiplicand (immediates); off = 0 takes p[0],p[2].. ; off = 1 p[1],p[3]..
template <int OFF>
__device__ __forceinline__ void row_mad_p(uint32_t *acc, uint32_t y) {
    acc[0] = ptx_mad_lo_cc(fp_p(OFF), y, acc[0]);
    acc[1] = ptx_madc_hi_cc(fp_p(OFF), y, acc[1]);
#pragma unroll
    for (int j = 2; j < 12; j += 2) {
        acc[j] = ptx_madc_lo_cc(fp_p(OFF + j), y, acc[j]);
        acc[j + 1] = ptx_madc_hi_cc(fp_p(OFF + j), y, acc[j + 1]);
    }
}
// acc = (acc >> 64) + Σ_{j even} x[j]·y·2^(32j), consuming the carry already in CC:
// the two-limb down-shift that turns last row's "aligned" array into this row's "shifted" one.
__device__ __forceinline__ void row_madc_rshift(uint32_t *acc, const uint32_t *x, uint32_t y) {
#pragma unroll
    for (int j = 0; j < 10; j += 2) {
        acc[j] = ptx_madc_lo_cc(x[j], y, acc[j + 2]);
        acc[j + 1] = ptx_madc_hi_cc(x[j], y, acc[j + 3]);
    }
    acc[10] = ptx_madc_lo_cc(x[10], y, 0);
    acc[11] = ptx_madc_hi(x[10], y, 0);
}

// one b-limb: al (aligned, columns 0..11) and sh (shifted, columns 1..12) both absorb a·bi and
// the Montgomery multiple of p that clears column 0.
template <bool FIRST>
__device__ __forceinline__ void mont_row(uint32_t *al, uint32_t *sh, const uint32_t *a, uint32_t bi) {
    if (FIRST) {
        row_mul(sh, a + 1, bi);
        row_mul(al, a, bi);
    } else {
        al[0] = ptx_add_cc(al[0], sh[1]);  // old column 1 lands on new column 0
        row_madc_rshift(sh, a + 1, bi);
        row_mad(al, a, bi);
        sh[11] = ptx_addc(sh[11], 0);
    }
    uint32_t m = al[0] * FP_M0;
    row_mad_p<1>(sh, m);
    row_mad_p<0>(al, m);
    sh[11] = ptx_addc(sh[11], 0);
}

// r = r - p if r >= p   (r < 2p on entry)
__device__ __forceinline__ void fp_final_sub(uint32_t *r) {
    uint32_t t[12];
    t[0] = ptx_sub_cc(r[0], fp_p(0));
#pragma unroll
    for (int i = 1; i < 12; i++) t[i] = ptx_subc_cc(r[i], fp_p(i));
    uint32_t borrow = ptx_subc(0, 0);  // 0 or 0xffffffff
#pragma unroll
    for (int i = 0; i < 12; i++) r[i] = borrow ? r[i] : t[i];
}

// r = a·b·2^-384 mod p
__device__ __forceinline__ void fp_mul(fp &r, const fp &a, const fp &b) {
    uint32_t even[12], odd[12];
    mont_row<true>(even, odd, a.l, b.l[0]);
    mont_row<false>(odd, even, a.l, b.l[1]);
#pragma unroll
    for (int i = 2; i < 12; i += 2) {
        mont_row<false>(even, odd, a.l, b.l[i]);
        mont_row<false>(odd, even, a.l, b.l[i + 1]);
    }
    // after an even number of rows the live value is even[k] + odd[k+1] at column k
    even[0] = ptx_add_cc(even[0], odd[1]);
#pragma unroll
    for (int i = 1; i < 11; i++) even[i] = ptx_addc_cc(even[i], odd[i + 1]);
    even[11] = ptx_addc(even[11], 0);
    fp_final_sub(even);
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = even[i];
}

// (A shared out-of-line copy of the product, called with operands in registers, shrinks the
// accumulation loop from ≈117 KB to ≈40 KB of SASS but measured 1–5 % SLOWER on B200: the ≈40
// register moves per call cost more than the instruction-cache pressure they remove.)
// Squaring reuses the product. A dedicated squaring (66 off-diagonal + 12 diagonal products, 222
// instead of 288 fused IMAD.WIDE) was measured on B200 and is SLOWER inside the accumulation
// kernel (6.24 ms vs 6.13 ms at n = 2^20): it trades 66 fmaheavy issues for ~190 serial
// IADD3.X of carry rippling and grows the loop body, which is already instruction-cache bound.
__device__ __forceinline__ void fp_sqr(fp &r, const fp &a) { fp_mul(r, a, a); }

__device__ __forceinline__ void fp_add(fp &r, const fp &a, const fp &b) {
    uint32_t t[12];
    t[0] = ptx_add_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < 11; i++) t[i] = ptx_addc_cc(a.l[i], b.l[i]);
    t[11] = ptx_addc(a.l[11], b.l[11]);  // p < 2^381: no carry out of 2p
    fp_final_sub(t);
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = t[i];
}
__device__ __forceinline__ void fp_sub(fp &r, const fp &a, const fp &b) {
    uint32_t t[12];
    t[0] = ptx_sub_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < 12; i++) t[i] = ptx_subc_cc(a.l[i], b.l[i]);
    uint32_t borrow = ptx_subc(0, 0);  // 0xffffffff when a < b
    t[0] = ptx_add_cc(t[0], fp_p(0) & borrow);
#pragma unroll
    for (int i = 1; i < 11; i++) t[i] = ptx_addc_cc(t[i], fp_p(i) & borrow);
    t[11] = ptx_addc(t[11], fp_p(11) & borrow);
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = t[i];
}
__device__ __forceinline__ void fp_dbl(fp &r, const fp &a) { fp_add(r, a, a); }

__device__ __forceinline__ bool fp_is_zero(const fp &a) {
    uint32_t t = a.l[0];
#pragma unroll
    for (int i = 1; i < 12; i++) t |= a.l[i];
    return t == 0;
}
__device__ __forceinline__ bool fp_eq(const fp &a, const fp &b) {
    uint32_t t = a.l[0] ^ b.l[0];
#pragma unroll
    for (int i = 1; i < 12; i++) t |= a.l[i] ^ b.l[i];
    return t == 0;
}
// r = -a mod p (0 stays 0)
__device__ __forceinline__ void fp_neg(fp &r, const fp &a) {
    uint32_t nz = fp_is_zero(a) ? 0u : 0xffffffffu;
    uint32_t t[12];
    t[0] = ptx_sub_cc(fp_p(0) & nz, a.l[0]);
#pragma unroll
    for (int i = 1; i < 11; i++) t[i] = ptx_subc_cc(fp_p(i) & nz, a.l[i]);
    t[11] = ptx_subc(fp_p(11) & nz, a.l[11]);
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = t[i];
}
// r = neg ? -a : a
__device__ __forceinline__ void fp_cneg(fp &r, const fp &a, bool neg) {
    fp t;
    fp_neg(t, a);
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = neg ? t.l[i] : a.l[i];
}
__device__ __forceinline__ void fp_set_zero(fp &r) {
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = 0;
}
__device__ __forceinline__ void fp_set_one(fp &r) {
    constexpr uint32_t ONE[12] = {0x0002fffd, 0x76090000, 0xc40c0002, 0xebf4000b,
                                  0x53c758ba, 0x5f489857, 0x70525745, 0x77ce5853,
                                  0xa256ec6d, 0x5c071a97, 0xfa80e493, 0x15f65ec3};
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = ONE[i];
}

// 48-byte vector load/store of one element (addresses are 8-byte aligned as in the reference's
// u64 arrays; 16-byte alignment holds for every array this library allocates or accepts).
__device__ __forceinline__ void fp_load(fp &r, const uint32_t *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        uint4 v = q[i];
        r.l[4 * i] = v.x; r.l[4 * i + 1] = v.y; r.l[4 * i + 2] = v.z; r.l[4 * i + 3] = v.w;
    }
}
__device__ __forceinline__ void fp_store(uint32_t *p, const fp &a) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
#pragma unroll
    for (int i = 0; i < 3; i++) q[i] = make_uint4(a.l[4 * i], a.l[4 * i + 1], a.l[4 * i + 2], a.l[4 * i + 3]);
}

// a^(p-2): used only to leave XYZZ/Jacobian form (tests, synthetic-base generation)
static __device__ __noinline__ void fp_inv(fp &r, const fp &a) {
    fp acc, base = a;
    fp_set_one(acc);
    for (int w = 0; w < 12; w++) {
        uint32_t e = FP_P[w] - (w == 0 ? 2u : 0u);
        for (int b = 0; b < 32; b++) {
            if ((e >> b) & 1) fp_mul(acc, acc, base);
            fp_sqr(base, base);
        }
    }
    r = acc;
}

}  // namespace b200msm
