// C++ mirror of the reference's arkworks surface for the MSM path (the reference's host language,
// Rust, has no toolchain in this image; rust/gpu.rs holds the same shim as Rust text).
//
// Mirrors, name for name:
//   impl VariableBaseMSM for G1Projective { fn msm(bases, scalars) -> Result<Self, usize> }
//                                                         reference src/g1.rs:602-632
//   impl VariableBaseMSM for G2Projective                  reference src/g2.rs:582-612
//   msm_bigint / msm_unchecked                             arkworks defaults the new build overrides
//   ScalarMul::{MulBase, NEGATION_IS_CHEAP}                reference src/g1.rs:593-600
// Types are the reference's #[repr(transparent)] layouts (src/g1.rs:54-56,435-437,
// src/g2.rs:66-68,415-417, src/scalar.rs:23-25): plain limb arrays, passed by pointer, no copies.
// Header-only; link with -lb200msm.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/b200msm.h"

namespace ark_blst {

using usize = std::size_t;

struct Scalar { uint64_t l[4]; };        // Montgomery Fr (blstrs::Scalar)
struct BigInt4 { uint64_t l[4]; };       // ark_ff::BigInt<4>, canonical little-endian
struct G1Affine { uint64_t l[12]; };     // blst_p1_affine
struct G2Affine { uint64_t l[24]; };     // blst_p2_affine
static_assert(sizeof(Scalar) == 32 && sizeof(BigInt4) == 32, "scalar layout");
static_assert(sizeof(G1Affine) == 96 && sizeof(G2Affine) == 192, "affine layout");

// Result<T, usize>
template <class T> struct Result {
    bool ok;
    T value;
    usize err;
    bool is_ok() const { return ok; }
    bool is_err() const { return !ok; }
    const T &unwrap() const {
        if (!ok) throw std::string("called unwrap() on Err(") + std::to_string(err) + ")";
        return value;
    }
    usize unwrap_err() const { return err; }
    static Result Ok(const T &v) { return Result{true, v, 0}; }
    static Result Err(usize e) { return Result{false, T{}, e}; }
};

// Engine binding — the counterpart of the reference's `Device::all()` + `devices[0]`
// (src/gpu.rs:233-234): which GPUs a sharded MSM runs on.  Mirrors rust/gpu.rs
// init_devices / shutdown_devices / device_count.  Optional: without it the first msm binds the
// current device only.  Binding another range while bound is Err(0) (shutdown_devices first).
inline Result<usize> init_devices(usize first_device, usize n_devices /* 0 = all visible */);
inline void shutdown_devices() { b200msm_shutdown(); }
inline usize device_count() { return static_cast<usize>(b200msm_device_count()); }
inline std::string last_error() { return b200msm_last_error(); }

namespace detail {
template <class Proj, class Aff>
Result<Proj> call(int (*f)(const uint64_t *, const uint64_t *, size_t, int, uint64_t *), const Aff *bases, usize nb,
                  const uint64_t *scalars, usize ns, int mont) {
    if (nb != ns) return Result<Proj>::Err(nb < ns ? nb : ns);   // arkworks: Err(min(len))
    Proj out{};
    int rc = f(reinterpret_cast<const uint64_t *>(bases), scalars, nb, mont, out.l);
    if (rc != 0) return Result<Proj>::Err(0);                      // reference GPU arm: Err(0), src/g1.rs:628-630
    return Result<Proj>::Ok(out);
}
}  // namespace detail

inline Result<usize> init_devices(usize first_device, usize n_devices) {
    if (b200msm_init(static_cast<int>(first_device), static_cast<int>(n_devices)) != 0) return Result<usize>::Err(0);
    return Result<usize>::Ok(device_count());
}

struct G1Projective {                    // blst_p1 (Jacobian)
    uint64_t l[18];
    using MulBase = G1Affine;
    static constexpr bool NEGATION_IS_CHEAP = true;
    bool is_zero() const { for (int i = 12; i < 18; i++) if (l[i]) return false; return true; }

    static Result<G1Projective> msm(const G1Affine *bases, usize nb, const Scalar *scalars, usize ns) {
        return detail::call<G1Projective>(b200msm_g1, bases, nb, reinterpret_cast<const uint64_t *>(scalars), ns, 1);
    }
    static G1Projective msm_bigint(const G1Affine *bases, usize nb, const BigInt4 *bigints, usize ns) {
        usize n = nb < ns ? nb : ns;     // infallible and truncating, like arkworks' msm_bigint
        return detail::call<G1Projective>(b200msm_g1, bases, n, reinterpret_cast<const uint64_t *>(bigints), n, 0).unwrap();
    }
    // CurveGroup::normalize_batch (reference src/g1.rs:536-543); throws on a device error
    static std::vector<G1Affine> normalize_batch(const G1Projective *v, usize n) {
        std::vector<G1Affine> out(n);
        if (b200msm_normalize_batch(B200MSM_G1, reinterpret_cast<const uint64_t *>(v), n, n ? out[0].l : nullptr) != 0)
            throw std::string("normalize_batch: ") + b200msm_last_error();
        return out;
    }
    static G1Projective msm_unchecked(const G1Affine *bases, usize nb, const Scalar *scalars, usize ns) {
        usize n = nb < ns ? nb : ns;
        return msm(bases, n, scalars, n).unwrap();
    }
};

struct G2Projective {                    // blst_p2 (Jacobian)
    uint64_t l[36];
    using MulBase = G2Affine;
    static constexpr bool NEGATION_IS_CHEAP = true;
    bool is_zero() const { for (int i = 24; i < 36; i++) if (l[i]) return false; return true; }

    static Result<G2Projective> msm(const G2Affine *bases, usize nb, const Scalar *scalars, usize ns) {
        return detail::call<G2Projective>(b200msm_g2, bases, nb, reinterpret_cast<const uint64_t *>(scalars), ns, 1);
    }
    static G2Projective msm_bigint(const G2Affine *bases, usize nb, const BigInt4 *bigints, usize ns) {
        usize n = nb < ns ? nb : ns;
        return detail::call<G2Projective>(b200msm_g2, bases, n, reinterpret_cast<const uint64_t *>(bigints), n, 0).unwrap();
    }
    static std::vector<G2Affine> normalize_batch(const G2Projective *v, usize n) {
        std::vector<G2Affine> out(n);
        if (b200msm_normalize_batch(B200MSM_G2, reinterpret_cast<const uint64_t *>(v), n, n ? out[0].l : nullptr) != 0)
            throw std::string("normalize_batch: ") + b200msm_last_error();
        return out;
    }
    static G2Projective msm_unchecked(const G2Affine *bases, usize nb, const Scalar *scalars, usize ns) {
        usize n = nb < ns ? nb : ns;
        return msm(bases, n, scalars, n).unwrap();
    }
};
static_assert(sizeof(G1Projective) == 144 && sizeof(G2Projective) == 288, "projective layout");

// Resident bases: a proving key's bases uploaded once, many scalar vectors run
// against them. The reference has no counterpart — its GPU arm re-uploads the bases and rebuilds
// the program on every call (src/gpu.rs:149-150,233-237). `precompute()` turns the upload into a
// fixed-base window table (b200msm_bases_precompute): one bucket set, no Horner chain.
template <class Proj> class ResidentBases {
    b200msm_bases *h_ = nullptr;
    usize n_ = 0;
    static constexpr int group() { return sizeof(Proj) == 144 ? B200MSM_G1 : B200MSM_G2; }

  public:
    using Affine = typename Proj::MulBase;
    ResidentBases(const Affine *bases, usize n) : n_(n) {
        if (b200msm_bases_upload(group(), reinterpret_cast<const uint64_t *>(bases), n, &h_) != 0)
            throw std::string("bases_upload: ") + b200msm_last_error();
    }
    ResidentBases(const ResidentBases &) = delete;
    ResidentBases &operator=(const ResidentBases &) = delete;
    ~ResidentBases() { b200msm_bases_free(h_); }
    usize len() const { return n_; }
    bool precompute(int window_bits = 0) { return b200msm_bases_precompute(h_, window_bits) == 0; }
    // same error convention as VariableBaseMSM::msm: more scalars than bases → Err(min(len)), device error → Err(0)
    Result<Proj> msm(const Scalar *scalars, usize ns) const { return run(reinterpret_cast<const uint64_t *>(scalars), ns, 1); }
    Result<Proj> msm_bigint(const BigInt4 *bigints, usize ns) const { return run(reinterpret_cast<const uint64_t *>(bigints), ns, 0); }

  private:
    Result<Proj> run(const uint64_t *scalars, usize ns, int mont) const {
        if (ns > n_) return Result<Proj>::Err(n_);
        Proj out{};
        if (b200msm_run(h_, scalars, ns, mont, out.l) != 0) return Result<Proj>::Err(0);
        return Result<Proj>::Ok(out);
    }
};

}  // namespace ark_blst
