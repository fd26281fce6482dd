// C++ twin of the reference's own MSM tests, through the host mirror (ark_blst_msm.hpp):
//   group_test MSM part        reference src/tests.rs:50-67   (10 bases × 10 scalars == naive fold)
//   custom_msm (3 identities)  reference src/g1.rs:695-709
//   length mismatch → Err(min) arkworks convention
// The expected values come from the CPU oracle (oracle/libmsm_ref.so), used as the checker only.
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../ark_blst_b200/host/ark_blst_msm.hpp"

extern "C" {
void ref_g1_msm_naive(const uint64_t *, const uint64_t *, size_t, int, uint64_t *);
void ref_g2_msm_naive(const uint64_t *, const uint64_t *, size_t, int, uint64_t *);
void ref_g1_to_affine(const uint64_t *, uint64_t *);
void ref_g2_to_affine(const uint64_t *, uint64_t *);
void ref_synth_scalars(uint64_t, size_t, int, uint64_t *);
void ref_synth_bases(int, uint64_t, size_t, const uint64_t *, uint64_t *, int);
}
using namespace ark_blst;

static int failures = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); failures++; } } while (0)

template <class P, int AW> static bool same_point(const P &a, const uint64_t *jac_expected, void (*to_aff)(const uint64_t *, uint64_t *)) {
    uint64_t x[AW], y[AW];
    to_aff(a.l, x);
    to_aff(jac_expected, y);
    return std::memcmp(x, y, sizeof x) == 0;  // equality after normalisation to affine
}

int main(int argc, char **argv) {
    // generators are passed in by the pytest driver as hex limbs (argv[1] = g1, argv[2] = g2)
    if (argc < 3) { std::printf("usage: test g1gen_hex g2gen_hex\n"); return 2; }
    uint64_t g1[12], g2[24];
    for (int i = 0; i < 12; i++) std::sscanf(argv[1] + 16 * i, "%16lx", &g1[i]);
    for (int i = 0; i < 24; i++) std::sscanf(argv[2] + 16 * i, "%16lx", &g2[i]);

    {   // bind every visible GPU (the reference takes devices[0], src/gpu.rs:233-234): a sharded MSM when there are several
        auto nd = init_devices(0, 0);
        CHECK(nd.is_ok() && nd.unwrap() >= 1 && device_count() == nd.unwrap());
        CHECK(init_devices(0, 0).is_ok());                   // same binding again is fine
        std::printf("devices bound: %zu\n", device_count());
    }
    {   // G1: group_test MSM part
        const size_t n = 10;
        std::vector<G1Affine> bases(n + 3);
        std::vector<Scalar> scalars(n + 3);
        ref_synth_bases(0, 4242, n, g1, bases[0].l, 1);
        ref_synth_scalars(4343, n + 3, 1, scalars[0].l);
        uint64_t exp[18];
        ref_g1_msm_naive(bases[0].l, scalars[0].l, n, 1, exp);
        auto res = G1Projective::msm(bases.data(), n, scalars.data(), n);
        CHECK(res.is_ok());
        CHECK((same_point<G1Projective, 12>(res.unwrap(), exp, ref_g1_to_affine)));
        // custom_msm: + 3 identity bases (all-zero affine), result stays the same and non-zero
        std::memset(&bases[n], 0, 3 * sizeof(G1Affine));
        auto res2 = G1Projective::msm(bases.data(), n + 3, scalars.data(), n + 3);
        CHECK(res2.is_ok() && !res2.unwrap().is_zero());
        CHECK((same_point<G1Projective, 12>(res2.unwrap(), exp, ref_g1_to_affine)));
        // msm_bigint on canonical limbs
        std::vector<BigInt4> big(n);
        ref_synth_scalars(4343, n, 0, big[0].l);
        CHECK((same_point<G1Projective, 12>(G1Projective::msm_bigint(bases.data(), n, big.data(), n), exp, ref_g1_to_affine)));
        // length mismatch
        auto bad = G1Projective::msm(bases.data(), 7, scalars.data(), 5);
        CHECK(bad.is_err() && bad.unwrap_err() == 5);
        // empty
        auto empty = G1Projective::msm(bases.data(), 0, scalars.data(), 0);
        CHECK(empty.is_ok() && empty.unwrap().is_zero());
    }
    {   // G2: group_test MSM part
        const size_t n = 10;
        std::vector<G2Affine> bases(n);
        std::vector<Scalar> scalars(n);
        ref_synth_bases(1, 5252, n, g2, bases[0].l, 1);
        ref_synth_scalars(5353, n, 1, scalars[0].l);
        uint64_t exp[36];
        ref_g2_msm_naive(bases[0].l, scalars[0].l, n, 1, exp);
        auto res = G2Projective::msm(bases.data(), n, scalars.data(), n);
        CHECK(res.is_ok());
        CHECK((same_point<G2Projective, 24>(res.unwrap(), exp, ref_g2_to_affine)));
    }
    {   // resident bases, plain and as a fixed-base window table: same group element as msm()
        const size_t n = 300;
        std::vector<G1Affine> bases(n);
        std::vector<Scalar> scalars(n);
        ref_synth_bases(0, 6262, n, g1, bases[0].l, 1);
        ref_synth_scalars(6363, n, 1, scalars[0].l);
        auto direct = G1Projective::msm(bases.data(), n, scalars.data(), n);
        CHECK(direct.is_ok());
        ResidentBases<G1Projective> pk(bases.data(), n);
        auto r1 = pk.msm(scalars.data(), n);
        CHECK(r1.is_ok() && (same_point<G1Projective, 12>(r1.unwrap(), direct.unwrap().l, ref_g1_to_affine)));
        CHECK(pk.precompute());
        auto r2 = pk.msm(scalars.data(), n);
        CHECK(r2.is_ok() && (same_point<G1Projective, 12>(r2.unwrap(), direct.unwrap().l, ref_g1_to_affine)));
        auto r3 = pk.msm(scalars.data(), 100);   // prefix
        auto d3 = G1Projective::msm(bases.data(), 100, scalars.data(), 100);
        CHECK(r3.is_ok() && (same_point<G1Projective, 12>(r3.unwrap(), d3.unwrap().l, ref_g1_to_affine)));
        std::vector<Scalar> more(n + 1);
        CHECK(pk.msm(more.data(), n + 1).is_err() && pk.msm(more.data(), n + 1).unwrap_err() == n);
    }
    std::printf(failures ? "FAILED (%d)\n" : "ok\n", failures);
    return failures ? 1 : 0;
}
