"""CPU suite: pins the oracle (T0 big-int, T1 C port) on every constant the reference holds for
this path and on the committed golden vectors. No GPU, no compute calls into libb200msm."""
import json
import os
import random
import re

import numpy as np
import pytest

from helpers import curve, points_to_limbs, scalars_to_limbs
from oracle import bls12381 as o

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def _golden():
    with open(os.path.join(HERE, "golden", "msm_vectors.json")) as f:
        return json.load(f)


def _dec_pt(C, v):
    if v is None:
        return None
    if C is o.G2:
        return ((int(v[0][0], 16), int(v[0][1], 16)), (int(v[1][0], 16), int(v[1][1], 16)))
    return (int(v[0], 16), int(v[1], 16))


# ---- constants the reference itself pins (literals restated; file:line cited) ----
def test_reference_constants():
    # src/fp.rs:25-32 modulus limbs, src/scalar.rs:476-481 Fr modulus limbs
    assert o.P == 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
    assert o.R_ORDER == 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
    x = o.BLS_X
    assert o.R_ORDER == x**4 - x**2 + 1 and o.P == (x - 1) ** 2 * o.R_ORDER // 3 + x
    # src/g1.rs:42 cofactor = (x-1)^2/3 ; src/g1.rs:44-51 COFACTOR_INV (raw Montgomery limbs)
    assert o.G1_COFACTOR == (x - 1) ** 2 // 3 == 76329603384216526031706109802092473003
    inv_limbs = [288839107172787499, 1152722415086798946, 2612889808468387987, 5124657601728438008]
    assert o.scalar_from_limbs(inv_limbs, True) == pow(o.G1_COFACTOR, -1, o.R_ORDER)
    assert o.scalar_from_limbs(inv_limbs, True) == 52435875175126190458656871551744051925719901746859129887267498875565241663483
    # src/g2.rs:42-63
    assert o.G2_COFACTOR == 305502333931268344200999753193121504214466019254188142667664032982267604182971884026507427359259977847832272839041616661285803823378372096355777062779109
    inv2 = [6746407649509787816, 1304054119431494378, 2461312685643913071, 5956596749362435284]
    assert o.scalar_from_limbs(inv2, True) == pow(o.G2_COFACTOR, -1, o.R_ORDER)


def test_reference_montgomery_kat():
    """src/fp.rs:714-721: raw limbs of blst_fp for (p-1)/2 — the reference's only hard-coded KAT;
    it pins R = 2^384, little-endian u64 limb order, and to_bytes_le byte order."""
    kat = [0xA1FAFFFFFFFE5557, 0x995BFFF976A3FFFE, 0x03F41D24D174CEB4, 0xF6547998C1995DBD, 0x778A468F507A6034, 0x020559931F7F8103]
    assert o.FpOps.to_limbs((o.P - 1) // 2) == kat
    assert o.FpOps.from_limbs(kat) == (o.P - 1) // 2
    # src/fp.rs:482-491: raw limbs p-1 (Montgomery) stand for the 2-adic root of unity -1·R^-1... as stored
    tw = [0xB9FEFFFFFFFFAAAA, 0x1EABFFFEB153FFFF, 0x6730D2A0F6B0F624, 0x64774B84F38512BF, 0x4B1BA7B6434BACD7, 0x1A0111EA397FE69A]
    assert o.limbs_to_int(tw) == o.P - 1


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_constants_against_reference_source_text():
    """when /root/reference is mounted, read the literals out of the Rust text itself"""
    def hexes(path, lo, hi):
        lines = open(os.path.join(REF, path)).read().split("\n")[lo - 1 : hi]
        return [int(h.replace("_", ""), 16) for h in re.findall(r"0x[0-9a-fA-F_]+", "\n".join(lines))]

    assert hexes("src/fp.rs", 25, 32) == o.P_LIMBS
    assert hexes("src/scalar.rs", 476, 481) == o.R_LIMBS
    assert hexes("src/fp.rs", 714, 721) == o.FpOps.to_limbs((o.P - 1) // 2)
    assert o.limbs_to_int(hexes("src/g1.rs", 42, 42)) == o.G1_COFACTOR
    assert o.limbs_to_int(hexes("src/g2.rs", 46, 53)) == o.G2_COFACTOR


def test_generators_and_external_kat():
    assert o.G1.is_on_curve(o.G1_GEN) and o.G2.is_on_curve(o.G2_GEN)
    assert o.G1.mul(o.G1_GEN, o.R_ORDER) is None and o.G2.mul(o.G2_GEN, o.R_ORDER) is None
    # 2·G1 as published in the EIP-2537 / IETF BLS test material (independent of this repo)
    d = o.G1.mul(o.G1_GEN, 2)
    assert d[0] == 0x0572CBEA904D67468808C8EB50A9450C9721DB309128012543902D0AC358A62AE28F75BB8F1C7C42C39A8C5529BF0F4E
    assert d[1] == 0x166A9D8CABC673A322FDA673779D8E3822BA3ECB8670E461F73BB9021D5FD76A4C56D9D4CD16BD1BBA86881979749D28


def test_group_laws_t0():
    rng = random.Random(3)
    for C in (o.G1, o.G2):
        a, b = rng.randrange(o.R_ORDER), rng.randrange(o.R_ORDER)
        A, B = C.mul(C.gen, a), C.mul(C.gen, b)
        assert C.eq(C.add_affine(A, B), C.mul(C.gen, (a + b) % o.R_ORDER))
        assert C.eq(C.add_affine(A, B), C.add_affine(B, A))            # src/tests.rs:31-34
        assert C.add_affine(A, C.neg(A)) is None                        # src/tests.rs:32
        assert C.eq(C.from_jac(C.jac_add(C.to_jac(A), C.to_jac(A))), C.mul(C.gen, 2 * a % o.R_ORDER))


# ---- golden vectors: T0 regenerates them, T1 (C port) reproduces them ----
def test_golden_constants():
    g = _golden()["constants"]
    assert int(g["p"], 16) == o.P and int(g["r"], 16) == o.R_ORDER
    assert [int(v, 16) for v in g["fp_mont_one_limbs"]] == o.int_to_limbs(o.MONT_R, 6)
    assert [int(v, 16) for v in g["fp_mont_r2_limbs"]] == o.int_to_limbs(o.MONT_R2, 6)
    assert [int(v, 16) for v in g["fr_mont_one_limbs"]] == [0x00000001FFFFFFFE, 0x5884B7FA00034802, 0x998C4FEFECBC4FF5, 0x1824B159ACC5056F]


@pytest.mark.parametrize("idx", range(16))
def test_golden_cases_t0_and_t1(cref, idx):
    case = _golden()["cases"][idx]
    g2 = case["group"] == "g2"
    C = curve(g2)
    bases = np.array([[int(v, 16) for v in row] for row in case["bases_limbs"]], dtype=np.uint64)
    sc = np.array([[int(v, 16) for v in row] for row in case["scalars_canonical_limbs"]], dtype=np.uint64)
    sm = np.array([[int(v, 16) for v in row] for row in case["scalars_montgomery_limbs"]], dtype=np.uint64)
    exp = _dec_pt(C, case["result_affine"])
    # T0 from the byte images
    pts = [C.affine_from_limbs(r) for r in bases.tolist()]
    assert C.eq(C.msm_naive(pts, [o.scalar_from_limbs(r, False) for r in sc.tolist()]), exp)
    assert [o.scalar_from_limbs(r, True) for r in sm.tolist()] == [o.scalar_from_limbs(r, False) for r in sc.tolist()]
    # T1, both scalar forms, Pippenger and naive, 1 and many threads
    for mont, s in ((0, sc), (1, sm)):
        for kw in ({}, {"nthreads": 1, "window": 4}):
            assert C.eq(C.jac_from_limbs(cref.msm(g2, bases, s, mont, **kw).tolist()), exp)
        assert C.eq(C.jac_from_limbs(cref.msm_naive(g2, bases, s, mont).tolist()), exp)


# ---- T1 vs T0, limb for limb ----
def test_c_field_ops_vs_bigint(cref):
    rng = random.Random(11)
    for _ in range(300):
        a, b = rng.randrange(o.P), rng.randrange(o.P)
        A = np.array(o.FpOps.to_limbs(a), dtype=np.uint64)
        B = np.array(o.FpOps.to_limbs(b), dtype=np.uint64)
        assert cref.fp_binop("ref_fp_mul", A, B).tolist() == o.FpOps.to_limbs(a * b % o.P)
        assert cref.fp_binop("ref_fp_add", A, B).tolist() == o.FpOps.to_limbs((a + b) % o.P)
        assert cref.fp_binop("ref_fp_sub", A, B).tolist() == o.FpOps.to_limbs((a - b) % o.P)
    for a, b in ((0, 0), (o.P - 1, o.P - 1), (1, o.P - 1), (0, 5)):
        A = np.array(o.FpOps.to_limbs(a), dtype=np.uint64)
        B = np.array(o.FpOps.to_limbs(b), dtype=np.uint64)
        assert cref.fp_binop("ref_fp_mul", A, B).tolist() == o.FpOps.to_limbs(a * b % o.P)
        assert cref.fp_binop("ref_fp_sub", A, B).tolist() == o.FpOps.to_limbs((a - b) % o.P)
    for _ in range(100):
        a = (rng.randrange(o.P), rng.randrange(o.P))
        b = (rng.randrange(o.P), rng.randrange(o.P))
        A = np.array(o.Fp2Ops.to_limbs(a), dtype=np.uint64)
        B = np.array(o.Fp2Ops.to_limbs(b), dtype=np.uint64)
        assert cref.fp_binop("ref_fp2_mul", A, B).tolist() == o.Fp2Ops.to_limbs(o.Fp2Ops.mul(a, b))


def test_synth_streams_agree(cref):
    n = 40
    assert [o.limbs_to_int(r) for r in cref.synth_scalars(2, n, False).tolist()] == o.synth_scalars(2, n)
    assert [o.scalar_from_limbs(r, True) for r in cref.synth_scalars(2, n, True).tolist()] == o.synth_scalars(2, n)
    assert [o.limbs_to_int(r) for r in cref.synth_dlogs(3, n).tolist()] == [o.synth_dlog(3, i) for i in range(n)]
    for g2 in (0, 1):
        C = curve(g2)
        assert cref.synth_bases(g2, 3, 12).tolist() == [C.affine_to_limbs(p) for p in o.synth_bases(C, 3, 12)]


def test_booth_digits(cref):
    rng = random.Random(5)
    for c in (2, 3, 8, 13, 16, 17, 20, 22):
        W = (256 + c - 1) // c
        for s in [0, 1, o.R_ORDER - 1, (1 << 255) - 1] + [rng.randrange(o.R_ORDER) for _ in range(20)]:
            lim = np.array(o.int_to_limbs(s, 4), dtype=np.uint64)
            d = [cref.lib().ref_booth_digit(cref._p(lim), w, c) for w in range(W)]
            assert sum(v << (c * w) for w, v in enumerate(d)) == s
            assert all(abs(v) <= 1 << (c - 1) for v in d)


@pytest.mark.parametrize("g2,n", [(0, 0), (0, 1), (0, 257), (0, 4096), (1, 300)])
def test_c_pippenger_vs_naive_and_dlog(cref, g2, n):
    bases = cref.synth_bases(g2, 21, n)
    sc = cref.synth_scalars(22, n, False)
    sm = cref.synth_scalars(22, n, True)
    r = cref.msm(g2, bases, sm, 1)
    assert cref.affine_equal(g2, r, cref.msm(g2, bases, sc, 0, nthreads=1))
    if n <= 300:
        assert cref.affine_equal(g2, r, cref.msm_naive(g2, bases, sc, 0))
    if n:
        assert cref.affine_equal(g2, r, cref.msm_by_dlog(g2, 21, sc))
    else:
        assert not r.any()


def test_c_identity_and_edge_scalars(cref):
    C = o.G1
    rng = random.Random(8)
    P = [C.mul(C.gen, rng.randrange(1, o.R_ORDER)) for _ in range(5)]
    pts = P + [None, None, P[0], C.neg(P[1])]
    sc = [0, 1, o.R_ORDER - 1, 12345, 1 << 254, 7, 9, 0, 1]
    exp = C.msm_naive(pts, sc)
    got = cref.msm(0, points_to_limbs(C, pts), scalars_to_limbs(sc, True), 1)
    assert C.eq(C.jac_from_limbs(got.tolist()), exp)


# ---- §8f-4 encodings: the oracle against published vectors and itself ----
def test_point_encoding_kats_and_roundtrip():
    g1 = "97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb"
    g2 = ("93e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e"
          "024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8")
    assert o.serialize_point(o.G1, o.G1_GEN, True).hex() == g1     # published zkcrypto / IETF encodings
    assert o.serialize_point(o.G2, o.G2_GEN, True).hex() == g2
    rng = random.Random(77)
    for C in (o.G1, o.G2):
        for _ in range(4):
            P_ = C.mul(C.gen, rng.randrange(1, o.R_ORDER))
            for comp in (True, False):
                st, q = o.deserialize_point(C, o.serialize_point(C, P_, comp), comp, True)
                assert st == 0 and C.eq(q, P_)
        for comp in (True, False):
            st, q = o.deserialize_point(C, o.serialize_point(C, None, comp), comp, True)
            assert st == 0 and q is None
    # Fp2 square roots: every square has a root, non-squares are rejected
    for _ in range(20):
        a = (rng.randrange(o.P), rng.randrange(o.P))
        s = o.fp2_sqrt(o.Fp2Ops.sqr(a))
        assert s is not None and o.Fp2Ops.eq(o.Fp2Ops.sqr(s), o.Fp2Ops.sqr(a))


# ---- round 2: the MULX/ADCX/ADOX product (the timed CPU baseline's inner loop) and blst's tile grid ----
def test_adx_product_equals_portable_and_bigint(cref):
    """oracle/field.h fp_mul_adx (inline asm, what the timed baseline runs on an ADX+BMI2 host) against the
    portable u128 product and the big-int oracle, limb for limb, incl. the operands at the edges of [0, p)"""
    L = cref.lib()
    rng = random.Random(29)
    edge = [0, 1, 2, o.P - 1, o.P - 2, (o.P - 1) // 2, o.MONT_R % o.P, o.MONT_RINV, (1 << 380) % o.P, (1 << 64) - 1, 1 << 64]
    pairs = [(a, b) for a in edge for b in edge] + [(rng.randrange(o.P), rng.randrange(o.P)) for _ in range(500)]
    for a, b in pairs:
        A = np.array(o.int_to_limbs(a, 6), dtype=np.uint64)   # raw limbs: the product is a·b·2^-384 of the raw values
        B = np.array(o.int_to_limbs(b, 6), dtype=np.uint64)
        want = o.int_to_limbs(a * b * o.MONT_RINV % o.P, 6)
        assert cref.fp_binop("ref_fp_mul_portable", A, B).tolist() == want
        if L.ref_mul_impl():
            assert cref.fp_binop("ref_fp_mul_adx", A, B).tolist() == want
        assert cref.fp_binop("ref_fp_mul", A, B).tolist() == want


def test_blst_tile_grid_restated(cref):
    """the (slices, windows, width) grid of blst's multi-threaded Pippenger driver as restated in
    oracle/msm_ref.c `breakdown`: widths cover the 255 scalar bits, and few CPUs never slice"""
    import ctypes

    L = cref.lib()
    for cpus in (2, 8, 16, 32, 64, 128, 256):
        for logn in range(5, 27):
            out = (ctypes.c_uint * 3)()
            w = L.ref_window_rule(1 << logn)
            L.ref_breakdown(255, w, cpus, out)
            nx, ny, wnd = list(out)
            assert nx >= 1 and ny * wnd >= 256 and abs(int(wnd) - int(w)) <= 8
            if 255 > w * cpus:
                assert nx == 1
    out = (ctypes.c_uint * 3)()
    L.ref_breakdown(255, 17, 16, out)          # 2^20 points on the 16-core bench box: 16 full-length tiles of width 16
    assert list(out) == [1, 16, 16]


def test_generator_literals_published():
    """the standard generators (IETF pairing-friendly-curves draft §4.2.1 / zkcrypto bls12_381), restated
    literally: pins the oracle's G1_GEN / G2_GEN from outside this repository"""
    assert o.G1_GEN == (
        0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
        0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1)
    assert o.G2_GEN == (
        (0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
         0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E),
        (0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
         0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE))
