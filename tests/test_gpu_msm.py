"""MSM parity: CUDA path (through the C-ABI) vs the oracle, bit-exact after affine normalisation —
the relation the reference's own test asserts (src/tests.rs:50-67)."""
import random

import numpy as np
import pytest

from helpers import curve, jac_to_point, points_to_limbs, raw_bigints_to_limbs, scalars_to_limbs
from oracle import bls12381 as o

pytestmark = pytest.mark.gpu


def _grp(eng, g2):
    return eng.G2Projective if g2 else eng.G1Projective


def test_digits_match_oracle(eng, cref):
    import ctypes

    n = 512
    sc = cref.synth_scalars(5, n, False)
    sc[0] = 0
    sc[1] = np.array(o.int_to_limbs(o.R_ORDER - 1, 4), dtype=np.uint64)
    sc[2] = np.array(o.int_to_limbs(1, 4), dtype=np.uint64)
    scm = np.array([o.scalar_to_limbs(o.limbs_to_int(r), True) for r in sc.tolist()], dtype=np.uint64)
    for c in (2, 7, 13, 16, 17, 20, 24):
        for mont, arr in ((0, sc), (1, scm)):
            nwin = ctypes.c_int()
            W = (256 + c - 1) // c
            out = np.zeros((W, n), dtype=np.int32)
            rc = eng._lib.lib.b200msm_dbg_digits(arr.ctypes.data_as(eng._lib.u64p), n, mont, c, out.ctypes.data_as(eng._lib.i32p), ctypes.byref(nwin))
            assert rc == 0 and nwin.value == W
            for i in range(0, n, 37):
                s = o.limbs_to_int(sc[i].tolist())
                digs = [int(out[w, i]) for w in range(W)]
                assert sum(d << (c * w) for w, d in enumerate(digs)) == s
                assert all(abs(d) <= 1 << (c - 1) for d in digs)
                assert digs == [cref.lib().ref_booth_digit(sc[i].ctypes.data_as(eng._lib.u64p), w, c) for w in range(W)]


@pytest.mark.parametrize("g2", [0, 1])
def test_reference_group_test_shape(eng, g2):
    """10 random bases × 10 random scalars vs the naive fold — src/tests.rs:50-67."""
    C = curve(g2)
    rng = random.Random(42 + g2)
    pts = [C.mul(C.gen, rng.randrange(1, o.R_ORDER)) for _ in range(10)]
    sc = [rng.randrange(o.R_ORDER) for _ in range(10)]
    exp = C.msm_naive(pts, sc)
    got = _grp(eng, g2).msm(points_to_limbs(C, pts), scalars_to_limbs(sc, True))
    assert C.eq(jac_to_point(C, got), exp)
    got = _grp(eng, g2).msm_bigint(points_to_limbs(C, pts), scalars_to_limbs(sc, False))
    assert C.eq(jac_to_point(C, got), exp)


@pytest.mark.parametrize("g2", [0, 1])
def test_identity_bases(eng, g2):
    """10 random + 3 identity bases (src/g1.rs:695-709): the case the reference's blst arm fails."""
    C = curve(g2)
    rng = random.Random(77 + g2)
    pts = [C.mul(C.gen, rng.randrange(1, o.R_ORDER)) for _ in range(10)] + [None] * 3
    sc = [rng.randrange(o.R_ORDER) for _ in range(13)]
    exp = C.msm_naive(pts, sc)
    assert exp is not None
    got = _grp(eng, g2).msm(points_to_limbs(C, pts), scalars_to_limbs(sc, True))
    assert C.eq(jac_to_point(C, got), exp)


@pytest.mark.parametrize("g2", [0, 1])
def test_edge_cases(eng, g2):
    C = curve(g2)
    G = _grp(eng, g2)
    rng = random.Random(91 + g2)
    w = 2 * C.F.nlimbs64
    # empty input → identity
    out = G.msm(np.zeros((0, w), dtype=np.uint64), np.zeros((0, 4), dtype=np.uint64))
    assert jac_to_point(C, out) is None
    P = C.mul(C.gen, rng.randrange(1, o.R_ORDER))
    Q = C.mul(C.gen, rng.randrange(1, o.R_ORDER))
    cases = [
        ([P], [0]),                                    # zero scalar → identity
        ([P], [1]),
        ([P], [o.R_ORDER - 1]),                        # = -P
        ([P, P], [5, o.R_ORDER - 5]),                  # cancels to identity
        ([P, C.neg(P)], [7, 7]),                       # P + (-P) inside one bucket
        ([P, P, P, P], [3, 3, 3, 3]),                  # duplicates → doubling inside a bucket
        ([P, Q, P, Q, P], [1, 1, 1, 1, 1]),            # all-equal scalars
        ([P, Q], [(1 << 255) - 19 - o.R_ORDER, 2]),
        ([None, None], [3, 4]),                        # only identities
    ]
    for pts, sc in cases:
        exp = C.msm_naive(pts, sc)
        for mont in (True, False):
            got = (G.msm if mont else G.msm_bigint)(points_to_limbs(C, pts), scalars_to_limbs(sc, mont))
            assert C.eq(jac_to_point(C, got), exp), (pts, sc, mont)
    # msm_bigint with un-reduced 256-bit integers (≥ r): reduced mod r on the device
    raw = [o.R_ORDER, o.R_ORDER + 5, (1 << 256) - 1, 2 * o.R_ORDER + 3]
    pts = [P, Q, P, Q]
    exp = C.msm_naive(pts, [v % o.R_ORDER for v in raw])
    got = G.msm_bigint(points_to_limbs(C, pts), raw_bigints_to_limbs(raw))
    assert C.eq(jac_to_point(C, got), exp)
    # length mismatch → Err(min(len)), arkworks' convention
    with pytest.raises(eng.MsmError) as ei:
        G.msm(points_to_limbs(C, [P, Q]), scalars_to_limbs([1], True))
    assert ei.value.value == 1


@pytest.mark.parametrize("g2,n", [(0, 1), (0, 2), (0, 33), (0, 1000), (0, 1 << 12), (0, (1 << 14) + 3), (1, 500), (1, 1 << 12)])
def test_vs_c_oracle(eng, cref, g2, n):
    """seeded synthetic inputs at sizes the CPU oracle finishes in seconds; ragged sizes included"""
    bases = cref.synth_bases(g2, 1000 + n, n)
    sm = cref.synth_scalars(2000 + n, n, True)
    sc = cref.synth_scalars(2000 + n, n, False)
    exp = cref.msm(g2, bases, sc, 0)
    got = _grp(eng, g2).msm(bases, sm)
    assert cref.affine_equal(g2, got, exp)
    got = _grp(eng, g2).msm_bigint(bases, sc)
    assert cref.affine_equal(g2, got, exp)
    assert cref.affine_equal(g2, cref.msm_by_dlog(g2, 1000 + n, sc), exp)


@pytest.mark.parametrize("c", [2, 5, 8, 11, 13, 15, 16, 17])
def test_every_window_width(eng, cref, c):
    n = 3000
    bases = cref.synth_bases(0, 31, n)
    sc = cref.synth_scalars(32, n, False)
    exp = cref.msm(0, bases, sc, 0)
    assert eng._lib.lib.b200msm_set_window_bits(c) == 0
    try:
        got = eng.G1Projective.msm_bigint(bases, sc)
    finally:
        eng._lib.lib.b200msm_set_window_bits(0)
    assert cref.affine_equal(0, got, exp)


def test_witness_like_scalars(eng, cref):
    """≈40 % zeros, ≈20 % ones, ≈10 % small, rest uniform: huge digit-1 bucket"""
    n = 1 << 13
    rng = random.Random(5)
    bases = cref.synth_bases(0, 55, n)
    sc = []
    for _ in range(n):
        u = rng.random()
        sc.append(0 if u < 0.4 else 1 if u < 0.6 else rng.randrange(1 << 32) if u < 0.7 else rng.randrange(o.R_ORDER))
    lim = scalars_to_limbs(sc, False)
    exp = cref.msm(0, bases, lim, 0)
    got = eng.G1Projective.msm_bigint(bases, lim)
    assert cref.affine_equal(0, got, exp)


def test_resident_bases_and_linearity(eng, cref):
    """upload once, run many; msm(P,s)+msm(P,t) = msm(P,s+t)"""
    n = 5000
    bases = cref.synth_bases(0, 9, n)
    s = cref.synth_scalars(10, n, False)
    t = cref.synth_scalars(11, n, False)
    st = np.array([o.int_to_limbs((o.limbs_to_int(a) + o.limbs_to_int(b)) % o.R_ORDER, 4) for a, b in zip(s.tolist(), t.tolist())], dtype=np.uint64)
    rb = eng.ResidentBases(eng.G1Projective, bases)
    a = rb.msm(s, montgomery=False)
    b = rb.msm(t, montgomery=False)
    ab = rb.msm(st, montgomery=False)
    # prefix run
    half = rb.msm(s[: n // 2], montgomery=False)
    rb.close()
    assert cref.affine_equal(0, cref.add(0, a, b), ab)
    assert cref.affine_equal(0, a, cref.msm(0, bases, s, 0))
    assert cref.affine_equal(0, half, cref.msm(0, bases[: n // 2], s[: n // 2], 0))


@pytest.mark.parametrize("g2", [0, 1])
def test_chunked_passes(eng, cref, g2):
    """the chunking the reference left as a TODO (src/gpu.rs:238-239): forced at small n here"""
    n = 5000 if not g2 else 1500
    bases = cref.synth_bases(g2, 41, n)
    sc = cref.synth_scalars(42, n, True)
    exp = cref.msm(g2, bases, sc, 1)
    L = eng._lib.lib
    for chunk in (n - 1, 1000, 333):
        assert L.b200msm_set_max_chunk(chunk) == 0
        try:
            got = _grp(eng, g2).msm(bases, sc)
        finally:
            L.b200msm_set_max_chunk(0)
        assert cref.affine_equal(g2, got, exp), chunk


@pytest.mark.parametrize("g2,n", [(0, 1), (0, 9), (0, 700), (0, 20000), (1, 1), (1, 9), (1, 700), (1, 6000)])
def test_glv_path(eng, cref, g2, n):
    """GLV split (k = k1 + k2·λ over P and φ(P) = (β·x, y); β² on G2) forced on AND forced off — the
    automatic plan picks either — must give the same group element; includes scalars around λ and
    identity bases"""
    bases = cref.synth_bases(g2, 71 + n, n)
    sc_int = [o.limbs_to_int(r) for r in cref.synth_scalars(72 + n, n, False).tolist()]
    lam = o.BLS_X * o.BLS_X - 1
    for k, v in enumerate([0, 1, lam - 1, lam, lam + 1, o.R_ORDER - 1, lam * lam, (1 << 128) - 1, 1 << 128]):
        if k < n:
            sc_int[k] = v % o.R_ORDER
    if n > 12:
        bases[11] = 0
    sc = scalars_to_limbs(sc_int, False)
    exp = cref.msm(g2, bases, sc, 0)
    L = eng._lib.lib
    for mode in (1, 0):
        assert L.b200msm_set_glv(mode) == 0
        try:
            got = _grp(eng, g2).msm_bigint(bases, sc)
            got_m = _grp(eng, g2).msm(bases, scalars_to_limbs(sc_int, True))
        finally:
            L.b200msm_set_glv(-1)
        assert cref.affine_equal(g2, got, exp) and cref.affine_equal(g2, got_m, exp), mode


@pytest.mark.parametrize("g2,c", [(0, 4), (0, 8), (0, 13), (0, 16), (1, 8), (1, 13), (1, 16)])
def test_glv_window_layouts(eng, cref, g2, c):
    """GLV with the width forced: c | 128 takes the top c bits of each 128-bit half unsigned and
    spreads that digit over the last two windows (no carry window), other widths keep the Booth
    carry window; halves whose top digit sits on either side of 2^(c−1), with and without the
    carry from below, are planted explicitly"""
    n = 3000 if not g2 else 1200
    lam = o.BLS_X * o.BLS_X - 1
    bases = cref.synth_bases(g2, 171 + c, n)
    sc_int = [o.limbs_to_int(r) for r in cref.synth_scalars(172 + c, n, False).tolist()]
    half = 1 << (c - 1)
    tops = [half - 1, half, half + 1, (lam >> (128 - c)) - 1, 1, 0]
    k = 0
    for t1 in tops:
        for carry in (0, 1):
            for t2 in (tops[0], tops[2]):
                k1 = (t1 << (128 - c)) | (carry << (127 - c)) | 5
                k2 = (t2 << (128 - c)) | ((1 - carry) << (127 - c)) | 9
                assert k1 < lam and k2 < lam
                sc_int[k] = (k1 + k2 * lam) % o.R_ORDER
                k += 1
    sc = scalars_to_limbs(sc_int, False)
    exp = cref.msm(g2, bases, sc, 0)
    L = eng._lib.lib
    assert L.b200msm_set_glv(1) == 0 and L.b200msm_set_window_bits(c) == 0
    try:
        got = _grp(eng, g2).msm_bigint(bases, sc)
    finally:
        L.b200msm_set_glv(-1)
        L.b200msm_set_window_bits(0)
    assert cref.affine_equal(g2, got, exp)


@pytest.mark.parametrize("g2,n,slices", [(0, 5, 8), (0, 1000, 2), (0, 20001, 4), (0, 20001, 8), (1, 3000, 4)])
def test_streamed_host_msm(eng, cref, g2, n, slices):
    """one-shot host-buffer MSM with the upload cut into slices that are accumulated into shared
    buckets (forced at small n here; the default starts at 2^18 points): same group element, with
    GLV on and off, uniform and witness-like scalars (heavy buckets fed by several slices)"""
    L = eng._lib.lib
    bases = cref.synth_bases(g2, 900 + n, n)
    rng = random.Random(n)
    uni = [o.limbs_to_int(r) for r in cref.synth_scalars(901 + n, n, False).tolist()]
    wit = [0 if (u := rng.random()) < 0.3 else 1 if u < 0.7 else rng.randrange(1 << 32) if u < 0.8 else uni[i] for i in range(n)]
    assert L.b200msm_set_stream_slices(slices, 1) == 0
    try:
        for sc_int in (uni, wit):
            sc = scalars_to_limbs(sc_int, False)
            exp = cref.msm(g2, bases, sc, 0)
            for mode in (-1, 0, 1):
                assert L.b200msm_set_glv(mode) == 0
                got = _grp(eng, g2).msm_bigint(bases, sc)
                assert cref.affine_equal(g2, got, exp), (mode, sc_int is wit)
            got = _grp(eng, g2).msm(bases, scalars_to_limbs(sc_int, True))
            assert cref.affine_equal(g2, got, exp)
            # resident bases (scalars streamed against the uploaded shard), then the same handle as a table
            L.b200msm_set_glv(-1)
            rb = eng.ResidentBases(_grp(eng, g2), bases)
            try:
                assert cref.affine_equal(g2, rb.msm(sc, montgomery=False), exp)
                rb.precompute(13 if n > 100 else 0)
                assert cref.affine_equal(g2, rb.msm(sc, montgomery=False), exp)
                if n > 10:
                    h = n // 3
                    assert cref.affine_equal(g2, rb.msm(sc[:h], montgomery=False), cref.msm(g2, bases[:h], sc[:h], 0))
            finally:
                rb.close()
    finally:
        L.b200msm_set_glv(-1)
        L.b200msm_set_stream_slices(8, 0)



def test_host_register_roundtrip(eng, cref):
    """b200msm_host_register page-locks the caller's (pageable) numpy buffers in place; results are
    unchanged, unregistering twice is an error status, never a crash"""
    L = eng._lib.lib
    n = 4000
    bases = cref.synth_bases(0, 321, n).copy()
    sc = cref.synth_scalars(322, n, True).copy()
    exp = cref.msm(0, bases, sc, 1)
    assert L.b200msm_init(-1, 1) == 0
    assert L.b200msm_host_register(bases.ctypes.data, bases.nbytes) == 0
    assert L.b200msm_host_register(sc.ctypes.data, sc.nbytes) == 0
    try:
        assert cref.affine_equal(0, eng.G1Projective.msm(bases, sc), exp)
    finally:
        assert L.b200msm_host_unregister(bases.ctypes.data) == 0
        assert L.b200msm_host_unregister(sc.ctypes.data) == 0
    assert L.b200msm_host_unregister(sc.ctypes.data) != 0
    assert L.b200msm_host_register(None, 16) != 0
    assert cref.affine_equal(0, eng.G1Projective.msm(bases, sc), exp)
