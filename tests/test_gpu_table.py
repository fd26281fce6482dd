"""Fixed-base window table for resident bases: table[w][i] = 2^(c·w)·P_i, one bucket
set for all windows, no Horner chain.  Same parity bar as the plain path: the result is the same
group element as the oracle's, bit-exact after affine normalisation (reference src/tests.rs:58-67),
and every table entry equals the oracle's 2^(c·w)·P_i limb for limb."""
import random

import numpy as np
import pytest

from helpers import curve, jac_to_point, points_to_limbs, scalars_to_limbs
from oracle import bls12381 as o

pytestmark = pytest.mark.gpu


def _grp(eng, g2):
    return eng.G2Projective if g2 else eng.G1Projective


@pytest.mark.parametrize("g2", [0, 1])
def test_table_entries_match_oracle(eng, g2):
    """every window of the device table vs big-int doubling; identity bases stay identities"""
    import torch

    C = curve(g2)
    rng = random.Random(300 + g2)
    pts = [C.mul(C.gen, rng.randrange(1, o.R_ORDER)) for _ in range(5)] + [None]
    n, aw, c = len(pts), (24 if g2 else 12), 13
    cc, W = eng.table_plan(g2, n, c)
    assert (cc, W) == (13, 20)
    bases = torch.from_numpy(points_to_limbs(C, pts).view(np.int64)).cuda()
    table = torch.zeros((W, n, aw), dtype=torch.int64, device="cuda")
    eng.table_build_device(g2, bases.data_ptr(), n, c, table.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = table.cpu().numpy().view(np.uint64)
    for w in (0, 1, 7, W - 1):
        exp = points_to_limbs(C, [None if p is None else C.mul(p, 1 << (c * w)) for p in pts])
        assert np.array_equal(got[w], exp), w


@pytest.mark.parametrize("g2,n,c", [(0, 1, 0), (0, 37, 0), (0, 3000, 11), (0, 3000, 16), (0, (1 << 14) + 5, 0), (0, 1 << 14, 20),
                                    (1, 700, 0), (1, 1 << 12, 13)])
def test_resident_table_vs_c_oracle(eng, cref, g2, n, c):
    """upload → precompute → run many; Montgomery and canonical scalars; prefix runs"""
    bases = cref.synth_bases(g2, 500 + n, n)
    if n > 20:
        bases[7] = 0  # an identity base
    sm = cref.synth_scalars(600 + n, n, True)
    sc = cref.synth_scalars(600 + n, n, False)
    exp = cref.msm(g2, bases, sc, 0)
    rb = eng.ResidentBases(_grp(eng, g2), bases)
    try:
        cc, W, nbytes = rb.precompute(c)
        assert W == (256 + cc - 1) // cc and (c == 0 or cc == c)
        assert nbytes >= W * n * bases.shape[1] * 8
        assert cref.affine_equal(g2, rb.msm(sm, montgomery=True), exp)
        assert cref.affine_equal(g2, rb.msm(sc, montgomery=False), exp)
        if n > 2:
            h = n // 2 + 1
            assert cref.affine_equal(g2, rb.msm(sc[:h], montgomery=False), cref.msm(g2, bases[:h], sc[:h], 0))
    finally:
        rb.close()


@pytest.mark.parametrize("g2", [0, 1])
def test_table_edge_scalars(eng, g2):
    """0, 1, r−1, cancellations, duplicates (doubling inside a bucket across windows), identities"""
    C = curve(g2)
    rng = random.Random(17 + g2)
    P = C.mul(C.gen, rng.randrange(1, o.R_ORDER))
    Q = C.mul(C.gen, rng.randrange(1, o.R_ORDER))
    pts = [P, Q, P, C.neg(P), None, Q, P, P]
    rb = eng.ResidentBases(_grp(eng, g2), points_to_limbs(C, pts))
    try:
        rb.precompute(10)
        cases = [
            [0] * 8,
            [1] * 8,
            [o.R_ORDER - 1] * 8,
            [5, 0, o.R_ORDER - 5, 0, 9, 0, 0, 0],            # cancels to the identity
            [7, 0, 0, 7, 0, 0, 0, 0],                        # P + (−P) in one bucket
            [3, 0, 3, 0, 0, 0, 3, 3],                        # duplicates → doubling inside a bucket
            [1 << 10, 1, 0, 0, 0, 0, 0, 0],                  # 2^c·P (window 1, bucket 1) meets Q (window 0, bucket 1)
            [(1 << 255) - 19 - o.R_ORDER, 2, 3, 4, 5, 6, 7, 8],
            [rng.randrange(o.R_ORDER) for _ in range(8)],
        ]
        for sc in cases:
            exp = C.msm_naive(pts, sc)
            for mont in (True, False):
                got = rb.msm(scalars_to_limbs(sc, mont), montgomery=mont)
                assert C.eq(jac_to_point(C, got), exp), (sc, mont)
    finally:
        rb.close()


def test_table_witness_like_scalars(eng, cref):
    """≈30 % zeros, ≈20 % ones, ≈15 % 2^c, ≈10 % small: bucket 1 collects window 0's ones and
    window 1's 2^c entries and goes down the block-cooperative heavy path"""
    n, c = 1 << 13, 12
    rng = random.Random(6)
    bases = cref.synth_bases(0, 56, n)
    sc = []
    for _ in range(n):
        u = rng.random()
        sc.append(0 if u < 0.3 else 1 if u < 0.5 else (1 << c) if u < 0.65 else rng.randrange(1 << 32) if u < 0.75 else rng.randrange(o.R_ORDER))
    lim = scalars_to_limbs(sc, False)
    exp = cref.msm(0, bases, lim, 0)
    rb = eng.ResidentBases(eng.G1Projective, bases)
    try:
        rb.precompute(c)
        assert cref.affine_equal(0, rb.msm(lim, montgomery=False), exp)
    finally:
        rb.close()


@pytest.mark.parametrize("g2", [0, 1])
def test_table_chunked_passes(eng, cref, g2):
    """forced chunking: every pass reads its slice of each window (same stride, shifted origin)"""
    n = 5000 if not g2 else 1500
    bases = cref.synth_bases(g2, 43, n)
    sc = cref.synth_scalars(44, n, True)
    exp = cref.msm(g2, bases, sc, 1)
    L = eng._lib.lib
    rb = eng.ResidentBases(_grp(eng, g2), bases)
    try:
        rb.precompute(0)
        for chunk in (n - 1, 1000, 333):
            assert L.b200msm_set_max_chunk(chunk) == 0
            try:
                got = rb.msm(sc, montgomery=True)
            finally:
                L.b200msm_set_max_chunk(0)
            assert cref.affine_equal(g2, got, exp), chunk
    finally:
        rb.close()


@pytest.mark.parametrize("g2,logn", [(0, 18), (0, 20), (1, 18)])
def test_table_dlog_closed_form(eng, cref, g2, logn):
    """Σ sᵢ·(kᵢ·G) = (Σ sᵢkᵢ mod r)·G at BASELINE sizes, table path vs plain path, inputs in HBM"""
    import torch

    n = 1 << logn
    aw = 24 if g2 else 12
    sb, ss = 0xB2000381_00001000 + logn, 177 + logn
    st = torch.cuda.current_stream().cuda_stream
    c, W = eng.table_plan(g2, n)
    table = torch.empty((W, n, aw), dtype=torch.int64, device="cuda")
    scalars = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    eng.synth_bases_device(g2, sb, n, table.data_ptr())
    eng.synth_scalars_device(ss, n, True, scalars.data_ptr())
    eng.table_build_device(g2, table.data_ptr(), n, c, table.data_ptr(), st)   # window 0 in place
    out = torch.zeros((2, 36 if g2 else 18), dtype=torch.int64, device="cuda")
    eng.run_table_device(g2, table.data_ptr(), n, c, scalars.data_ptr(), n, True, out[0].data_ptr(), st)
    eng.run_device(g2, table.data_ptr(), scalars.data_ptr(), n, True, out[1].data_ptr(), st)
    torch.cuda.synchronize()
    got = out.cpu().numpy().view(np.uint64)
    exp = cref.msm_by_dlog(g2, sb, cref.synth_scalars(ss, n, False))
    assert cref.affine_equal(g2, got[0], exp)
    assert cref.affine_equal(g2, got[1], exp)
