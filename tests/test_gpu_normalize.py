"""§8f-3: batch normalisation to affine on the device vs the oracle (limb-exact: the affine
Montgomery image is canonical).  Mirrors CurveGroup::normalize_batch (src/g1.rs:536-543),
exercised by the reference at src/tests.rs:63."""
import random

import numpy as np
import pytest

from helpers import curve
from oracle import bls12381 as o

pytestmark = pytest.mark.gpu


def _rand_jac(C, pt, rng):
    F = C.F
    n = F.nlimbs64
    if pt is None:
        # identity: Z = 0 with arbitrary X, Y (blst leaves them unspecified)
        x = F.to_limbs((rng.randrange(o.P), rng.randrange(o.P)) if F is o.Fp2Ops else rng.randrange(o.P))
        return x + x + [0] * n
    z = (rng.randrange(1, o.P), rng.randrange(o.P)) if F is o.Fp2Ops else rng.randrange(1, o.P)
    z2 = F.sqr(z)
    return F.to_limbs(F.mul(pt[0], z2)) + F.to_limbs(F.mul(pt[1], F.mul(z2, z))) + F.to_limbs(z)


@pytest.mark.parametrize("g2,n", [(0, 1), (0, 2), (0, 777), (0, 50000), (1, 3), (1, 3001)])
def test_normalize_batch(eng, cref, g2, n):
    C = curve(g2)
    rng = random.Random(600 + n + g2)
    base_pts = [C.mul(C.gen, rng.randrange(1, o.R_ORDER)) for _ in range(min(n, 40))]
    pts = [base_pts[i % len(base_pts)] for i in range(n)]
    for i in range(0, n, 97):        # sprinkle identities, including first/last where they fall
        pts[i] = None
    if n > 5:
        pts[-1] = None
    proj = np.array([_rand_jac(C, p, rng) for p in pts], dtype=np.uint64)
    grp = eng.G2Projective if g2 else eng.G1Projective
    got = grp.normalize_batch(proj)
    exp = np.array([C.affine_to_limbs(p) for p in pts], dtype=np.uint64)
    assert np.array_equal(got, exp)


def test_normalize_then_msm_roundtrip(eng, cref):
    """the reference's own sequence: normalize_batch(bases) then msm(affines, scalars) (src/tests.rs:63-67)"""
    C = o.G1
    rng = random.Random(9)
    pts = [C.mul(C.gen, rng.randrange(1, o.R_ORDER)) for _ in range(10)]
    sc = [rng.randrange(o.R_ORDER) for _ in range(10)]
    proj = np.array([_rand_jac(C, p, rng) for p in pts], dtype=np.uint64)
    aff = eng.G1Projective.normalize_batch(proj)
    got = eng.G1Projective.msm(aff, np.array([o.scalar_to_limbs(s, True) for s in sc], dtype=np.uint64))
    assert C.eq(C.jac_from_limbs([int(v) for v in got]), C.msm_naive(pts, sc))
