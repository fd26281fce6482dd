"""XYZZ point kernels (madd / add / dbl, and the quad-distributed add / dbl = ops 3 / 4) against the
oracle, including every exceptional case."""
import random

import numpy as np
import pytest

from helpers import curve, points_to_limbs
from oracle import bls12381 as o

pytestmark = pytest.mark.gpu


def _xyzz_limbs(C, pt, rng):
    """random XYZZ representative of affine pt: (x·z², y·z³, z², z³); infinity = zeros"""
    F = C.F
    if pt is None:
        return [0] * (4 * F.nlimbs64)
    z = (rng.randrange(1, o.P), rng.randrange(o.P)) if F is o.Fp2Ops else rng.randrange(1, o.P)
    zz = F.sqr(z)
    zzz = F.mul(zz, z)
    return F.to_limbs(F.mul(pt[0], zz)) + F.to_limbs(F.mul(pt[1], zzz)) + F.to_limbs(zz) + F.to_limbs(zzz)


def _cases(C, rng, n):
    g = C.gen
    pts = [C.mul(g, rng.randrange(1, o.R_ORDER)) for _ in range(n)]
    qs = [C.mul(g, rng.randrange(1, o.R_ORDER)) for _ in range(n)]
    # exceptional cases: acc = inf, q = inf, q = acc (doubling), q = -acc (cancel), both inf
    pts[0] = None
    qs[1] = None
    qs[2] = pts[2]
    qs[3] = C.neg(pts[3])
    pts[4] = None
    qs[4] = None
    return pts, qs


@pytest.mark.parametrize("g2", [0, 1])
@pytest.mark.parametrize("op", [0, 1, 2, 3, 4])
def test_point_ops(eng, g2, op):
    C = curve(g2)
    rng = random.Random(300 + 10 * g2 + op)
    n = 24
    pts, qs = _cases(C, rng, n)
    acc = np.array([_xyzz_limbs(C, p, rng) for p in pts], dtype=np.uint64)
    if op == 0:
        q = points_to_limbs(C, qs)
    else:
        q = np.array([_xyzz_limbs(C, p, rng) for p in qs], dtype=np.uint64)
    out = np.zeros((n, 3 * C.F.nlimbs64), dtype=np.uint64)
    u = eng._lib.u64p
    rc = eng._lib.lib.b200msm_dbg_point_op(g2, op, acc.ctypes.data_as(u), q.ctypes.data_as(u), out.ctypes.data_as(u), n)
    assert rc == 0, eng._lib.lib.b200msm_last_error()
    for i in range(n):
        exp = C.add_affine(pts[i], qs[i]) if op in (0, 1, 3) else C.add_affine(pts[i], pts[i])
        got = C.jac_from_limbs([int(v) for v in out[i]])
        assert C.eq(got, exp), (i, op)
