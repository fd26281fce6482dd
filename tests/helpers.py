"""Shared input builders for the parity tests (numpy uint64 in the reference's layouts)."""
import random

import numpy as np

from oracle import bls12381 as o


def rand_fp_limbs(rng, n, edge=True):
    vals = [rng.randrange(o.P) for _ in range(n)]
    if edge and n >= 6:
        vals[:6] = [0, 1, o.P - 1, o.P - 2, (o.P - 1) // 2, o.MONT_RINV]
    return vals, np.array([o.FpOps.to_limbs(v) for v in vals], dtype=np.uint64)


def rand_fp2_limbs(rng, n, edge=True):
    vals = [(rng.randrange(o.P), rng.randrange(o.P)) for _ in range(n)]
    if edge and n >= 6:
        vals[:6] = [(0, 0), (1, 0), (0, 1), (o.P - 1, o.P - 1), (0, o.P - 1), (o.P - 1, 0)]
    return vals, np.array([o.Fp2Ops.to_limbs(v) for v in vals], dtype=np.uint64)


def curve(g2):
    return o.G2 if g2 else o.G1


def scalars_to_limbs(scalars, mont):
    return np.array([o.scalar_to_limbs(s, mont) for s in scalars], dtype=np.uint64).reshape(-1, 4)


def raw_bigints_to_limbs(values):
    """un-reduced 256-bit integers as BigInt<4> limbs (msm_bigint may be fed these)"""
    return np.array([o.int_to_limbs(v, 4) for v in values], dtype=np.uint64).reshape(-1, 4)


def points_to_limbs(C, pts):
    w = 2 * C.F.nlimbs64
    return np.array([C.affine_to_limbs(p) for p in pts], dtype=np.uint64).reshape(-1, w)


def jac_to_point(C, limbs):
    return C.jac_from_limbs([int(v) for v in limbs])
