"""Limb-for-limb parity of the sm_100a field kernels with the big-int oracle (bit-exact)."""
import random

import numpy as np
import pytest

from helpers import rand_fp2_limbs, rand_fp_limbs
from oracle import bls12381 as o

pytestmark = pytest.mark.gpu

OPS = {0: "mul", 1: "add", 2: "sub", 3: "sqr", 4: "neg", 5: "inv"}


def _run(eng, fp2, op, a, b):
    import ctypes

    out = np.zeros_like(a)
    u = eng._lib.u64p
    rc = eng._lib.lib.b200msm_dbg_field_op(int(fp2), op, a.ctypes.data_as(u), b.ctypes.data_as(u), out.ctypes.data_as(u), a.shape[0])
    assert rc == 0, eng._lib.lib.b200msm_last_error()
    return out


@pytest.mark.parametrize("op", [0, 1, 2, 3, 4])
def test_fp_ops(eng, op):
    rng = random.Random(100 + op)
    n = 4096
    av, a = rand_fp_limbs(rng, n)
    bv, b = rand_fp_limbs(rng, n)
    rng.shuffle(bv)
    b = np.array([o.FpOps.to_limbs(v) for v in bv], dtype=np.uint64)
    got = _run(eng, False, op, a, b)
    F = o.FpOps
    ref = {0: lambda x, y: F.mul(x, y), 1: F.add, 2: F.sub, 3: lambda x, y: F.sqr(x), 4: lambda x, y: F.neg(x)}[op]
    exp = np.array([F.to_limbs(ref(x, y)) for x, y in zip(av, bv)], dtype=np.uint64)
    assert np.array_equal(got, exp), OPS[op]


def test_fp_inv(eng):
    rng = random.Random(7)
    av, a = rand_fp_limbs(rng, 256, edge=False)
    got = _run(eng, False, 5, a, a)
    exp = np.array([o.FpOps.to_limbs(o.FpOps.inv(x)) for x in av], dtype=np.uint64)
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("op", [0, 1, 2, 3, 4, 5])
def test_fp2_ops(eng, op):
    rng = random.Random(200 + op)
    n = 2048 if op != 5 else 128
    av, a = rand_fp2_limbs(rng, n, edge=(op != 5))
    bv, b = rand_fp2_limbs(rng, n)
    got = _run(eng, True, op, a, b)
    F = o.Fp2Ops
    ref = {0: F.mul, 1: F.add, 2: F.sub, 3: lambda x, y: F.sqr(x), 4: lambda x, y: F.neg(x), 5: lambda x, y: F.inv(x)}[op]
    exp = np.array([F.to_limbs(ref(x, y)) for x, y in zip(av, bv)], dtype=np.uint64)
    assert np.array_equal(got, exp), OPS[op]
