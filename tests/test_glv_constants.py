"""CPU suite: the constants the device-side GLV path is built on (csrc/scalar.cuh,
csrc/accumulate.cuh), read out of the CUDA sources and checked against the big-int oracle — the
same way tests/test_oracle.py pins the field constants to the reference's Rust text.
  k = k1 + k2·λ (mod r), λ = z² − 1, φ(x, y) = (β·x, y) = λ·P on G1 and (β²·x, y) = λ·Q on G2,
  Barrett constant μ = ⌊2^256 / λ⌋, and the decomposition's bounds (both halves < 2^128)."""
import os
import random
import re

from oracle import bls12381 as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "ark_blst_b200", "csrc")


def _array(fname, name):
    text = open(os.path.join(CSRC, fname)).read()
    m = re.search(name + r"\[\d+\]\s*=\s*\{([^}]*)\}", text)
    assert m, name
    limbs = [int(v, 16) for v in re.findall(r"0x[0-9a-fA-F]+", m.group(1))]
    return sum(v << (32 * i) for i, v in enumerate(limbs))


def test_lambda_and_barrett_constant():
    lam = _array("scalar.cuh", "GLV_LAMBDA")
    mu = _array("scalar.cuh", "GLV_MU")
    assert lam == o.BLS_X * o.BLS_X - 1
    assert (lam * lam + lam + 1) % o.R_ORDER == 0          # a primitive cube root of unity mod r
    assert mu == (1 << 256) // lam
    assert lam.bit_length() == 128 and (o.R_ORDER // lam) < (1 << 128)


def test_decomposition_identity_and_bounds():
    """what glv_decompose computes: k2 = ⌊k/λ⌋, k1 = k mod λ — both below 2^128, top digits below
    λ's, so an 8×16-bit unsigned-top recoding never needs a ninth carry window"""
    lam = o.BLS_X * o.BLS_X - 1
    rng = random.Random(9)
    for k in [0, 1, lam - 1, lam, lam + 1, o.R_ORDER - 1, lam * lam % o.R_ORDER] + [rng.randrange(o.R_ORDER) for _ in range(2000)]:
        k2, k1 = divmod(k, lam)
        assert (k1 + k2 * lam) % o.R_ORDER == k
        assert k1 < lam < (1 << 128) and k2 <= lam + 1
        for h in (k1, k2):
            top = (h >> 112) + ((h >> 111) & 1)             # unsigned top digit + Booth carry from below
            assert top <= 1 << 16
        # Barrett estimate with μ is at most 2 below the true quotient (the device corrects ≤ 2 times)
        q = (k * ((1 << 256) // lam)) >> 256
        assert 0 <= k2 - q <= 2


def test_endomorphism_constants_act_as_lambda():
    beta = _array("accumulate.cuh", "GLV_BETA") * o.MONT_RINV % o.P
    beta_sq = _array("accumulate.cuh", "GLV_BETA_SQ") * o.MONT_RINV % o.P
    lam = o.BLS_X * o.BLS_X - 1
    assert beta != 1 and pow(beta, 3, o.P) == 1 and beta_sq == beta * beta % o.P
    rng = random.Random(11)
    for _ in range(3):
        k = rng.randrange(1, o.R_ORDER)
        P = o.G1.mul(o.G1.gen, k)
        assert o.G1.eq((P[0] * beta % o.P, P[1]), o.G1.mul(P, lam))
        Q = o.G2.mul(o.G2.gen, k)
        x = (Q[0][0] * beta_sq % o.P, Q[0][1] * beta_sq % o.P)
        assert o.G2.eq((x, Q[1]), o.G2.mul(Q, lam))
