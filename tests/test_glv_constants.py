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


# ---- round 2: the four-part decomposition on G2 (csrc/gls4.cuh, csrc/accumulate.cuh k_psi_tables) ----
def _psi(Q):
    F, P = o.Fp2Ops, o.P
    conj = lambda a: (a[0], (-a[1]) % P)
    gx = F.inv(o.fp2_pow((1, 1), (P - 1) // 3))
    gy = F.inv(o.fp2_pow((1, 1), (P - 1) // 2))
    return (F.mul(conj(Q[0]), gx), F.mul(conj(Q[1]), gy)), gx, gy


def test_psi_constants_and_eigenvalue():
    """γx = (1+u)^-(p-1)/3 = (0, γx1), γy = (1+u)^-(p-1)/2 as the CUDA source holds them (Montgomery limbs), and
    ψ(Q) = [z]·Q on the order-r subgroup of the twist — what makes |z|·Q = −ψ(Q)"""
    G = o.G2.mul(o.G2.gen, 0xC0FFEE)
    q1, gx, gy = _psi(G)
    mont = lambda name: _array("accumulate.cuh", name) * o.MONT_RINV % o.P
    assert gx[0] == 0 and mont("PSI_GX_C1") == gx[1]
    assert (mont("PSI_GY_C0"), mont("PSI_GY_C1")) == gy
    assert o.G2.eq(q1, o.G2.mul(G, o.BLS_X % o.R_ORDER))
    assert (o.P - o.BLS_X) % o.R_ORDER == 0                       # p ≡ z (mod r): ψ is the p-power Frobenius carried to the twist
    q2, _, _ = _psi(q1)
    beta = _array("accumulate.cuh", "GLV_BETA") * o.MONT_RINV % o.P
    assert q2 == ((G[0][0] * beta % o.P, G[0][1] * beta % o.P), o.Fp2Ops.neg(G[1]))   # ψ²(x, y) = (β·x, −y)
    text = open(os.path.join(CSRC, "gls4.cuh")).read()
    assert int(re.search(r"GLS4_Z = (0x[0-9a-f]+)ull", text).group(1), 16) == -o.BLS_X


def test_four_part_decomposition_on_the_host(tmp_path):
    """csrc/gls4.cuh is plain integer code: compiled with g++ here, its base-|z| digits checked against Python's
    divmod, and Σ k_i·image_i(Q) = k·Q against the big-int oracle"""
    import ctypes
    import subprocess

    import numpy as np

    src = tmp_path / "g.cpp"
    src.write_text('#include "%s"\nextern "C" void gls4(const uint32_t *k, uint32_t *out, int n) { for (int i = 0; i < n; i++) {'
                   ' uint32_t p[4][8]; b200msm::gls4_decompose(k + 8 * i, p); for (int d = 0; d < 4; d++) {'
                   ' out[8 * i + 2 * d] = p[d][0]; out[8 * i + 2 * d + 1] = p[d][1];'
                   ' for (int j = 2; j < 8; j++) if (p[d][j]) out[8 * i] = 0xdeadbeef; } } }\n' % os.path.join(CSRC, "gls4.cuh"))
    so = str(tmp_path / "libg.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", so, str(src)], check=True)
    L = ctypes.CDLL(so)
    Z, r = -o.BLS_X, o.R_ORDER
    rng = random.Random(5)
    vals = [0, 1, Z - 1, Z, Z + 1, Z * Z - 1, Z * Z, Z ** 3 - 1, Z ** 3, r - 1, (1 << 64) - 1, 1 << 64, (1 << 192) + 5] + [rng.randrange(r) for _ in range(20000)]
    a = np.array([[(v >> (32 * i)) & 0xFFFFFFFF for i in range(8)] for v in vals], dtype=np.uint32)
    out = np.zeros_like(a)
    L.gls4(a.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), len(vals))
    assert Z ** 4 > r
    for v, row in zip(vals, out):
        ds = [int(row[2 * d]) | (int(row[2 * d + 1]) << 32) for d in range(4)]
        assert sum(d * Z ** i for i, d in enumerate(ds)) == v and all(d < Z for d in ds)
        top = [(d >> 48) + ((d >> 47) & 1) for d in ds]         # unsigned top digit + Booth carry (c = 16): fits two windows' buckets
        assert max(top) <= 1 << 16
    # the identity the engine relies on: k·Q = k0·Q + k1·(−ψQ) + k2·ψ²Q + k3·(−ψ³Q)
    Q = o.G2.mul(o.G2.gen, 987654321)
    q1, _, _ = _psi(Q)
    q2, _, _ = _psi(q1)
    q3, _, _ = _psi(q2)
    imgs = [Q, o.G2.neg(q1), q2, o.G2.neg(q3)]
    k = vals[-1]
    acc = None
    for i, im in enumerate(imgs):
        acc = o.G2.add_affine(acc, o.G2.mul(im, (k // Z ** i) % Z))
    assert o.G2.eq(acc, o.G2.mul(Q, k))
