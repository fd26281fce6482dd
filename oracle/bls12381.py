"""T0 oracle: big-int BLS12-381 G1/G2 arithmetic and the naive MSM  Σ sᵢ·Pᵢ.

TEST INFRASTRUCTURE ONLY.  Nothing under ark_blst_b200/ may import this file; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may.

PARITY UNPINNED by reference vectors: the reference (nikkolasg/ark-blst) ships no MSM golden
vectors (its tests are unseeded-random property tests) and the arithmetic lives in un-vendored
crates (blst =0.3.10, blstrs ^0.6.1 fork, Cargo.toml:16,22,59) that cannot be built here (no
cargo/rustc).  This oracle therefore follows
  * the reference's *definition* of the MSM result:  fold(acc + b.mul(s))   src/tests.rs:58-61
  * the reference's constants, each asserted below:
        Fp modulus limbs                 src/fp.rs:25-32
        (p-1)/2 limbs (the only KAT)     src/fp.rs:714-721
        Fr modulus limbs                 src/scalar.rs:476-481
        G1 cofactor                      src/g1.rs:42
        G2 cofactor                      src/g2.rs:45-54
  * the reference's data layouts: Fp = 6×u64 LE Montgomery (R=2^384)  src/fp.rs:482-491,532;
    Fp2 = (c0,c1) src/fp2.rs:450-454; Scalar = 4×u64 LE Montgomery (R=2^256) src/scalar.rs:23-25;
    G1Affine/G1Projective are repr(transparent) over blst_p1_affine / blst_p1
    src/g1.rs:54-56,435-437 (x,y[,z] each an Fp; infinity affine = all-zero, Jacobian z = 0).
The MSM value is a unique group element, so any correct evaluation equals blst's after
normalisation to affine; that is the equality every parity test uses.
"""
from __future__ import annotations

# ---------------------------------------------------------------------------------------------
# constants (asserted against the reference's literals in tests/test_oracle_constants.py)
# ---------------------------------------------------------------------------------------------
P_LIMBS = [  # src/fp.rs:25-32
    0xB9FE_FFFF_FFFF_AAAB, 0x1EAB_FFFE_B153_FFFF, 0x6730_D2A0_F6B0_F624,
    0x6477_4B84_F385_12BF, 0x4B1B_A7B6_434B_ACD7, 0x1A01_11EA_397F_E69A,
]
R_LIMBS = [  # src/scalar.rs:476-481
    0xFFFF_FFFF_0000_0001, 0x53BD_A402_FFFE_5BFE, 0x3339_D808_09A1_D805, 0x73ED_A753_299D_7D48,
]


def limbs_to_int(limbs, bits=64):
    v = 0
    for i, l in enumerate(limbs):
        v |= int(l) << (bits * i)
    return v


def int_to_limbs(v, n, bits=64):
    mask = (1 << bits) - 1
    return [(v >> (bits * i)) & mask for i in range(n)]


P = limbs_to_int(P_LIMBS)
R_ORDER = limbs_to_int(R_LIMBS)
BLS_X = -0xD201000000010000  # curve parameter; r = x^4 - x^2 + 1, h1 = (x-1)^2/3
MONT_R = (1 << 384) % P       # Fp Montgomery one  (blstrs::fp::R, src/fp.rs:532)
MONT_R2 = (MONT_R * MONT_R) % P
MONT_RINV = pow(MONT_R, -1, P)
FR_MONT_R = (1 << 256) % R_ORDER
FR_MONT_RINV = pow(FR_MONT_R, -1, R_ORDER)
P_INV64 = (-pow(P, -1, 1 << 64)) % (1 << 64)   # 0x89f3fffcfffcfffd
R_INV64 = (-pow(R_ORDER, -1, 1 << 64)) % (1 << 64)

G1_COFACTOR = limbs_to_int([0x8C00AAAB0000AAAB, 0x396C8C005555E156])  # src/g1.rs:42
G2_COFACTOR = limbs_to_int([  # src/g2.rs:45-54
    0xCF1C38E31C7238E5, 0x1616EC6E786F0C70, 0x21537E293A6691AE, 0xA628F1CB4D9E82EF,
    0xA68A205B2E5A7DDF, 0xCD91DE4547085ABA, 0x91D50792876A202, 0x5D543A95414E7F1,
])

# standard generators (zkcrypto/IETF); validated on-curve and of order r in the tests
G1_GEN = (
    0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
    0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1,
)
G2_GEN = (
    (0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
     0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E),
    (0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
     0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE),
)


# ---------------------------------------------------------------------------------------------
# field towers: a tiny "field ops" record so G1 (Fp) and G2 (Fp2) share the curve code
# ---------------------------------------------------------------------------------------------
class FpOps:
    name = "fp"
    zero = 0
    one = 1
    b = 4  # y^2 = x^3 + 4
    nlimbs64 = 6

    @staticmethod
    def add(a, b): return (a + b) % P
    @staticmethod
    def sub(a, b): return (a - b) % P
    @staticmethod
    def mul(a, b): return (a * b) % P
    @staticmethod
    def sqr(a): return (a * a) % P
    @staticmethod
    def neg(a): return (-a) % P
    @staticmethod
    def inv(a): return pow(a, -1, P)
    @staticmethod
    def is_zero(a): return a % P == 0
    @staticmethod
    def eq(a, b): return (a - b) % P == 0
    @staticmethod
    def to_limbs(a):  # Montgomery, 6×u64 LE
        return int_to_limbs((a * MONT_R) % P, 6)
    @staticmethod
    def from_limbs(l):
        return (limbs_to_int(l[:6]) * MONT_RINV) % P


class Fp2Ops:
    name = "fp2"
    zero = (0, 0)
    one = (1, 0)
    b = (4, 4)  # y^2 = x^3 + 4(1+u), u^2 = -1
    nlimbs64 = 12

    @staticmethod
    def add(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)
    @staticmethod
    def sub(a, b): return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)
    @staticmethod
    def mul(a, b):
        return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
    @staticmethod
    def sqr(a):
        return (((a[0] + a[1]) * (a[0] - a[1])) % P, (2 * a[0] * a[1]) % P)
    @staticmethod
    def neg(a): return ((-a[0]) % P, (-a[1]) % P)
    @staticmethod
    def inv(a):
        d = pow(a[0] * a[0] + a[1] * a[1], -1, P)
        return ((a[0] * d) % P, (-a[1] * d) % P)
    @staticmethod
    def is_zero(a): return a[0] % P == 0 and a[1] % P == 0
    @staticmethod
    def eq(a, b): return (a[0] - b[0]) % P == 0 and (a[1] - b[1]) % P == 0
    @staticmethod
    def to_limbs(a):  # c0 then c1 (src/fp2.rs:450-454), each Montgomery 6×u64 LE
        return FpOps.to_limbs(a[0]) + FpOps.to_limbs(a[1])
    @staticmethod
    def from_limbs(l):
        return (FpOps.from_limbs(l[0:6]), FpOps.from_limbs(l[6:12]))


# ---------------------------------------------------------------------------------------------
# curve arithmetic. Points: None = infinity, else affine (x, y); Jacobian (X, Y, Z) internally.
# ---------------------------------------------------------------------------------------------
class Curve:
    def __init__(self, F, gen, name):
        self.F, self.gen, self.name = F, gen, name

    # -- affine ground truth (slow, obviously correct) --
    def is_on_curve(self, pt):
        if pt is None:
            return True
        F = self.F
        x, y = pt
        return F.eq(F.sqr(y), F.add(F.mul(F.sqr(x), x), F.b))

    def neg(self, pt):
        return None if pt is None else (pt[0], self.F.neg(pt[1]))

    def add_affine(self, p, q):
        F = self.F
        if p is None:
            return q
        if q is None:
            return p
        if F.eq(p[0], q[0]):
            if F.eq(p[1], q[1]) and not F.is_zero(p[1]):
                lam = F.mul(F.mul(F.sqr(p[0]), _small(F, 3)), F.inv(F.add(p[1], p[1])))
            else:
                return None
        else:
            lam = F.mul(F.sub(q[1], p[1]), F.inv(F.sub(q[0], p[0])))
        x3 = F.sub(F.sub(F.sqr(lam), p[0]), q[0])
        y3 = F.sub(F.mul(lam, F.sub(p[0], x3)), p[1])
        return (x3, y3)

    # -- Jacobian (fast path of the oracle) --
    def to_jac(self, pt):
        F = self.F
        return (F.one, F.one, F.zero) if pt is None else (pt[0], pt[1], F.one)

    def from_jac(self, j):
        F = self.F
        X, Y, Z = j
        if F.is_zero(Z):
            return None
        zi = F.inv(Z)
        zi2 = F.sqr(zi)
        return (F.mul(X, zi2), F.mul(Y, F.mul(zi2, zi)))

    def jac_dbl(self, j):
        F = self.F
        X, Y, Z = j
        if F.is_zero(Z) or F.is_zero(Y):
            return (F.one, F.one, F.zero)
        A = F.sqr(X); B = F.sqr(Y); C = F.sqr(B)
        t = F.sub(F.sub(F.sqr(F.add(X, B)), A), C)
        D = F.add(t, t)
        E = F.add(F.add(A, A), A)
        Fq = F.sqr(E)
        X3 = F.sub(Fq, F.add(D, D))
        C8 = F.add(C, C); C8 = F.add(C8, C8); C8 = F.add(C8, C8)
        Y3 = F.sub(F.mul(E, F.sub(D, X3)), C8)
        Z3 = F.mul(F.add(Y, Y), Z)
        return (X3, Y3, Z3)

    def jac_add(self, p, q):
        F = self.F
        if F.is_zero(p[2]):
            return q
        if F.is_zero(q[2]):
            return p
        Z1Z1 = F.sqr(p[2]); Z2Z2 = F.sqr(q[2])
        U1 = F.mul(p[0], Z2Z2); U2 = F.mul(q[0], Z1Z1)
        S1 = F.mul(F.mul(p[1], q[2]), Z2Z2); S2 = F.mul(F.mul(q[1], p[2]), Z1Z1)
        if F.eq(U1, U2):
            if F.eq(S1, S2):
                return self.jac_dbl(p)
            return (F.one, F.one, F.zero)
        H = F.sub(U2, U1); Rr = F.sub(S2, S1)
        HH = F.sqr(H); HHH = F.mul(H, HH); V = F.mul(U1, HH)
        X3 = F.sub(F.sub(F.sqr(Rr), HHH), F.add(V, V))
        Y3 = F.sub(F.mul(Rr, F.sub(V, X3)), F.mul(S1, HHH))
        Z3 = F.mul(F.mul(p[2], q[2]), H)
        return (X3, Y3, Z3)

    def mul(self, pt, k):
        """k·pt by plain double-and-add (the shape of mul_bigint, src/g1.rs:331-341)."""
        F = self.F
        if k < 0:
            return self.mul(self.neg(pt), -k)
        acc = (F.one, F.one, F.zero)
        if pt is None or k == 0:
            return None
        base = self.to_jac(pt)
        for bit in bin(k)[2:]:
            acc = self.jac_dbl(acc)
            if bit == "1":
                acc = self.jac_add(acc, base)
        return self.from_jac(acc)

    def msm_naive(self, bases, scalars):
        """The reference's definition: fold(acc + b.mul(s))  (src/tests.rs:58-61)."""
        F = self.F
        acc = (F.one, F.one, F.zero)
        for b, s in zip(bases, scalars):
            acc = self.jac_add(acc, self.to_jac(self.mul(b, s % R_ORDER)))
        return self.from_jac(acc)

    def eq(self, p, q):
        if p is None or q is None:
            return p is None and q is None
        return self.F.eq(p[0], q[0]) and self.F.eq(p[1], q[1])

    # -- blst byte layouts --
    def affine_to_limbs(self, pt):
        """blst_p{1,2}_affine: x then y, Montgomery limbs; infinity = all zero."""
        F = self.F
        if pt is None:
            return [0] * (2 * F.nlimbs64)
        return F.to_limbs(pt[0]) + F.to_limbs(pt[1])

    def affine_from_limbs(self, l):
        F = self.F
        n = F.nlimbs64
        if all(int(v) == 0 for v in l[: 2 * n]):
            return None
        return (F.from_limbs(l[0:n]), F.from_limbs(l[n : 2 * n]))

    def jac_from_limbs(self, l):
        """blst_p{1,2}: X, Y, Z Montgomery limbs; infinity iff Z == 0. Returns affine/None."""
        F = self.F
        n = F.nlimbs64
        j = (F.from_limbs(l[0:n]), F.from_limbs(l[n : 2 * n]), F.from_limbs(l[2 * n : 3 * n]))
        return self.from_jac(j)

    def jac_to_limbs(self, pt):
        F = self.F
        if pt is None:
            return [0] * (3 * F.nlimbs64)
        return F.to_limbs(pt[0]) + F.to_limbs(pt[1]) + F.to_limbs(F.one)


def _small(F, k):
    v = F.zero
    for _ in range(k):
        v = F.add(v, F.one)
    return v


G1 = Curve(FpOps, G1_GEN, "g1")
G2 = Curve(Fp2Ops, G2_GEN, "g2")


# ---------------------------------------------------------------------------------------------
# scalars (Fr): layouts of the two entry points
# ---------------------------------------------------------------------------------------------
def scalar_to_limbs(s, montgomery):
    """`msm` receives &[Scalar] = Montgomery 4×u64 (src/scalar.rs:23-25); `msm_bigint` receives
    canonical BigInt<4> (src/scalar.rs:458-463)."""
    s %= R_ORDER
    if montgomery:
        s = (s * FR_MONT_R) % R_ORDER
    return int_to_limbs(s, 4)


def scalar_from_limbs(l, montgomery):
    v = limbs_to_int(l[:4])
    if montgomery:
        v = (v * FR_MONT_RINV) % R_ORDER
    return v


# ---------------------------------------------------------------------------------------------
# deterministic inputs shared by oracle / C port / CUDA generator: counter-based splitmix64
# ---------------------------------------------------------------------------------------------
M64 = (1 << 64) - 1


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & M64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def synth_scalar(seed, i):
    """255-bit draw from four splitmix64 words of counter 4i..4i+3, minus r once if ≥ r.
    (Every implementation — this file, oracle/msm_ref.c, csrc/synth.cu — must agree bit for bit.)"""
    v = 0
    for j in range(4):
        v |= splitmix64((seed + 4 * i + j) & M64) << (64 * j)
    v &= (1 << 255) - 1
    if v >= R_ORDER:
        v -= R_ORDER
    return v


def synth_dlog(seed, i):
    """discrete log kᵢ of synthetic base i (Pᵢ = kᵢ·G); never zero."""
    k = synth_scalar(seed ^ 0x5EED_BA5E_5EED_BA5E, i)
    return k if k != 0 else 1


def synth_bases(curve, seed, n):
    return [curve.mul(curve.gen, synth_dlog(seed, i)) for i in range(n)]


def synth_scalars(seed, n):
    return [synth_scalar(seed, i) for i in range(n)]


def msm_by_dlog(curve, seed_bases, seed_scalars, n, scalars=None):
    """T2 oracle: Σ sᵢ·(kᵢ·G) = (Σ sᵢkᵢ mod r)·G — O(n) Fr work, no EC oracle needed."""
    acc = 0
    for i in range(n):
        s = synth_scalar(seed_scalars, i) if scalars is None else scalars[i]
        acc = (acc + s * synth_dlog(seed_bases, i)) % R_ORDER
    return curve.mul(curve.gen, acc)


# ---------------------------------------------------------------------------------------------
# ZCash / IETF point encodings. Reference: CanonicalSerialize/Deserialize for
# G1Affine / G2Affine, src/g1.rs:358-431, src/g2.rs:338-411 → blstrs to_compressed /
# to_uncompressed / from_*_unchecked, then Valid::check (on curve ∧ torsion free, src/g1.rs:386-396).
# Big-endian coordinates; G2 coordinates are written c1 then c0; three flag bits in byte 0:
#   0x80 compressed, 0x40 infinity (everything else zero), 0x20 y is the lexicographically
#   larger of (y, −y) (compressed form only).
# ---------------------------------------------------------------------------------------------
def fp_sqrt(a):
    """p ≡ 3 (mod 4): a^((p+1)/4), or None if a is a non-residue"""
    r = pow(a, (P + 1) // 4, P)
    return r if r * r % P == a % P else None


def fp2_pow(a, e):
    F = Fp2Ops
    r = F.one
    for bit in bin(e)[2:]:
        r = F.sqr(r)
        if bit == "1":
            r = F.mul(r, a)
    return r


def fp2_sqrt(a):
    """Adj / Rodríguez-Henríquez Algorithm 9 for p ≡ 3 (mod 4); None if no root"""
    F = Fp2Ops
    if F.is_zero(a):
        return F.zero
    a1 = fp2_pow(a, (P - 3) // 4)
    alpha = F.mul(F.sqr(a1), a)
    a0 = F.mul((alpha[0], (-alpha[1]) % P), alpha)  # α^p · α (Frobenius = conjugation)
    if F.eq(a0, (P - 1, 0)):
        return None
    x0 = F.mul(a1, a)
    if F.eq(alpha, (P - 1, 0)):
        x = ((-x0[1]) % P, x0[0])  # u·x0
    else:
        b = fp2_pow(F.add(F.one, alpha), (P - 1) // 2)
        x = F.mul(b, x0)
    return x if F.eq(F.sqr(x), a) else None


def _lex_largest(F, y):
    half = (P - 1) // 2
    if F is Fp2Ops:
        return y[1] > half or (y[1] == 0 and y[0] > half)
    return y > half


def _coord_bytes(F, v):
    return (v[1].to_bytes(48, "big") + v[0].to_bytes(48, "big")) if F is Fp2Ops else v.to_bytes(48, "big")


def _coord_from(F, b):
    if F is Fp2Ops:
        c1, c0 = int.from_bytes(b[:48], "big"), int.from_bytes(b[48:96], "big")
        return None if c1 >= P or c0 >= P else (c0, c1)
    v = int.from_bytes(b[:48], "big")
    return None if v >= P else v


def serialize_point(curve, pt, compressed):
    F = curve.F
    cb = 96 if F is Fp2Ops else 48
    if pt is None:
        out = bytearray(cb if compressed else 2 * cb)
        out[0] = 0xC0 if compressed else 0x40
        return bytes(out)
    out = bytearray(_coord_bytes(F, pt[0]) + (b"" if compressed else _coord_bytes(F, pt[1])))
    if compressed:
        out[0] |= 0x80 | (0x20 if _lex_largest(F, pt[1]) else 0)
    return bytes(out)


def deserialize_point(curve, data, compressed, validate):
    """→ (status, point): status 0 ok, 1 malformed (what blstrs from_*_unchecked rejects; the
    reference unwrap()s it), 2 fails Valid::check (not on curve / not in the r-torsion)."""
    F = curve.F
    cb = 96 if F is Fp2Ops else 48
    data = bytes(data)
    flags = data[0] & 0xE0
    body = bytes([data[0] & 0x1F]) + data[1:]
    if bool(flags & 0x80) != bool(compressed):
        return 1, None
    if flags & 0x40:
        ok = not any(body) and not (flags & 0x20)
        return (0, None) if ok else (1, None)
    if not compressed and (flags & 0x20):
        return 1, None
    x = _coord_from(F, body[:cb])
    if x is None:
        return 1, None
    if compressed:
        rhs = F.add(F.mul(F.sqr(x), x), F.b)
        y = fp2_sqrt(rhs) if F is Fp2Ops else fp_sqrt(rhs)
        if y is None:
            return 1, None
        if _lex_largest(F, y) != bool(flags & 0x20):
            y = F.neg(y)
        pt = (x, y)
    else:
        y = _coord_from(F, body[cb : 2 * cb])
        if y is None:
            return 1, None
        pt = (x, y)
    if validate:
        if not curve.is_on_curve(pt) or curve.mul(pt, R_ORDER) is not None:
            return 2, pt
    return 0, pt
