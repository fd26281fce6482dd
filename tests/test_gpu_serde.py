"""§8f-4: point (de)serialisation on the device vs the oracle, byte-exact / limb-exact.
Mirrors CanonicalSerialize / CanonicalDeserialize of G1Affine / G2Affine (src/g1.rs:358-431,
src/g2.rs:338-411); the reference pins this wire format cross-implementation at
src/tests.rs:70-96 (bytes must round-trip through ark_bls12_381)."""
import random

import numpy as np
import pytest

from helpers import curve, points_to_limbs
from oracle import bls12381 as o

pytestmark = pytest.mark.gpu

# published compressed encodings of the generators (zkcrypto / IETF BLS material)
G1_GEN_COMPRESSED = "97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb"
G2_GEN_COMPRESSED = ("93e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e"
                     "024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8")


def _grp(eng, g2):
    return eng.G2Projective if g2 else eng.G1Projective


@pytest.mark.parametrize("g2", [0, 1])
def test_generator_kat(eng, g2):
    C = curve(g2)
    got = _grp(eng, g2).serialize(points_to_limbs(C, [C.gen]), compressed=True)
    assert bytes(got[0]).hex() == (G2_GEN_COMPRESSED if g2 else G1_GEN_COMPRESSED)
    aff, st = _grp(eng, g2).deserialize(got, compressed=True, validate=True)
    assert st[0] == 0 and aff[0].tolist() == C.affine_to_limbs(C.gen)


@pytest.mark.parametrize("g2", [0, 1])
@pytest.mark.parametrize("compressed", [True, False])
def test_roundtrip_and_oracle_bytes(eng, g2, compressed):
    C = curve(g2)
    rng = random.Random(900 + 2 * g2 + compressed)
    pts = [C.mul(C.gen, rng.randrange(1, o.R_ORDER)) for _ in range(40)] + [None, C.neg(C.gen)]
    limbs = points_to_limbs(C, pts)
    got = _grp(eng, g2).serialize(limbs, compressed=compressed)
    exp = [o.serialize_point(C, p, compressed) for p in pts]
    assert [bytes(r) for r in got] == exp
    for validate in (True, False):
        aff, st = _grp(eng, g2).deserialize(got, compressed=compressed, validate=validate)
        assert not st.any()
        assert np.array_equal(aff, limbs)


@pytest.mark.parametrize("g2", [0, 1])
def test_rejections_match_oracle(eng, g2):
    """malformed flags, non-canonical x, x with no y, on-curve points outside the r-torsion,
    off-curve uncompressed points: status codes equal the oracle's"""
    C = curve(g2)
    F = C.F
    rng = random.Random(950 + g2)
    cb = 96 if g2 else 48
    good = o.serialize_point(C, C.mul(C.gen, 12345), True)
    cases = [good]
    cases.append(bytes([good[0] & 0x7F]) + good[1:])                      # compression flag missing
    cases.append(bytes([0xC0]) + b"\x00" * (cb - 2) + b"\x01")            # infinity with stray bits
    cases.append(bytes([0xE0]) + b"\x00" * (cb - 1))                      # infinity with sort flag
    cases.append(bytes([0x9F]) + b"\xff" * (cb - 1))                      # x ≥ p
    k = 0
    while len(cases) < 13:                                                # random x: no root / wrong subgroup / fine
        k += 1
        x = (rng.randrange(o.P), rng.randrange(o.P)) if g2 else rng.randrange(o.P)
        body = bytearray(o._coord_bytes(F, x))
        body[0] |= 0x80 | (0x20 if k % 2 else 0)
        cases.append(bytes(body))
    data = np.frombuffer(b"".join(cases), dtype=np.uint8).reshape(len(cases), cb)
    for validate in (True, False):
        aff, st = _grp(eng, g2).deserialize(data, compressed=True, validate=validate)
        for i, c in enumerate(cases):
            est, ept = o.deserialize_point(C, c, True, validate)
            assert st[i] == est, (i, validate)
            if est != 1:
                assert aff[i].tolist() == C.affine_to_limbs(ept), (i, validate)
    assert {0, 1, 2} <= set(int(o.deserialize_point(C, c, True, True)[0]) for c in cases)
    # uncompressed: off-curve point is status 2 under validation, accepted without it
    P_ = C.mul(C.gen, 777)
    bad = (P_[0], F.add(P_[1], F.one))
    raw = np.frombuffer(o.serialize_point(C, bad, False), dtype=np.uint8).reshape(1, 2 * cb)
    assert _grp(eng, g2).deserialize(raw, compressed=False, validate=True)[1][0] == 2
    assert _grp(eng, g2).deserialize(raw, compressed=False, validate=False)[1][0] == 0


def test_deserialize_then_msm(eng, cref):
    """proving-key shaped flow: compressed bytes → affine on device → msm"""
    n = 300
    bases = cref.synth_bases(0, 4242, n)
    sc = cref.synth_scalars(4343, n, True)
    blob = eng.G1Projective.serialize(bases, compressed=True)
    aff, st = eng.G1Projective.deserialize(blob, compressed=True, validate=True)
    assert not st.any() and np.array_equal(aff, bases)
    assert cref.affine_equal(0, eng.G1Projective.msm(aff, sc), cref.msm(0, bases, sc, 1))
