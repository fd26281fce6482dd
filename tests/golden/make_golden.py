"""Generates tests/golden/msm_vectors.json from the big-int oracle (oracle/bls12381.py, T0).

The reference ships no MSM vectors (its tests are unseeded-random, src/tests.rs:50-67) and cannot
be built here (no cargo), so these fixtures are OUR oracle's outputs, frozen: they pin the C port,
the CUDA path and future refactors of the oracle itself to one another.  Inputs are given as
byte-exact limb images in the reference's layouts, outputs as affine coordinates (hex integers,
non-Montgomery) — the normal form the reference's equality check reduces to.
Run:  python tests/golden/make_golden.py
"""
import json
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import bls12381 as o  # noqa: E402


def enc_pt(C, p):
    if p is None:
        return None
    if C is o.G2:
        return [[hex(p[0][0]), hex(p[0][1])], [hex(p[1][0]), hex(p[1][1])]]
    return [hex(p[0]), hex(p[1])]


def case(C, name, pts, scalars):
    return {
        "name": name,
        "group": C.name,
        "bases_limbs": [[hex(v) for v in C.affine_to_limbs(p)] for p in pts],
        "scalars_canonical_limbs": [[hex(v) for v in o.scalar_to_limbs(s, False)] for s in scalars],
        "scalars_montgomery_limbs": [[hex(v) for v in o.scalar_to_limbs(s, True)] for s in scalars],
        "result_affine": enc_pt(C, C.msm_naive(pts, scalars)),
    }


def main():
    out = {"constants": {
        "p": hex(o.P), "r": hex(o.R_ORDER), "fp_mont_one_limbs": [hex(v) for v in o.int_to_limbs(o.MONT_R, 6)],
        "fp_mont_r2_limbs": [hex(v) for v in o.int_to_limbs(o.MONT_R2, 6)],
        "fr_mont_one_limbs": [hex(v) for v in o.int_to_limbs(o.FR_MONT_R, 4)],
        "g1_2G": enc_pt(o.G1, o.G1.mul(o.G1_GEN, 2)),
        "g2_2G": enc_pt(o.G2, o.G2.mul(o.G2_GEN, 2)),
    }, "cases": []}
    for C, seed in ((o.G1, 1), (o.G2, 2)):
        rng = random.Random(seed)
        g = C.gen
        P = [C.mul(g, rng.randrange(1, o.R_ORDER)) for _ in range(12)]
        S = [rng.randrange(o.R_ORDER) for _ in range(12)]
        out["cases"].append(case(C, "reference_group_test_shape_n10", P[:10], S[:10]))
        out["cases"].append(case(C, "three_identity_bases_n13", P[:10] + [None] * 3, S[:10] + [5, 6, 7]))
        out["cases"].append(case(C, "single_generator_times_one", [g], [1]))
        out["cases"].append(case(C, "r_minus_one", [P[0]], [o.R_ORDER - 1]))
        out["cases"].append(case(C, "cancellation_to_identity", [P[0], P[0]], [9, o.R_ORDER - 9]))
        out["cases"].append(case(C, "duplicates_and_negations", [P[0], P[0], C.neg(P[0]), P[1], P[1]], [3, 3, 3, 1, 1]))
        out["cases"].append(case(C, "zero_scalars", P[:4], [0, 0, 0, 0]))
        out["cases"].append(case(C, "synthetic_stream_seed7_n16", o.synth_bases(C, 7, 16), o.synth_scalars(8, 16)))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "msm_vectors.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path, len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
