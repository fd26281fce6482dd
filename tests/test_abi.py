"""CPU suite: the C-ABI library loads, exports every symbol include/b200msm.h declares, and the
host-side mirror keeps the reference's error behaviour. No compute calls (no GPU here)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "b200msm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200msm_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(eng):
    names = _declared()
    assert len(names) >= 20
    lib = ctypes.CDLL(eng._lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(eng._lib.SIGNATURES), "ctypes table and header disagree"
    out = subprocess.check_output(["nm", "-D", "--defined-only", eng._lib.LIB_PATH], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert exported == set(names), "library exports exactly the header's symbols"


def test_header_compiles_as_c():
    src = '#include "b200msm.h"\nint main(void){ return B200MSM_OK; }\n'
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-x", "c", "-"],
                   input=src, text=True, check=True)


def test_version_and_sm100_sass(eng):
    assert b"sm_100a" in eng._lib.lib.b200msm_version()
    out = subprocess.run(["cuobjdump", "-lelf", eng._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_length_mismatch_is_err_min_len(eng):
    """arkworks' convention for VariableBaseMSM::msm; checked before any device work"""
    with pytest.raises(eng.MsmError) as ei:
        eng.G1Projective.msm(np.zeros((3, 12), dtype=np.uint64), np.zeros((2, 4), dtype=np.uint64))
    assert ei.value.value == 2
    with pytest.raises(eng.MsmError) as ei:
        eng.G2Projective.msm_bigint(np.zeros((1, 24), dtype=np.uint64), np.zeros((5, 4), dtype=np.uint64))
    assert ei.value.value == 1
    with pytest.raises(ValueError):
        eng.G1Projective.msm(np.zeros((3, 11), dtype=np.uint64), np.zeros((3, 4), dtype=np.uint64))


def test_no_device_fails_loudly(eng):
    """There is no CPU fallback: without a usable sm_100 device the call errs with Err(0)
    (reference GPU arm's convention, src/g1.rs:628-630) and a message. Skipped on a GPU box."""
    try:
        import torch

        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(eng.MsmError) as ei:
        eng.G1Projective.msm(np.zeros((2, 12), dtype=np.uint64), np.zeros((2, 4), dtype=np.uint64))
    assert ei.value.value == 0
    assert eng._lib.lib.b200msm_last_error()


def test_product_does_not_touch_the_oracle():
    """the oracle is test infrastructure: nothing under ark_blst_b200/ may reference it"""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ark_blst_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")) or f == "Makefile":
                t = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"\boracle\b|msm_ref|libmsm_ref", t) and f != "__init__.py":
                    for line in t.splitlines():
                        if re.search(r"(import|include|dlopen|CDLL).*(oracle|msm_ref)", line):
                            bad.append((f, line.strip()))
    assert not bad, bad


def test_layout_mirror_compiles():
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "test_layout_mirror.c")], check=True)


def test_cpp_mirror_compiles():
    """the C++ host mirror of VariableBaseMSM is header-only and must compile stand-alone"""
    src = '#include "ark_blst_b200/host/ark_blst_msm.hpp"\nint main(){ return ark_blst::G1Projective::NEGATION_IS_CHEAP ? 0 : 1; }\n'
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-I", ROOT, "-x", "c++", "-"], input=src, text=True, check=True)


def _plan(eng, group, n, glv):
    out = (ctypes.c_int * 4)()
    assert eng._lib.lib.b200msm_plan_query(group, n, glv, out) == 0
    return tuple(out)


def test_planner_host_logic(eng):
    """auto_plan needs no device. Without GLV and without batched-affine rounds it lands on the work model's canonical c* of SURVEY
    §8(d) at 2^16 / 2^20 / 2^24 (13 / 16 / 20; 16 instead of 18 at 2^22, whose top window would be
    degenerate) with c·W ≥ 256; GLV takes c = 16 with the unsigned top digit (8 + 1 windows) where
    it was measured to pay and is never picked automatically above 2^23 points."""
    from bench import work_model

    L = eng._lib.lib
    for g in (0, 1):
        assert L.b200msm_set_batch_affine(0) == 0      # the canonical widths are those of the plain XYZZ accumulation
        try:
            for logn, c_exp in ((16, 13), (20, 16), (22, 16), (24, 20)):
                c, W, glv, nb = _plan(eng, g, 1 << logn, 0)
                assert (c, glv) == (c_exp, 0) and c * W >= 256 and W == -(-256 // c) and nb == W << (c - 1)
                if logn != 22:
                    assert c == work_model(1 << logn, g)[0]
        finally:
            L.b200msm_set_batch_affine(-1)
        for logn in (16, 20, 22, 24):                   # with the batched-affine rounds the model may go one width down
            c, W, glv, nb = _plan(eng, g, 1 << logn, 0)
            assert glv == 0 and c * W >= 256 and W == -(-256 // c) and abs(c - work_model(1 << logn, g)[0]) <= 2
        for logn in (16, 18, 20):   # two 128-bit parts over (P, phi(P)) on G1; four 64-bit parts over the psi images on G2
            assert _plan(eng, g, 1 << logn, -1)[:3] in (((16, 5, 4), (16, 9, 2)) if g else ((16, 9, 2),))
            assert _plan(eng, g, 1 << logn, 1)[:3] == (16, 9, 2)
            assert _plan(eng, g, 1 << logn, 2)[:3] == ((16, 5, 4) if g else (16, 9, 2))
        for logn in (24, 26):
            assert _plan(eng, g, 1 << logn, -1)[2] == 0
        # forced GLV: a width dividing 128 has 128/c + 1 windows, any other keeps the carry window (c·W ≥ 129)
        c, W, glv, _ = _plan(eng, g, 1 << 12, 1)
        assert glv in (1, 2) and (W == 128 // c + 1 if 128 % c == 0 else c * W >= 129)
    out = (ctypes.c_int * 4)()
    assert eng._lib.lib.b200msm_plan_query(7, 1024, 0, out) != 0
    assert eng._lib.lib.b200msm_plan_query(0, 0, 0, out) != 0


def test_table_plan_host_logic(eng):
    """fixed-base table widths: only those whose top window keeps ≥ c − 5 bits (10, 13, 16, 20), wide
    enough to fill the GPU from 2^18 points (G1), and windows × points below 2^31"""
    for g in (0, 1):
        for logn in range(8, 27):
            c, W = eng.table_plan(g, 1 << logn)
            assert c in (10, 13, 16, 20) and W == -(-256 // c)
            assert 255 - (W - 1) * c >= c - 5
        assert eng.table_plan(g, 1 << 20) == (20, 13)
    assert eng.table_plan(0, 1 << 18)[0] == 20
    with pytest.raises(eng._lib.B200MsmError):
        eng.table_plan(0, 1 << 28)
    with pytest.raises(eng._lib.B200MsmError):
        eng.table_plan(0, 1 << 20, 24)
    assert eng.table_plan(0, 1000, 11) == (11, 24)   # an explicit width is taken as given


def test_planner_cpp_unit(tmp_path):
    """tests/cpp/test_plan.cpp: the planner header compiled with plain g++ (no CUDA, no device)"""
    exe = str(tmp_path / "test_plan")
    subprocess.run(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_plan.cpp")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


def test_bench_counts_the_nonzero_digits_of_witness_like_scalars(cref):
    """bench.py's Groth16 witness-like numerator counts the bucket additions actually needed (VERDICT r1 weak #9:
    the uniform-scalar model printed frac 1.19 on scalars that are 70 % zeros and ones)"""
    import bench

    s = bench.witness_scalars(123, 3000)
    for c in (13, 18):
        W = -(-256 // c)
        want = sum(1 for i in range(s.shape[0]) for w in range(W)
                   if cref.lib().ref_booth_digit(cref._p(np.ascontiguousarray(s[i])), w, c) != 0)
        assert bench.count_nonzero_booth_digits(s, c) == want
    uni = cref.synth_scalars(5, 3000, False)
    c, W, _, _ = bench.work_model(1 << 22, False)
    assert bench.count_nonzero_booth_digits(s, c) < 0.45 * bench.count_nonzero_booth_digits(uni, c)
    assert bench.work_model_counted(bench.count_nonzero_booth_digits(s, c), 1 << 22, False) < bench.work_model(1 << 22, False)[2]


def test_divsteps_inversion_on_the_host(tmp_path):
    """csrc/modinv.cuh is plain integer code: compiled with g++ here and checked against Python's pow(x, −1, p)
    (the device runs the same source; tests/test_gpu_batch_affine.py checks it there)"""
    import random

    from oracle import bls12381 as o

    src = tmp_path / "w.cpp"
    src.write_text('#include "%s"\nextern "C" void mi_inv(const uint32_t *a, uint32_t *out, int n) {'
                   ' for (int i = 0; i < n; i++) b200msm::mi_inverse_u32(out + 12 * i, a + 12 * i); }\n'
                   % os.path.join(ROOT, "ark_blst_b200", "csrc", "modinv.cuh"))
    so = str(tmp_path / "libmi.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", so, str(src)], check=True)
    L = ctypes.CDLL(so)
    rng = random.Random(1)
    p = o.P
    vals = [0, 1, 2, p - 1, p - 2, (p - 1) // 2, 1 << 380, 3, 1 << 30, (1 << 30) - 1, 1 << 60] + [rng.randrange(p) for _ in range(5000)]
    a = np.array([[(v >> (32 * i)) & 0xFFFFFFFF for i in range(12)] for v in vals], dtype=np.uint32)
    out = np.zeros_like(a)
    L.mi_inv(a.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), len(vals))
    for v, row in zip(vals, out):
        assert sum(int(x) << (32 * i) for i, x in enumerate(row)) == (pow(v, -1, p) if v else 0)
