"""ctypes loader for the T1 CPU oracle oracle/libmsm_ref.so (built by oracle/Makefile).

TEST INFRASTRUCTURE ONLY — importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never from ark_blst_b200/.
Arrays are numpy uint64, C-contiguous, in the reference's layouts (see oracle/bls12381.py).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmsm_ref.so")
_lib = None


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("msm_ref.c", "field.h", "ec_tmpl.h")]
    if (
        force
        or not os.path.exists(_SO)
        or any(os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    ):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libmsm_ref.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        u64p = ctypes.POINTER(ctypes.c_uint64)
        for name in ("ref_g1_msm", "ref_g2_msm"):
            f = getattr(_lib, name)
            f.argtypes = [u64p, u64p, ctypes.c_size_t, ctypes.c_int, u64p, ctypes.c_int, ctypes.c_int]
            f.restype = ctypes.c_int
        for name in ("ref_g1_msm_naive", "ref_g2_msm_naive"):
            f = getattr(_lib, name)
            f.argtypes = [u64p, u64p, ctypes.c_size_t, ctypes.c_int, u64p]
            f.restype = None
        _lib.ref_synth_scalars.argtypes = [ctypes.c_uint64, ctypes.c_size_t, ctypes.c_int, u64p]
        _lib.ref_synth_dlogs.argtypes = [ctypes.c_uint64, ctypes.c_size_t, u64p]
        _lib.ref_synth_bases.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_size_t, u64p, u64p, ctypes.c_int]
        _lib.ref_fr_dot_synth.argtypes = [ctypes.c_uint64, u64p, ctypes.c_size_t, u64p]
        _lib.ref_booth_digit.argtypes = [u64p, ctypes.c_uint, ctypes.c_uint]
        _lib.ref_booth_digit.restype = ctypes.c_int
        _lib.ref_window_rule.argtypes = [ctypes.c_size_t]
        _lib.ref_window_rule.restype = ctypes.c_uint
    return _lib


def _p(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))


def ncores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def msm(g2, bases, scalars, mont, nthreads=None, window=0):
    """Pippenger. bases (n, 12|24) u64, scalars (n, 4) u64 → Jacobian (18|36,) u64."""
    n = scalars.shape[0]
    out = np.zeros(36 if g2 else 18, dtype=np.uint64)
    f = lib().ref_g2_msm if g2 else lib().ref_g1_msm
    rc = f(_p(bases), _p(scalars), n, int(mont), _p(out), nthreads or ncores(), window)
    assert rc == 0
    return out


def msm_naive(g2, bases, scalars, mont):
    n = scalars.shape[0]
    out = np.zeros(36 if g2 else 18, dtype=np.uint64)
    (lib().ref_g2_msm_naive if g2 else lib().ref_g1_msm_naive)(_p(bases), _p(scalars), n, int(mont), _p(out))
    return out


def to_affine(g2, jac):
    out = np.zeros(24 if g2 else 12, dtype=np.uint64)
    jac = np.ascontiguousarray(jac, dtype=np.uint64)
    (lib().ref_g2_to_affine if g2 else lib().ref_g1_to_affine)(_p(jac), _p(out))
    return out


def add(g2, a, b):
    out = np.zeros(36 if g2 else 18, dtype=np.uint64)
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = np.ascontiguousarray(b, dtype=np.uint64)
    (lib().ref_g2_add if g2 else lib().ref_g1_add)(_p(a), _p(b), _p(out))
    return out


def mul(g2, aff, k):
    out = np.zeros(36 if g2 else 18, dtype=np.uint64)
    aff = np.ascontiguousarray(aff, dtype=np.uint64)
    k = np.ascontiguousarray(k, dtype=np.uint64)
    (lib().ref_g2_mul if g2 else lib().ref_g1_mul)(_p(aff), _p(k), _p(out))
    return out


def generator_limbs(g2):
    from . import bls12381 as o

    c = o.G2 if g2 else o.G1
    return np.array(c.affine_to_limbs(c.gen), dtype=np.uint64)


def synth_scalars(seed, n, mont):
    out = np.zeros((n, 4), dtype=np.uint64)
    lib().ref_synth_scalars(seed, n, int(mont), _p(out))
    return out


def synth_dlogs(seed, n):
    out = np.zeros((n, 4), dtype=np.uint64)
    lib().ref_synth_dlogs(seed, n, _p(out))
    return out


def synth_bases(g2, seed, n, nthreads=None):
    out = np.zeros((n, 24 if g2 else 12), dtype=np.uint64)
    gen = generator_limbs(g2)
    lib().ref_synth_bases(int(g2), seed, n, _p(gen), _p(out), nthreads or ncores())
    return out


def msm_by_dlog(g2, seed_bases, scalars_canon):
    """T2: (Σ sᵢkᵢ mod r)·G for the synthetic bases of `seed_bases` → Jacobian limbs."""
    n = scalars_canon.shape[0]
    dot = np.zeros(4, dtype=np.uint64)
    lib().ref_fr_dot_synth(seed_bases, _p(np.ascontiguousarray(scalars_canon)), n, _p(dot))
    return mul(g2, generator_limbs(g2), dot)


def fp_binop(name, a, b):
    out = np.zeros_like(a)
    f = getattr(lib(), name)
    f(_p(a), _p(b), _p(out))
    return out


def affine_equal(g2, jac_a, jac_b):
    """group equality after normalisation to affine — the parity relation of src/tests.rs:66-67"""
    return np.array_equal(to_affine(g2, jac_a), to_affine(g2, jac_b))
