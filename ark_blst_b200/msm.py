"""Host-side mirror of the reference's arkworks surface for the MSM path.

Mirrors `impl VariableBaseMSM for G1Projective / G2Projective` (reference src/g1.rs:602-632,
src/g2.rs:582-612): `msm(bases, scalars)` takes affine bases and Montgomery `Scalar`s,
`msm_bigint(bases, bigints)` takes canonical BigInt<4>s; both return the projective sum or
raise `MsmError(usize)` — the Python spelling of `Result<Self, usize>`:
  * length mismatch → MsmError(min(len))   (arkworks' convention for `msm`)
  * any device error → MsmError(0)          (reference GPU arm, src/g1.rs:628-630)
Arrays are numpy uint64 in the reference's own memory layouts (see include/b200msm.h); nothing is
converted on the host.  All arithmetic happens in libb200msm.so on the GPU; there is no fallback.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import lib

G1, G2 = 0, 1


class MsmError(Exception):
    """Err(usize) of `VariableBaseMSM::msm`."""

    def __init__(self, value, detail=""):
        super().__init__(f"Err({value}) {detail}".strip())
        self.value = value


def _u64(a, cols):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if a.ndim == 1:
        a = a.reshape(-1, cols)
    if a.ndim != 2 or a.shape[1] != cols:
        raise ValueError(f"expected (n, {cols}) uint64 array, got {a.shape}")
    return a


def _ptr(a):
    return a.ctypes.data_as(_lib.u64p)


class _Group:
    GROUP = G1
    AFFINE_WORDS = 12   # u64 per MulBase (G1Affine = blst_p1_affine, src/g1.rs:54-56)
    PROJ_WORDS = 18     # u64 per Self (G1Projective = blst_p1, src/g1.rs:435-437)
    NEGATION_IS_CHEAP = True  # ScalarMul const, src/g1.rs:595

    @classmethod
    def _call(cls, bases, scalars, mont):
        bases = _u64(bases, cls.AFFINE_WORDS)
        scalars = _u64(scalars, 4)
        if bases.shape[0] != scalars.shape[0]:
            raise MsmError(min(bases.shape[0], scalars.shape[0]), "length mismatch")
        out = np.zeros(cls.PROJ_WORDS, dtype=np.uint64)
        f = lib.b200msm_g2 if cls.GROUP == G2 else lib.b200msm_g1
        rc = f(_ptr(bases), _ptr(scalars), scalars.shape[0], int(mont), _ptr(out))
        if rc != 0:
            raise MsmError(0, (lib.b200msm_last_error() or b"").decode())
        return out

    @classmethod
    def msm(cls, bases, scalars):
        """VariableBaseMSM::msm(&[MulBase], &[ScalarField]) — scalars in Montgomery form."""
        return cls._call(bases, scalars, True)

    @classmethod
    def msm_bigint(cls, bases, bigints):
        """VariableBaseMSM::msm_bigint(&[MulBase], &[BigInt<4>]) — canonical little-endian limbs."""
        return cls._call(bases, bigints, False)

    @classmethod
    def normalize_batch(cls, projective):
        """CurveGroup::normalize_batch(&[Self]) -> Vec<Affine> (reference src/g1.rs:536-543)."""
        proj = _u64(projective, cls.PROJ_WORDS)
        out = np.zeros((proj.shape[0], cls.AFFINE_WORDS), dtype=np.uint64)
        _lib.check(lib.b200msm_normalize_batch(cls.GROUP, _ptr(proj), proj.shape[0], _ptr(out)), "normalize_batch")
        return out

    @classmethod
    def serialize(cls, affine, compressed=True):
        """CanonicalSerialize::serialize_{compressed,uncompressed} for a batch of affine points
        (reference src/g1.rs:358-373) → (n, 48|96|192) uint8"""
        aff = _u64(affine, cls.AFFINE_WORDS)
        eb = cls.AFFINE_WORDS * 8 // (2 if compressed else 1)
        out = np.zeros((aff.shape[0], eb), dtype=np.uint8)
        u8 = ctypes.POINTER(ctypes.c_uint8)
        _lib.check(lib.b200msm_serialize(cls.GROUP, _ptr(aff), aff.shape[0], int(compressed), out.ctypes.data_as(u8)), "serialize")
        return out

    @classmethod
    def deserialize(cls, data, compressed=True, validate=True):
        """CanonicalDeserialize::deserialize_with_mode for a batch (reference src/g1.rs:398-431)
        → (affine (n, 12|24) uint64, status (n,) uint8: 0 ok, 1 malformed, 2 fails Valid::check)"""
        eb = cls.AFFINE_WORDS * 8 // (2 if compressed else 1)
        data = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1, eb)
        n = data.shape[0]
        aff = np.zeros((n, cls.AFFINE_WORDS), dtype=np.uint64)
        st = np.zeros(n, dtype=np.uint8)
        u8 = ctypes.POINTER(ctypes.c_uint8)
        _lib.check(lib.b200msm_deserialize(cls.GROUP, data.ctypes.data_as(u8), n, int(compressed), int(validate), _ptr(aff),
                                           st.ctypes.data_as(u8)), "deserialize")
        return aff, st

    @classmethod
    def msm_unchecked(cls, bases, scalars):
        """arkworks' msm_unchecked: truncates to the shorter input instead of erring."""
        bases = _u64(bases, cls.AFFINE_WORDS)
        scalars = _u64(scalars, 4)
        n = min(bases.shape[0], scalars.shape[0])
        return cls._call(bases[:n], scalars[:n], True)


class G1Projective(_Group):
    GROUP, AFFINE_WORDS, PROJ_WORDS = G1, 12, 18


class G2Projective(_Group):
    GROUP, AFFINE_WORDS, PROJ_WORDS = G2, 24, 36


class ResidentBases:
    """bases uploaded once (sharded across the bound GPUs), many scalar vectors."""

    def __init__(self, group_cls, bases):
        self.cls = group_cls
        bases = _u64(bases, group_cls.AFFINE_WORDS)
        self.n = bases.shape[0]
        self._h = ctypes.c_void_p()
        _lib.check(lib.b200msm_bases_upload(group_cls.GROUP, _ptr(bases), self.n, ctypes.byref(self._h)), "bases_upload")

    def precompute(self, window_bits=0):
        """Turn the resident bases into a fixed-base window table (table[w][i] = 2^(c·w)·P_i): later
        msm() calls use one bucket set for all windows and no Horner chain. Returns (c, windows, bytes)."""
        _lib.check(lib.b200msm_bases_precompute(self._h, int(window_bits)), "bases_precompute")
        c, w, b = ctypes.c_int(), ctypes.c_int(), ctypes.c_size_t()
        _lib.check(lib.b200msm_bases_table_info(self._h, ctypes.byref(c), ctypes.byref(w), ctypes.byref(b)), "bases_table_info")
        return c.value, w.value, b.value

    def msm(self, scalars, montgomery=True):
        scalars = _u64(scalars, 4)
        if scalars.shape[0] > self.n:
            raise MsmError(min(self.n, scalars.shape[0]), "more scalars than resident bases")
        out = np.zeros(self.cls.PROJ_WORDS, dtype=np.uint64)
        rc = lib.b200msm_run(self._h, _ptr(scalars), scalars.shape[0], int(montgomery), _ptr(out))
        if rc != 0:
            raise MsmError(0, (lib.b200msm_last_error() or b"").decode())
        return out

    def close(self):
        if self._h:
            lib.b200msm_bases_free(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- device-pointer helpers (torch tensors / raw pointers already in HBM) ----
def run_device(group, d_bases, d_scalars, n, montgomery, d_out, stream=0):
    _lib.check(lib.b200msm_run_device(group, d_bases, d_scalars, n, int(montgomery), d_out, stream), "run_device")


def table_plan(group, n, window_bits=0):
    """(c, windows) of the fixed-base table for n points; the table holds windows × n affine points"""
    c, w = ctypes.c_int(int(window_bits)), ctypes.c_int()
    _lib.check(lib.b200msm_table_plan(group, n, ctypes.byref(c), ctypes.byref(w)), "table_plan")
    return c.value, w.value


def table_build_device(group, d_bases, n, window_bits, d_table, stream=0):
    _lib.check(lib.b200msm_table_build_device(group, d_bases, n, int(window_bits), d_table, stream), "table_build_device")


def run_table_device(group, d_table, stride, window_bits, d_scalars, n, montgomery, d_out, stream=0):
    _lib.check(lib.b200msm_run_table_device(group, d_table, stride, int(window_bits), d_scalars, n, int(montgomery), d_out, stream),
               "run_table_device")


def sum_partials_device(group, d_partials, count, d_out, stream=0):
    _lib.check(lib.b200msm_sum_partials_device(group, d_partials, count, d_out, stream), "sum_partials_device")


def synth_bases_device(group, seed, n, d_out, stream=0):
    _lib.check(lib.b200msm_synth_bases_device(group, seed, n, d_out, stream), "synth_bases_device")


def synth_scalars_device(seed, n, montgomery, d_out, stream=0):
    _lib.check(lib.b200msm_synth_scalars_device(seed, n, int(montgomery), d_out, stream), "synth_scalars_device")


def imad_peak():
    out = (ctypes.c_double * 3)()
    _lib.check(lib.b200msm_imad_peak(out), "imad_peak")
    return {"imad_per_s": out[0], "imad_wide_x2_per_s": out[1], "sm_mhz_est": out[2]}


def last_plan():
    out = (ctypes.c_int * 4)()
    _lib.check(lib.b200msm_last_plan(out), "last_plan")
    return {"window_bits": out[0], "windows": out[1], "glv": out[2], "table": out[3]}


def last_phase_ms():
    out = (ctypes.c_double * 8)()
    _lib.check(lib.b200msm_last_phase_ms(out), "last_phase_ms")
    names = ["digits", "sort", "bounds_order", "accumulate", "reduce", "combine", "total", "valid"]
    return dict(zip(names, list(out)))
