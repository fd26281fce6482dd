"""ctypes binding of libb200msm.so (the C-ABI of include/b200msm.h).

There is no fallback of any kind: if the shared library is missing or fails to load, importing
this module raises, and every MSM call fails loudly when no sm_100 device is usable.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200MSM_LIB: another build of the same library (A/B experiments with kernel variants); still no fallback
LIB_PATH = os.environ.get("B200MSM_LIB") or os.path.join(_HERE, "libb200msm.so")

u64p = ctypes.POINTER(ctypes.c_uint64)
i32p = ctypes.POINTER(ctypes.c_int32)
f64p = ctypes.POINTER(ctypes.c_double)
vp = ctypes.c_void_p

# name → (restype, argtypes); every symbol include/b200msm.h declares
SIGNATURES = {
    "b200msm_init": (ctypes.c_int, [ctypes.c_int, ctypes.c_int]),
    "b200msm_shutdown": (None, []),
    "b200msm_device_count": (ctypes.c_int, []),
    "b200msm_last_error": (ctypes.c_char_p, []),
    "b200msm_version": (ctypes.c_char_p, []),
    "b200msm_g1": (ctypes.c_int, [u64p, u64p, ctypes.c_size_t, ctypes.c_int, u64p]),
    "b200msm_g2": (ctypes.c_int, [u64p, u64p, ctypes.c_size_t, ctypes.c_int, u64p]),
    "b200msm_bases_upload": (ctypes.c_int, [ctypes.c_int, u64p, ctypes.c_size_t, ctypes.POINTER(vp)]),
    "b200msm_bases_free": (ctypes.c_int, [vp]),
    "b200msm_run": (ctypes.c_int, [vp, u64p, ctypes.c_size_t, ctypes.c_int, u64p]),
    "b200msm_bases_precompute": (ctypes.c_int, [vp, ctypes.c_int]),
    "b200msm_bases_table_info": (ctypes.c_int, [vp, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_size_t)]),
    "b200msm_table_plan": (ctypes.c_int, [ctypes.c_int, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "b200msm_table_build_device": (ctypes.c_int, [ctypes.c_int, vp, ctypes.c_size_t, ctypes.c_int, vp, vp]),
    "b200msm_run_table_device": (ctypes.c_int, [ctypes.c_int, vp, ctypes.c_size_t, ctypes.c_int, vp, ctypes.c_size_t, ctypes.c_int, vp, vp]),
    "b200msm_run_device": (ctypes.c_int, [ctypes.c_int, vp, vp, ctypes.c_size_t, ctypes.c_int, vp, vp]),
    "b200msm_sum_partials_device": (ctypes.c_int, [ctypes.c_int, vp, ctypes.c_int, vp, vp]),
    "b200msm_normalize_batch": (ctypes.c_int, [ctypes.c_int, u64p, ctypes.c_size_t, u64p]),
    "b200msm_normalize_batch_device": (ctypes.c_int, [ctypes.c_int, vp, ctypes.c_size_t, vp, vp]),
    "b200msm_deserialize": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_uint8), ctypes.c_size_t, ctypes.c_int, ctypes.c_int, u64p, ctypes.POINTER(ctypes.c_uint8)]),
    "b200msm_serialize": (ctypes.c_int, [ctypes.c_int, u64p, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(ctypes.c_uint8)]),
    "b200msm_launch_count": (ctypes.c_ulonglong, []),
    "b200msm_set_window_bits": (ctypes.c_int, [ctypes.c_int]),
    "b200msm_set_glv": (ctypes.c_int, [ctypes.c_int]),
    "b200msm_set_heavy_factor": (ctypes.c_int, [ctypes.c_int]),
    "b200msm_set_batch_affine": (ctypes.c_int, [ctypes.c_int]),
    "b200msm_set_graphs": (ctypes.c_int, [ctypes.c_int]),
    "b200msm_set_lane": (ctypes.c_int, [ctypes.c_int]),
    "b200msm_host_register": (ctypes.c_int, [vp, ctypes.c_size_t]),
    "b200msm_host_unregister": (ctypes.c_int, [vp]),
    "b200msm_set_stream_slices": (ctypes.c_int, [ctypes.c_int, ctypes.c_size_t]),
    "b200msm_set_max_chunk": (ctypes.c_int, [ctypes.c_size_t]),
    "b200msm_set_profiling": (ctypes.c_int, [ctypes.c_int]),
    "b200msm_last_phase_ms": (ctypes.c_int, [f64p]),
    "b200msm_last_plan": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "b200msm_plan_query": (ctypes.c_int, [ctypes.c_int, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]),
    "b200msm_synth_bases_device": (ctypes.c_int, [ctypes.c_int, ctypes.c_uint64, ctypes.c_size_t, vp, vp]),
    "b200msm_synth_scalars_device": (ctypes.c_int, [ctypes.c_uint64, ctypes.c_size_t, ctypes.c_int, vp, vp]),
    "b200msm_imad_peak": (ctypes.c_int, [f64p]),
    "b200msm_dbg_field_op": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, u64p, u64p, u64p, ctypes.c_size_t]),
    "b200msm_dbg_point_op": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, u64p, u64p, u64p, ctypes.c_size_t]),
    "b200msm_dbg_digits": (ctypes.c_int, [u64p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, i32p, ctypes.POINTER(ctypes.c_int)]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C ark_blst_b200/csrc). There is no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        f = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
        f.restype = res
        f.argtypes = args
    return lib


lib = _load()


class B200MsmError(RuntimeError):
    def __init__(self, code, where):
        msg = lib.b200msm_last_error()
        super().__init__(f"{where}: status {code}: {msg.decode() if msg else ''}")
        self.code = code


def check(rc, where):
    if rc != 0:
        raise B200MsmError(rc, where)
