"""End-to-end time of the one-shot host-buffer call b200msm_g1/g2 (pinned host memory, H2D inside)
for several slice counts of the streamed upload (development aid). usage: gpu_e2e.py g1:20 g2:20 ..."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
import ark_blst_b200 as eng
from oracle import cref

L = eng._lib.lib
for spec in sys.argv[1:]:
    g, logn = spec.split(":")
    g2 = 1 if g == "g2" else 0
    n = 1 << int(logn)
    aw = 24 if g2 else 12
    bases = torch.empty((n, aw), dtype=torch.int64, device="cuda")
    scalars = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    eng.synth_bases_device(g2, 1, n, bases.data_ptr())
    eng.synth_scalars_device(2, n, True, scalars.data_ptr())
    hb = torch.empty((n, aw), dtype=torch.int64, pin_memory=True); hb.copy_(bases)
    hs = torch.empty((n, 4), dtype=torch.int64, pin_memory=True); hs.copy_(scalars)
    torch.cuda.synchronize()
    hb_np, hs_np = hb.numpy().view(np.uint64), hs.numpy().view(np.uint64)
    import os
    if os.environ.get("PAGEABLE"):   # ordinary (pageable) host memory, as a Rust Vec would be
        hb_np, hs_np = hb_np.copy(), hs_np.copy()
        if os.environ.get("PAGEABLE") == "register":   # ... then page-locked in place through the library
            L.b200msm_init(-1, 1)
            assert L.b200msm_host_register(hb_np.ctypes.data, hb_np.nbytes) == 0
            assert L.b200msm_host_register(hs_np.ctypes.data, hs_np.nbytes) == 0
    grp = eng.G2Projective if g2 else eng.G1Projective
    exp = cref.msm_by_dlog(g2, 1, cref.synth_scalars(2, n, False))
    res = {}
    for slices in (1, 2, 4, 8):
        L.b200msm_set_stream_slices(slices, 1 if slices > 1 else 0)
        for _ in range(3): out = grp.msm(hb_np, hs_np)
        t0 = time.perf_counter()
        for _ in range(10): out = grp.msm(hb_np, hs_np)
        res[slices] = {"ms": round((time.perf_counter() - t0) * 100, 3), "parity": bool(cref.affine_equal(g2, out, exp))}
    rb = eng.ResidentBases(grp, hb_np)
    by_slices = {}
    for slices in (1, 2, 4, 8):
        L.b200msm_set_stream_slices(slices, 1 if slices > 1 else 0)
        for _ in range(3): out = rb.msm(hs_np)
        t0 = time.perf_counter()
        for _ in range(10): out = rb.msm(hs_np)
        by_slices[slices] = round((time.perf_counter() - t0) * 100, 3)
    print(json.dumps({"resident_by_slices": by_slices}), flush=True)
    L.b200msm_set_stream_slices(8, 0)
    for _ in range(3): out = rb.msm(hs_np)
    t0 = time.perf_counter()
    for _ in range(10): out = rb.msm(hs_np)
    res_ms = round((time.perf_counter() - t0) * 100, 3)
    rb.precompute()
    by_slices = {}
    for slices in (1, 2, 4, 8):
        L.b200msm_set_stream_slices(slices, 1 if slices > 1 else 0)
        for _ in range(3): out = rb.msm(hs_np)
        t0 = time.perf_counter()
        for _ in range(10): out = rb.msm(hs_np)
        by_slices[slices] = round((time.perf_counter() - t0) * 100, 3)
    print(json.dumps({"table_by_slices": by_slices}), flush=True)
    L.b200msm_set_stream_slices(8, 0)
    for _ in range(3): out = rb.msm(hs_np)
    t0 = time.perf_counter()
    for _ in range(10): out = rb.msm(hs_np)
    tbl_ms = round((time.perf_counter() - t0) * 100, 3)
    rb.close()
    print(json.dumps({"group": g, "logn": int(logn), "pageable": bool(os.environ.get("PAGEABLE")), "e2e_by_slices": res,
                      "resident_ms": res_ms, "resident_table_ms": tbl_ms, "table_parity": bool(cref.affine_equal(g2, out, exp))}), flush=True)
