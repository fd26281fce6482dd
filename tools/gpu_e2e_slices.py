"""End-to-end time of b200msm_g1/g2(pinned host buffers) under the current slice schedule (development aid;
sweep with B200MSM_SLICE_RATIO / B200MSM_SLICE_K). usage: gpu_e2e_slices.py g1:20 ..."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
import ark_blst_b200 as eng
from oracle import cref
for spec in sys.argv[1:]:
    g, logn = spec.split(":"); g2 = int(g == "g2"); n = 1 << int(logn); aw = 24 if g2 else 12
    db = torch.empty((n, aw), dtype=torch.int64, device="cuda"); ds = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    eng.synth_bases_device(g2, 1, n, db.data_ptr()); eng.synth_scalars_device(2, n, True, ds.data_ptr())
    hb = torch.empty((n, aw), dtype=torch.int64, pin_memory=True); hb.copy_(db)
    hs = torch.empty((n, 4), dtype=torch.int64, pin_memory=True); hs.copy_(ds); torch.cuda.synchronize()
    hb_np, hs_np = hb.numpy().view(np.uint64), hs.numpy().view(np.uint64)
    grp = eng.G2Projective if g2 else eng.G1Projective
    exp = cref.msm_by_dlog(g2, 1, cref.synth_scalars(2, n, False))
    rb = eng.ResidentBases(grp, hb_np)
    res = {}
    for name, call in (("oneshot", lambda: grp.msm(hb_np, hs_np)), ("resident", lambda: rb.msm(hs_np))):
        for _ in range(4): r = call()
        t0 = time.perf_counter()
        for _ in range(20): r = call()
        res[name] = round((time.perf_counter() - t0) * 50, 3)
        res[name + "_parity"] = bool(cref.affine_equal(g2, r, exp))
    rb.close()
    print(json.dumps({"spec": spec, "ratio": os.environ.get("B200MSM_SLICE_RATIO"), "K": os.environ.get("B200MSM_SLICE_K"), **res}), flush=True)
