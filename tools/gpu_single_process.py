"""Single-process N-GPU host-buffer MSM timing with the library's host-side trace (development aid).
usage: B200MSM_TRACE=1 python tools/gpu_single_process.py <ndev> g1:22 g2:20"""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import torch
import ark_blst_b200 as eng
from oracle import cref
L = eng._lib.lib
nd = int(sys.argv[1])
assert L.b200msm_init(0, nd) == 0
for spec in sys.argv[2:]:
    g, logn = spec.split(":"); g2 = int(g == "g2"); n = 1 << int(logn); aw = 24 if g2 else 12
    db = torch.empty((n, aw), dtype=torch.int64, device="cuda:0"); ds = torch.empty((n, 4), dtype=torch.int64, device="cuda:0")
    eng.synth_bases_device(g2, 7, n, db.data_ptr(), 0); eng.synth_scalars_device(8, n, True, ds.data_ptr(), 0)
    hb = torch.empty((n, aw), dtype=torch.int64, pin_memory=True); hb.copy_(db)
    hs = torch.empty((n, 4), dtype=torch.int64, pin_memory=True); hs.copy_(ds); torch.cuda.synchronize()
    hb_np, hs_np = hb.numpy().view(np.uint64), hs.numpy().view(np.uint64)
    grp = eng.G2Projective if g2 else eng.G1Projective
    exp = cref.msm_by_dlog(g2, 7, cref.synth_scalars(8, n, False))
    rb = eng.ResidentBases(grp, hb_np)
    for name, call in (("oneshot", lambda: grp.msm(hb_np, hs_np)), ("resident", lambda: rb.msm(hs_np))):
        for _ in range(3): r = call()
        sys.stderr.write(f"---- timed {spec} {name}\n")
        t0 = time.perf_counter()
        for _ in range(4): r = call()
        print(json.dumps({"spec": spec, "ndev": nd, "path": name, "ms": (time.perf_counter() - t0) * 250, "parity": bool(cref.affine_equal(g2, r, exp))}), flush=True)
    rb.close()
L.b200msm_shutdown()
