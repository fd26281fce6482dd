"""Quick GPU sanity + phase timing (development aid; bench.py is the contract)."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import torch
import ark_blst_b200 as eng
from oracle import cref

L = eng._lib.lib
print(eng.imad_peak())
L.b200msm_set_profiling(1)
for g2, logn in ((0, 16), (0, 20), (1, 18), (1, 20), (0, 22)):
    n = 1 << logn
    aw = 24 if g2 else 12
    bases = torch.empty((n, aw), dtype=torch.int64, device="cuda")
    scalars = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    t0 = time.time()
    eng.synth_bases_device(g2, 1, n, bases.data_ptr()); torch.cuda.synchronize()
    tg = time.time() - t0
    eng.synth_scalars_device(2, n, True, scalars.data_ptr())
    out = torch.zeros(36 if g2 else 18, dtype=torch.int64, device="cuda")
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        eng.run_device(g2, bases.data_ptr(), scalars.data_ptr(), n, True, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize(); dt = time.time() - t0
    ph = eng.last_phase_ms()
    ok = None
    if logn <= 20:
        exp = cref.msm_by_dlog(g2, 1, cref.synth_scalars(2, n, False))
        ok = cref.affine_equal(g2, out.cpu().numpy().view(np.uint64), exp)
    print(json.dumps({"g2": g2, "logn": logn, "gen_s": round(tg, 3), "wall_ms": round(dt * 1e3, 3), "ok": ok, "phases": {k: round(v, 3) for k, v in ph.items()}}))
