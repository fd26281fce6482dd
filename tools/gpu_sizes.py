"""Device-resident MSM timing over a list of (group, logn) with phase breakdown and a dlog parity
check (development aid; bench.py is the contract). usage: gpu_sizes.py g1:24 g2:20 ..."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
import ark_blst_b200 as eng
from oracle import cref

L = eng._lib.lib
L.b200msm_set_profiling(1)
import os
if os.environ.get("GLV"): L.b200msm_set_glv(int(os.environ["GLV"]))
if os.environ.get("WBITS"): L.b200msm_set_window_bits(int(os.environ["WBITS"]))
if os.environ.get("BA"): L.b200msm_set_batch_affine(int(os.environ["BA"]))
peak = eng.imad_peak()["imad_per_s"]
sys.path.insert(0, ".")
from bench import work_model, FPMUL_IMAD
for spec in sys.argv[1:]:
    g, logn = spec.split(":")
    g2 = 1 if g == "g2" else 0
    n = 1 << int(logn)
    aw = 24 if g2 else 12
    bases = torch.empty((n, aw), dtype=torch.int64, device="cuda")
    scalars = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    t0 = time.time()
    eng.synth_bases_device(g2, 1, n, bases.data_ptr()); torch.cuda.synchronize()
    tg = time.time() - t0
    eng.synth_scalars_device(2, n, True, scalars.data_ptr())
    out = torch.zeros(36 if g2 else 18, dtype=torch.int64, device="cuda")
    best = None
    for it in range(4):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.run_device(g2, bases.data_ptr(), scalars.data_ptr(), n, True, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        e1.record(); e1.synchronize()
        ms = e0.elapsed_time(e1)
        if it and (best is None or ms < best): best = ms
    ph = eng.last_phase_ms()
    plan = eng.last_plan()
    exp = cref.msm_by_dlog(g2, 1, cref.synth_scalars(2, n, False))
    ok = cref.affine_equal(g2, out.cpu().numpy().view(np.uint64), exp)
    c, W, tot, acc = work_model(n, g2)
    print(json.dumps({"group": g, "logn": int(logn), "c*": c, "W*": W, "plan": plan, "gen_s": round(tg, 2), "ms": round(best, 3), "parity": bool(ok),
                      "msm_frac_of_imad_peak": round(tot * FPMUL_IMAD / (best * 1e-3) / peak, 3),
                      "acc_frac": round(acc * FPMUL_IMAD / (ph["accumulate"] * 1e-3) / peak, 3),
                      "mem_gb": round(torch.cuda.mem_get_info()[1] / 1e9 - torch.cuda.mem_get_info()[0] / 1e9, 1),
                      "phases": {k: round(v, 3) for k, v in ph.items() if k != "valid"}}))
    del bases, scalars
    torch.cuda.empty_cache()
