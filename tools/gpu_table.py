"""Fixed-base window table (b200msm_table_build_device / b200msm_run_table_device): build time,
device-resident MSM time with phase breakdown, and a dlog parity check, next to the plain path on
the same inputs (development aid; bench.py is the contract). usage: gpu_table.py g1:20 g1:20:18 g2:20 ..."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
import ark_blst_b200 as eng
from oracle import cref
from bench import work_model, FPMUL_IMAD

L = eng._lib.lib
L.b200msm_set_profiling(1)
import os
if os.environ.get("TBL_HEAVY"): L.b200msm_set_heavy_factor(int(os.environ["TBL_HEAVY"]))
peak = eng.imad_peak()["imad_per_s"]
st = torch.cuda.current_stream().cuda_stream


def timed(f, reps=4):
    best = None
    for it in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); e1.synchronize()
        ms = e0.elapsed_time(e1)
        if it and (best is None or ms < best): best = ms
    return best


for spec in sys.argv[1:]:
    parts = spec.split(":")
    g2 = 1 if parts[0] == "g2" else 0
    logn = int(parts[1]); n = 1 << logn
    c, W = eng.table_plan(g2, n, int(parts[2]) if len(parts) > 2 else 0)
    aw = 24 if g2 else 12
    table = torch.empty((W, n, aw), dtype=torch.int64, device="cuda")
    scalars = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    eng.synth_bases_device(g2, 1, n, table.data_ptr())
    eng.synth_scalars_device(2, n, True, scalars.data_ptr())
    torch.cuda.synchronize(); t0 = time.time()
    eng.table_build_device(g2, table.data_ptr(), n, c, table.data_ptr(), st)
    torch.cuda.synchronize(); tb = time.time() - t0
    out = torch.zeros((2, 36 if g2 else 18), dtype=torch.int64, device="cuda")
    ms_t = timed(lambda: eng.run_table_device(g2, table.data_ptr(), n, c, scalars.data_ptr(), n, True, out[0].data_ptr(), st))
    ph_t = eng.last_phase_ms()
    ms_p = timed(lambda: eng.run_device(g2, table.data_ptr(), scalars.data_ptr(), n, True, out[1].data_ptr(), st))
    ph_p = eng.last_phase_ms()
    exp = cref.msm_by_dlog(g2, 1, cref.synth_scalars(2, n, False))
    got = out.cpu().numpy().view(np.uint64)
    cs, Ws, tot, acc = work_model(n, g2)
    print(json.dumps({"group": parts[0], "logn": logn, "table_c": c, "table_W": W, "table_gib": round(W * n * aw * 8 / 2**30, 2),
                      "build_s": round(tb, 3), "table_ms": round(ms_t, 3), "plain_ms": round(ms_p, 3),
                      "parity_table": bool(cref.affine_equal(g2, got[0], exp)), "parity_plain": bool(cref.affine_equal(g2, got[1], exp)),
                      "table_frac_of_plain_roofline": round(tot * FPMUL_IMAD / (ms_t * 1e-3) / peak, 3),
                      "table_phases": {k: round(v, 3) for k, v in ph_t.items() if k != "valid"},
                      "plain_phases": {k: round(v, 3) for k, v in ph_p.items() if k != "valid"}}), flush=True)
    del table, scalars
    torch.cuda.empty_cache()
