"""Repeats the back-to-back table + plain MSM of tests/test_gpu_table.py::test_table_dlog_closed_form (and a few
neighbours) many times in one process and counts parity failures (development aid for race hunting)."""
import json, os, sys
import numpy as np
sys.path.insert(0, ".")
import torch
import ark_blst_b200 as eng
from oracle import cref

L = eng._lib.lib
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
fails = {"table": 0, "plain": 0, "plain_again": 0, "g2": 0}
st = torch.cuda.current_stream().cuda_stream
junk = []
for it in range(reps):
    for g2, logn in ((0, 20), (0, 18)):
        n = 1 << logn; aw = 24 if g2 else 12
        sb, ss = 0xB2000381_00001000 + logn + it, 177 + logn + it
        c, W = eng.table_plan(g2, n)
        table = torch.empty((W, n, aw), dtype=torch.int64, device="cuda")
        scalars = torch.empty((n, 4), dtype=torch.int64, device="cuda")
        eng.synth_bases_device(g2, sb, n, table.data_ptr())
        eng.synth_scalars_device(ss, n, True, scalars.data_ptr())
        eng.table_build_device(g2, table.data_ptr(), n, c, table.data_ptr(), st)
        out = torch.zeros((3, 36 if g2 else 18), dtype=torch.int64, device="cuda")
        eng.run_table_device(g2, table.data_ptr(), n, c, scalars.data_ptr(), n, True, out[0].data_ptr(), st)
        eng.run_device(g2, table.data_ptr(), scalars.data_ptr(), n, True, out[1].data_ptr(), st)
        eng.run_device(g2, table.data_ptr(), scalars.data_ptr(), n, True, out[2].data_ptr(), st)
        torch.cuda.synchronize()
        got = out.cpu().numpy().view(np.uint64)
        exp = cref.msm_by_dlog(g2, sb, cref.synth_scalars(ss, n, False))
        for k, name in enumerate(("table", "plain", "plain_again")):
            if not cref.affine_equal(g2, got[k], exp):
                fails[name] += 1
        if it % 3 == 0:
            junk.append(torch.empty((1 << 20) * (it + 1), dtype=torch.int64, device="cuda"))   # perturb the allocator
        del table, scalars, out
    if it % 4 == 3:
        junk.clear(); torch.cuda.empty_cache()
print(json.dumps({"reps": reps, "fails": fails, "env": {k: v for k, v in os.environ.items() if k.startswith("B200MSM")}}))
