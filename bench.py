#!/usr/bin/env python
"""bench.py — BLS12-381 MSM on B200 through the b200msm C-ABI, with the host-CPU baseline beside it.

Contract (one JSON line from rank 0):
  python bench.py --gpus N --steps K --warmup W                 our arm (CUDA path)
  python bench.py --impl reference --gpus N --steps K --warmup W  the reference's CPU path, timed
                                                                  on the box's host cores
A "step" is one whole MSM (hot path: digits → sort → accumulate → reduce → combine) over one
batch of synthetic input.  Workload at N=1: BASELINE.json configs[1], "G1 MSM 2^20 points on
1×B200".  At N>1 every rank holds its own 2^20-point shard (weak scaling: an N·2^20-point MSM),
computes a partial sum, the 144-byte partials are all-gathered over NCCL and rank 0 adds them.

`value` / `ms_per_step`: device time per MSM with bases and scalars already resident in HBM.
`e2e`: the same MSM through the reference-facing call b200msm_g1(host bases, host scalars) —
H2D of 128 B/point and D2H of the result inside the timed region.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "BLS12-381 G1 MSM ms @2^20/2^24, 1-8 B200, vs blst Pippenger on host"
SEED_BASES = 0xB200_0381_0000_0000
SEED_SCALARS = 0xB200_0381_5CA1_A400
FPMUL_IMAD = 588  # 32-bit IMAD per Fp product


def work_model(n, g2, c=None):
    """MSM(n) = n·W·(1−2^−c)·madd + W·2^(c−1)·2·add + W·(c·dbl+add) Fp-mul at the
    work-minimising c*. Returns (c, W, total Fp-mul, accumulate-only Fp-mul)."""
    madd, add, dbl = (28, 40, 25) if g2 else (10, 14, 9)
    best = None
    for cc in ([c] if c else range(2, 23)):
        W = math.ceil(256 / cc)
        acc = n * W * (1 - 2.0 ** -cc) * madd
        tot = acc + W * 2 ** (cc - 1) * 2 * add + W * (cc * dbl + add)
        if best is None or tot < best[2]:
            best = (cc, W, tot, acc)
    return best


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = sorted(s for s, p in zip(sm, pw) if p >= 0.5 * max(pw)) or sorted(sm)
        return {"sm_mhz": load[len(load) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own CPU path for this MSM (src/g1.rs:602-619 → blstrs multi_exp → blst
    Pippenger), timed on the host cores. blst cannot be built here (Rust, un-vendored, no cargo),
    so this times oracle/libmsm_ref.so — the C restatement of that algorithm — with every host
    thread: cpu_baseline.kind = "port"."""
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    from oracle import cref
    import numpy as np

    g2 = args.group == "g2"
    n_total = (1 << args.logn) * args.gpus
    n_s = min(n_total, 1 << args.ref_sample_logn)
    cores = cref.ncores()
    t0 = time.time()
    bases = cref.synth_bases(int(g2), SEED_BASES, n_s)
    scal = cref.synth_scalars(SEED_SCALARS, n_s, True)
    gen_s = time.time() - t0
    times = []
    out = None
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        out = cref.msm(int(g2), bases, scal, 1, nthreads=cores)
        dt = (time.perf_counter() - t0) * 1e3
        if it >= args.warmup:
            times.append(dt)
    ok = cref.affine_equal(int(g2), out, cref.msm_by_dlog(int(g2), SEED_BASES, cref.synth_scalars(SEED_SCALARS, n_s, False)))
    ms_sample = sum(times) / len(times)
    scale = n_total / n_s
    ms = ms_sample * scale
    sample = ("full workload" if n_s == n_total else
              f"2^{args.ref_sample_logn} of {n_total} points per step, time scaled linearly x{scale:g}")
    line = {
        "impl": "reference", "metric": METRIC, "value": ms, "unit": "ms", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "points_per_s": n_total / (ms * 1e-3),
        "config": {"workload": f"{args.group.upper()} MSM 2^{args.logn} points per GPU x {args.gpus} GPU(s) = {n_total} points",
                   "window_rule": "blst pippenger_window_size", "input_gen_s": round(gen_s, 2), "parity_vs_dlog": bool(ok)},
        "cpu_baseline": {"value": ms, "unit": "ms", "cores": cores, "kind": "port", "sample": sample,
                         "what": "oracle/libmsm_ref.so: C restatement of blst 0.3.10 Pippenger (portable u128 Montgomery, "
                                 "no hand-written asm), pthreads over (window x slice) tiles"},
        "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch

    import ark_blst_b200 as eng

    rank, world, local = dist_env()
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the MSM engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    L = eng._lib.lib
    eng._lib.check(L.b200msm_init(local, 1), "init")

    g2 = args.group == "g2"
    G = eng.G2 if g2 else eng.G1
    n = 1 << args.logn
    n_total = n * world
    aw, jw = (24, 36) if g2 else (12, 18)
    seed_b = SEED_BASES + 1000003 * rank
    seed_s = SEED_SCALARS + 1000003 * rank
    stream = torch.cuda.current_stream().cuda_stream

    bases = torch.empty((n, aw), dtype=torch.int64, device=dev)
    scalars = torch.empty((n, 4), dtype=torch.int64, device=dev)
    eng.synth_bases_device(G, seed_b, n, bases.data_ptr(), stream)
    eng.synth_scalars_device(seed_s, n, True, scalars.data_ptr(), stream)  # Montgomery: what `msm` receives
    torch.cuda.synchronize()

    peak = eng.imad_peak() if rank == 0 else None
    partial = torch.zeros(jw, dtype=torch.int64, device=dev)
    gathered = torch.zeros((world, jw), dtype=torch.int64, device=dev)
    result = torch.zeros(jw, dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    tbl = {}

    def step_device(table=False):
        if table:
            eng.run_table_device(G, tbl["t"].data_ptr(), n, tbl["c"], scalars.data_ptr(), n, True, partial.data_ptr(), stream)
        else:
            eng.run_device(G, bases.data_ptr(), scalars.data_ptr(), n, True, partial.data_ptr(), stream)
        if world > 1:
            dist.all_gather_into_tensor(gathered.view(-1), partial)
            if rank == 0:
                eng.sum_partials_device(G, gathered.data_ptr(), world, result.data_ptr(), stream)
        else:
            result.copy_(partial)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    L.b200msm_set_profiling(1)
    for _ in range(args.warmup):
        flush.zero_()
        step_device()
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = L.b200msm_launch_count()
    phases = []
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                      # L2 flush between steps, outside the per-step events
        ev[k][0].record()
        step_device()
        ev[k][1].record()
        ev[k][1].synchronize()
        phases.append(eng.last_phase_ms())
    plan = eng.last_plan()
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    launches = (L.b200msm_launch_count() - launches0) / args.steps
    dev_ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())

    # ---- e2e: the reference-facing call with HOST buffers (pinned), H2D + D2H inside ----
    hb = torch.empty((n, aw), dtype=torch.int64, pin_memory=True)
    hs = torch.empty((n, 4), dtype=torch.int64, pin_memory=True)
    hb.copy_(bases); hs.copy_(scalars)
    torch.cuda.synchronize()
    hb_np, hs_np = hb.numpy().view(np.uint64), hs.numpy().view(np.uint64)
    grp = eng.G2Projective if g2 else eng.G1Projective
    hpart = torch.zeros(jw, dtype=torch.int64, pin_memory=True)

    def step_e2e():
        out = grp.msm(hb_np, hs_np)                         # b200msm_g1(host, host) → 144 B back
        if world > 1:
            hpart.numpy().view(np.uint64)[:] = out
            partial.copy_(hpart, non_blocking=True)
            dist.all_gather_into_tensor(gathered.view(-1), partial)
            if rank == 0:
                eng.sum_partials_device(G, gathered.data_ptr(), world, result.data_ptr(), stream)
                return result.cpu().numpy().view(np.uint64)
            torch.cuda.synchronize()
        return out

    for _ in range(max(1, args.warmup)):
        e2e_out = step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_out = step_e2e()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None

    my_partial = partial.cpu().numpy().view(np.uint64)

    # ---- resident bases with a fixed-base window table (b200msm_bases_precompute): the prover
    # shape — the proving key's bases stay on the device, only scalars arrive per MSM.  Reported
    # beside the headline, never instead of it: `value` and `e2e` above take fresh bases per call.
    table_info = None
    if not args.no_table:
        c_t, W_t = eng.table_plan(G, n)
        tbl["c"] = c_t
        tbl["t"] = torch.empty((W_t, n, aw), dtype=torch.int64, device=dev)
        tbl["t"][0].copy_(bases)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.table_build_device(G, tbl["t"].data_ptr(), n, c_t, tbl["t"].data_ptr(), stream)
        torch.cuda.synchronize()
        build_ms = (time.perf_counter() - t0) * 1e3
        for _ in range(args.warmup):
            flush.zero_()
            step_device(True)
        barrier()
        tev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        tphases = []
        for k in range(args.steps):
            flush.zero_()
            tev[k][0].record()
            step_device(True)
            tev[k][1].record()
            tev[k][1].synchronize()
            tphases.append(eng.last_phase_ms())
        barrier()
        t = torch.tensor([sum(a.elapsed_time(b) for a, b in tev) / args.steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tbl_ms = float(t.item())
        tbl_partial = partial.cpu().numpy().view(np.uint64).copy()
        tbl_total = result.cpu().numpy().view(np.uint64).copy() if rank == 0 else None
        # end to end: b200msm_bases_upload + b200msm_bases_precompute once, then b200msm_run(host scalars) per MSM
        rb = eng.ResidentBases(grp, hb_np)

        def step_e2e_resident():
            out = rb.msm(hs_np, montgomery=True)
            if world > 1:
                hpart.numpy().view(np.uint64)[:] = out
                partial.copy_(hpart, non_blocking=True)
                dist.all_gather_into_tensor(gathered.view(-1), partial)
                if rank == 0:
                    eng.sum_partials_device(G, gathered.data_ptr(), world, result.data_ptr(), stream)
                    return result.cpu().numpy().view(np.uint64)
                torch.cuda.synchronize()
            return out

        def time_resident():
            for _ in range(max(1, args.warmup)):
                out = step_e2e_resident()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                out = step_e2e_resident()
            barrier()
            tt = torch.tensor([(time.perf_counter() - t0) * 1e3 / args.steps], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return out, float(tt.item())

        res_e2e_out, res_e2e_ms = time_resident()      # resident bases as uploaded (no table): scalars H2D only
        rb.precompute(c_t)
        tbl_e2e_out, tbl_e2e_ms = time_resident()      # the same handle as a fixed-base window table
        t = torch.tensor([tbl_e2e_ms], dtype=torch.float64, device=dev)
        rb.close()
        tph = {k: sum(p[k] for p in tphases) / len(tphases) for k in tphases[0] if k != "valid"}
        table_info = {"ms_per_msm": tbl_ms, "points_per_s": n_total / (tbl_ms * 1e-3), "window_bits": c_t, "windows": W_t,
                      "table_bytes_per_gpu": W_t * n * aw * 8, "build_ms_once": build_ms,
                      "phases_ms": {k: round(v, 4) for k, v in tph.items()},
                      "e2e_ms": float(t.item()), "e2e_h2d_bytes_per_step": n * 32 * world,
                      "e2e_ms_resident_without_table": res_e2e_ms,
                      "api": "b200msm_bases_upload + b200msm_bases_precompute once; per MSM b200msm_run(handle, host scalars) "
                             "(device figure: b200msm_run_table_device)"}

    # ---- parity of what was timed (the oracle is the checker only) + CPU baseline ----
    from oracle import cref

    exp_partial = cref.msm_by_dlog(int(g2), seed_b, cref.synth_scalars(seed_s, n, False))
    ok = cref.affine_equal(int(g2), my_partial, exp_partial)
    okt = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    if world > 1:
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        exp_all = [torch.zeros(jw, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(exp_all, torch.from_numpy(exp_partial.view(np.int64)).to(dev))
    if table_info is not None:
        okt2 = torch.tensor([1 if cref.affine_equal(int(g2), tbl_partial, exp_partial) else 0], dtype=torch.int32, device=dev)
        if world > 1:
            dist.all_reduce(okt2, op=dist.ReduceOp.MIN)
        okt = torch.minimum(okt, okt2)
    parity = bool(okt.item())
    if rank == 0:
        total = result.cpu().numpy().view(np.uint64)
        if world > 1:
            exp_total = np.zeros(jw, dtype=np.uint64)
            for e in exp_all:
                exp_total = cref.add(int(g2), exp_total, e.cpu().numpy().view(np.uint64))
            parity = parity and cref.affine_equal(int(g2), total, exp_total)
        parity = parity and cref.affine_equal(int(g2), e2e_out, total)
        if table_info is not None:
            parity = parity and cref.affine_equal(int(g2), tbl_total, total) and cref.affine_equal(int(g2), tbl_e2e_out, total)
            parity = parity and cref.affine_equal(int(g2), res_e2e_out, total)
            table_info["frac_of_plain_msm_imad"] = None  # filled below

        cpu = None
        if world == 1 and not args.no_cpu:
            cores = cref.ncores()
            best = None
            reps = 2 if args.logn <= 20 else 1
            for _ in range(reps):
                t0 = time.perf_counter()
                cpu_out = cref.msm(int(g2), hb_np, hs_np, 1, nthreads=cores)
                dt = (time.perf_counter() - t0) * 1e3
                best = dt if best is None else min(best, dt)
            parity = parity and cref.affine_equal(int(g2), cpu_out, total)
            cpu = {"value": best, "unit": "ms", "cores": cores, "kind": "port",
                   "sample": f"the full 2^{args.logn}-point workload, best of {reps} runs, same inputs as the GPU step",
                   "what": "oracle/libmsm_ref.so: C restatement of blst 0.3.10 Pippenger (portable u128 Montgomery, "
                           "no hand-written asm); real blst cannot be built here (Rust, no cargo)"}

        c, W, _, fpmul_acc = work_model(n, g2)            # what each GPU's accumulate kernel runs
        _, _, fpmul_total, _ = work_model(n_total, g2)     # single-problem numerator
        ph = {k: sum(p[k] for p in phases) / len(phases) for k in phases[0] if k != "valid"}
        acc_s = ph["accumulate"] * 1e-3
        imad_peak = peak["imad_per_s"]
        achieved = fpmul_acc * FPMUL_IMAD / acc_s
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(f"k_accumulate_{args.group}_2^{args.logn}")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": dev_ms, "unit": "ms", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": False, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32x12 Montgomery limbs (integer)", "data": "synthetic",
            "points_per_s": n_total / (dev_ms * 1e-3),
            "config": {"workload": f"{args.group.upper()} MSM 2^{args.logn} points per GPU x {world} GPU(s) = {n_total} points",
                       "window_bits": c, "windows": W,
                       "engine_plan": {**plan, "note": "glv 2 = scalars split k1 + k2*lambda over (P, phi(P)), unsigned top digit; "
                                                       "window_bits/windows above are the canonical c* of the work model (roofline numerator)"},
                       "scalars": "uniform mod r, Montgomery form (VariableBaseMSM::msm)",
                       "bases": "random subgroup points k_i*G, affine, resident in HBM",
                       "l2": "flushed between steps (256 MiB memset, outside the per-step events)",
                       "parity": "GPU result == (sum s_i k_i)*G and == CPU oracle result" if parity else "MISMATCH"},
            "parity_ok": parity,
            "wall_ms_per_step_incl_flush": wall_ms / args.steps,
            "phases_ms": {k: round(v, 4) for k, v in ph.items()},
            "roofline": {
                "bound": "imad", "kernel": f"k_accumulate<{'fp2' if g2 else 'fp'}> (+heavy-bucket kernels)",
                "achieved": achieved / 1e12, "peak": imad_peak / 1e12, "unit": "TIMAD/s",
                "frac": achieved / imad_peak, "traffic": traffic,
                "peak_source": "measured live by b200msm_imad_peak (mad.lo.u32 issue rate; MEASURED_PEAKS.json has no integer figure)",
                "algorithmic_imad_per_launch": fpmul_acc * FPMUL_IMAD,
                "kernel_ms": ph["accumulate"],
                "whole_msm": {"algorithmic_imad": fpmul_total * FPMUL_IMAD,
                              "frac": fpmul_total * FPMUL_IMAD / (dev_ms * 1e-3) / (imad_peak * world)},
                "gather_gbs": n * W * (192 if g2 else 96) / acc_s / 1e9,
                "hbm_peak_gbs": _measured_hbm(),
            },
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": n * (aw * 8 + 32) * world,
                    "d2h_bytes_per_step": jw * 8 * world, "points_per_s": n_total / (e2e_ms * 1e-3),
                    "api": "b200msm_g1/b200msm_g2(host bases, host scalars) via ark_blst_b200.G?Projective.msm, pinned host memory"},
            "resident_table": table_info,
            "gpu_launches": launches,
            "clocks": clocks,
            "imad_peak": peak,
        }
        if table_info is not None:  # same numerator as the headline (the plain MSM's algorithmic work at c*)
            table_info["frac_of_plain_msm_imad"] = fpmul_total * FPMUL_IMAD / (table_info["ms_per_msm"] * 1e-3) / (imad_peak * world)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_groth16(args):
    """BASELINE.json configs[4]: a Groth16-prover-shaped batch — three G1 MSMs and one G2 MSM of
    2^logn points each, issued back to back, every MSM sharded over the ranks (non-default mode:
    `--workload groth16`, default logn 22). Device-resident inputs; one JSON line from rank 0."""
    import numpy as np
    import torch

    import ark_blst_b200 as eng
    from oracle import cref

    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    L = eng._lib.lib
    eng._lib.check(L.b200msm_init(local, 1), "init")
    n_total = 1 << args.logn
    lo, hi = n_total * rank // world, n_total * (rank + 1) // world
    n = hi - lo
    stream = torch.cuda.current_stream().cuda_stream
    jobs = []  # (group, bases, scalars, seeds)
    for k, g2 in enumerate((1, 0, 0, 0)):  # G2 first: the tail left exposed at the end is a G1 one
        aw = 24 if g2 else 12
        sb, ss = SEED_BASES + 7919 * k + 1000003 * rank, SEED_SCALARS + 7919 * k + 1000003 * rank
        b = torch.empty((n, aw), dtype=torch.int64, device=dev)
        s = torch.empty((n, 4), dtype=torch.int64, device=dev)
        eng.synth_bases_device(g2, sb, n, b.data_ptr(), stream)
        eng.synth_scalars_device(ss, n, True, s.data_ptr(), stream)
        canon = None
        if args.scalars == "witness":   # ≈40 % zeros, ≈20 % ones, ≈10 % below 2^32, rest uniform (canonical form)
            rng = np.random.default_rng(ss & 0xffffffff)
            u = rng.random(n)
            canon = cref.synth_scalars(ss, n, False)
            small = rng.integers(0, 1 << 32, size=n, dtype=np.uint64)
            canon[u < 0.7] = 0
            canon[(u >= 0.4) & (u < 0.6), 0] = 1
            mid = (u >= 0.6) & (u < 0.7)
            canon[mid, 0] = small[mid]
            s.copy_(torch.from_numpy(canon.view(np.int64)))
        tc = 0
        if args.table:  # resident proving key: fixed-base window table per MSM, built once outside the timed region
            tc, tw = eng.table_plan(g2, n)
            t = torch.empty((tw, n, aw), dtype=torch.int64, device=dev)
            t[0].copy_(b)
            eng.table_build_device(g2, t.data_ptr(), n, tc, t.data_ptr(), stream)
            b = t
        jobs.append((g2, b, s, sb, ss, torch.zeros(36 if g2 else 18, dtype=torch.int64, device=dev), tc, canon))
    torch.cuda.synchronize()
    peak = eng.imad_peak()["imad_per_s"] if rank == 0 else None

    lanes = max(1, min(args.lanes, len(jobs)))
    lane_streams = [torch.cuda.Stream(device=dev) for _ in range(lanes)] if lanes > 1 else []

    def step():
        # the four MSMs go out on `lanes` engine lanes / streams (b200msm_set_lane): the latency-bound tail of one
        # overlaps the accumulation of the next; the partials are gathered once all four are in
        cur = torch.cuda.current_stream()
        for k, (g2, b, s, _, _, part, tc, _c) in enumerate(jobs):
            st = stream
            if lanes > 1:
                L.b200msm_set_lane(k % lanes)
                lane_streams[k % lanes].wait_stream(cur)
                st = lane_streams[k % lanes].cuda_stream
            if tc:
                eng.run_table_device(g2, b.data_ptr(), n, tc, s.data_ptr(), n, _c is None, part.data_ptr(), st)
            else:
                eng.run_device(g2, b.data_ptr(), s.data_ptr(), n, _c is None, part.data_ptr(), st)
        if lanes > 1:
            L.b200msm_set_lane(0)
            for ls in lane_streams:
                cur.wait_stream(ls)
        outs = []
        for g2, b, s, _, _, part, tc, _c in jobs:
            if world > 1:
                gathered = torch.empty((world, part.numel()), dtype=torch.int64, device=dev)
                dist.all_gather_into_tensor(gathered.view(-1), part)
                if rank == 0:
                    res = torch.zeros_like(part)
                    eng.sum_partials_device(g2, gathered.data_ptr(), world, res.data_ptr(), stream)
                    outs.append(res)
            else:
                outs.append(part)
        return outs

    for _ in range(args.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        outs = step()
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ok = True
    for g2, b, s, sb, ss, part, _, canon in jobs:   # every rank checks its own partial against the dlog closed form
        exp = cref.msm_by_dlog(g2, sb, canon if canon is not None else cref.synth_scalars(ss, n, False))
        ok = ok and cref.affine_equal(g2, part.cpu().numpy().view(np.uint64), exp)
    okt = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    if world > 1:
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    if rank == 0:
        imad = (3 * work_model(n_total, False)[2] + work_model(n_total, True)[2]) * FPMUL_IMAD
        print(json.dumps({
            "metric": "Groth16-shaped batch: 3xG1 + 1xG2 MSM, ms per batch", "value": ms, "unit": "ms", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32x12 Montgomery limbs (integer)", "data": "synthetic",
            "config": {"workload": f"3xG1 + 1xG2 MSM at 2^{args.logn} points each, sharded over {world} GPU(s), {'witness-like' if args.scalars == 'witness' else 'uniform'} scalars",
                       "lanes": lanes,
                       "bases": "resident fixed-base window tables (b200msm_table_build_device, built once)" if args.table else "resident affine bases"},
            "parity_ok": bool(okt.item()),
            "roofline": {"bound": "imad", "achieved": imad / (ms * 1e-3) / 1e12, "peak": peak * world / 1e12, "unit": "TIMAD/s",
                         "frac": imad / (ms * 1e-3) / (peak * world), "traffic": None, "algorithmic_imad": imad}}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def _measured_hbm():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return 6650.0  # B200_PROFILING.md fallback


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--group", default="g1", choices=["g1", "g2"])
    ap.add_argument("--logn", type=int, default=None, help="log2 of points per GPU (default 20; groth16 workload: total points, default 22)")
    ap.add_argument("--workload", default="msm", choices=["msm", "groth16"])
    ap.add_argument("--ref-sample-logn", type=int, default=20, help="reference arm: points actually timed per step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--scalars", default="uniform", choices=["uniform", "witness"], help="groth16 workload: scalar distribution")
    ap.add_argument("--lanes", type=int, default=4, help="groth16 workload: engine lanes / streams the four MSMs are spread over (1 = back to back on one stream)")
    ap.add_argument("--table", action="store_true", help="groth16 workload: run every MSM against a resident fixed-base window table")
    ap.add_argument("--no-table", action="store_true", help="skip the resident-bases fixed-base-table leg")
    args = ap.parse_args()
    if args.logn is None:
        args.logn = 22 if args.workload == "groth16" else 20
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.workload == "groth16" and args.impl == "b200":
        return run_groth16(args)
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
