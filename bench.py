#!/usr/bin/env python
"""bench.py — BLS12-381 MSM on B200 through the b200msm C-ABI, with the host-CPU baseline beside it.

Contract (one JSON line from rank 0):
  python bench.py --gpus N --steps K --warmup W                 our arm (CUDA path)
  python bench.py --impl reference --gpus N --steps K --warmup W  the reference's CPU path, timed
                                                                  on the box's host cores
A "step" is one whole MSM (hot path: digits → sort → accumulate → reduce → combine) over one
batch of synthetic input.  Headline workload at N=1: BASELINE.json configs[1], "G1 MSM 2^20 points
on 1×B200".  At N>1 every rank holds its own 2^20-point shard (weak scaling: an N·2^20-point MSM),
computes a partial sum, the 144-byte partials are all-gathered over NCCL and rank 0 adds them.

`value` / `ms_per_step`: device time per MSM with bases and scalars already resident in HBM.
`e2e`: the same MSM through the reference-facing call b200msm_g1(host bases, host scalars) —
H2D of 128 B/point and D2H of the result inside the timed region (pinned host memory;
`e2e_pageable` is the same from ordinary pageable memory, what a Rust Vec is).

Beside the headline the same line carries one block per remaining BASELINE.json config, each with
its own parity check (`--blocks` selects; default all):
  c0_2p16      configs[0]  G1 2^16, the benches/group.rs shape: GPU ms, and the CPU port's ms at N=1
  strong_2p24  configs[2]  G1 2^24 sharded over the N ranks (strong scaling: total work fixed)
  g2_2p20      configs[3]  G2 2^20 sharded over the N ranks
  groth16_2p22 configs[4]  3×G1 + 1×G2 at 2^22, every MSM sharded over the N ranks, uniform and
                           witness-like scalars
  single_process (N>1)     the path a Rust caller takes: ONE process, b200msm_init(0, N),
                           b200msm_g1(host, host) sharding G1 2^24 over the N GPUs inside the library
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
ALL_BLOCKS = ("c0_2p16", "g2_2p20", "strong_2p24", "groth16_2p22", "single_process")


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


def work_model_counted(nonzero_digits, n, g2):
    """The same model with the accumulate term COUNTED instead of assumed uniform: `nonzero_digits`
    is the number of non-zero Booth digits of the actual scalars at the canonical c*(n) — the
    bucket additions an MSM of these scalars has to perform (witness-like scalars are mostly 0/1)."""
    madd, add, dbl = (28, 40, 25) if g2 else (10, 14, 9)
    c, W, _, _ = work_model(n, g2)
    return nonzero_digits * madd + W * 2 ** (c - 1) * 2 * add + W * (c * dbl + add)


def executed_accumulate_fpmul(n, g2, plan, ba_rounds):
    """Fp products the accumulation phase actually EXECUTES under the engine's own plan (not the SURVEY 8d
    model, which counts the XYZZ formulation at c*): with R batched-affine pairing rounds, round r runs
    entries/2^r slots at 6 products each (Fp2: 17) plus the second level of the batch inversion (3/K per slot,
    K = 32), and the XYZZ mixed additions (10 / 28) only see what is left, entries/2^R."""
    madd, aff = (28, 17) if g2 else (10, 6)
    lvl2 = (9 if g2 else 3) / 32.0
    c, W, glv = plan["window_bits"], plan["windows"], plan["glv"]
    parts = 1 if glv == 0 else (2 if glv <= 2 else 4)          # b200msm_last_plan: 1 / 2 two parts, 3 / 4 four parts (G2)
    entries = n * parts * (W - 1 if glv in (2, 4) else W) * (1 - 2.0 ** -c)
    tot = 0.0
    for r in range(1, ba_rounds + 1):
        tot += entries / 2 ** r * (aff + lvl2)
    return tot + entries / 2 ** ba_rounds * madd


def count_nonzero_booth_digits(canon, c):
    """canon: (n, 4) uint64 canonical scalars. Number of non-zero signed c-bit window digits
    d_w = u_w + b[wc−1] − 2^c·b[wc+c−1] over W = ⌈256/c⌉ windows (numpy, exact)."""
    import numpy as np

    n = canon.shape[0]
    W = math.ceil(256 / c)
    limbs = np.concatenate([canon, np.zeros((n, 2), dtype=np.uint64)], axis=1)
    total = 0
    for w in range(W):
        lo = w * c - 1                       # bits [lo, lo + c] → c+1 bits (bit −1 of window 0 is 0)
        if lo < 0:
            v = (limbs[:, 0] << np.uint64(1)) & np.uint64((1 << (c + 1)) - 1)
        else:
            q, r = divmod(lo, 64)
            v = limbs[:, q] >> np.uint64(r)
            if r + c + 1 > 64:
                v = v | (limbs[:, q + 1] << np.uint64(64 - r))
            v = v & np.uint64((1 << (c + 1)) - 1)
        u = (v >> np.uint64(1)) & np.uint64((1 << c) - 1)
        d = u.astype(np.int64) + (v & np.uint64(1)).astype(np.int64) - (((v >> np.uint64(c)) & np.uint64(1)).astype(np.int64) << c)
        total += int(np.count_nonzero(d))
    return total


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


CPU_WHAT = ("oracle/libmsm_ref.so: C restatement of blst 0.3.10 Pippenger (window rule, Booth digits, XYZZ buckets, blst's "
            "(slices x windows) tile grid on pthreads); Montgomery product = {mul}; real blst cannot be built here (Rust, "
            "un-vendored, no cargo)")


def cpu_what(cref):
    adx = bool(cref.lib().ref_mul_impl())
    return CPU_WHAT.format(mul="MULX/ADCX/ADOX inline asm (blst's mulx_mont_384 instruction mix)" if adx
                           else "portable u128 (this CPU lacks ADX/BMI2)")


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own CPU path for this MSM (src/g1.rs:602-619 → blstrs multi_exp → blst
    Pippenger), timed on the host cores. blst cannot be built here (Rust, un-vendored, no cargo),
    so this times oracle/libmsm_ref.so — the C restatement of that algorithm, with an ADX/MULX
    Montgomery product where the CPU has it — with every host thread: cpu_baseline.kind = "port".
    The FULL workload of our arm's config is timed at every N (N·2^20 points, no extrapolation)."""
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    from oracle import cref
    import numpy as np

    g2 = args.group == "g2"
    n_total = (1 << args.logn) * args.gpus
    cores = cref.ncores()
    t0 = time.time()
    bases = cref.synth_bases(int(g2), SEED_BASES, n_total)
    scal = cref.synth_scalars(SEED_SCALARS, n_total, True)
    gen_s = time.time() - t0
    times = []
    out = None
    steps, warmup = args.steps, args.warmup
    budget_s = 200.0      # the whole arm ends within a few minutes: fewer timed steps when one is long (reported)
    t_arm = time.perf_counter()
    it = 0
    while it < warmup + steps:
        t0 = time.perf_counter()
        out = cref.msm(int(g2), bases, scal, 1, nthreads=cores)
        dt = (time.perf_counter() - t0) * 1e3
        if it >= warmup:
            times.append(dt)
        it += 1
        if it == 1 and dt * 1e-3 * (warmup + steps) > budget_s:
            warmup = 1
            steps = max(2, min(steps, int(budget_s / (dt * 1e-3)) - 1))
    ok = cref.affine_equal(int(g2), out, cref.msm_by_dlog(int(g2), SEED_BASES, cref.synth_scalars(SEED_SCALARS, n_total, False)))
    ms = sum(times) / len(times)
    # configs[0] on the CPU beside it (cheap): G1 2^16, the benches/group.rs:18-26 shape
    n16 = 1 << 16
    b16 = cref.synth_bases(0, SEED_BASES + 16, n16)
    s16 = cref.synth_scalars(SEED_SCALARS + 16, n16, True)
    t16 = []
    for _ in range(4):
        t0 = time.perf_counter()
        cref.msm(0, b16, s16, 1, nthreads=cores)
        t16.append((time.perf_counter() - t0) * 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": ms, "unit": "ms", "n_gpus": args.gpus,
        "steps": len(times), "warmup": warmup, "ms_per_step": ms, "higher_is_better": False,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "points_per_s": n_total / (ms * 1e-3),
        "config": {"workload": f"{args.group.upper()} MSM 2^{args.logn} points per GPU x {args.gpus} GPU(s) = {n_total} points",
                   "window_rule": "blst pippenger_window_size + breakdown", "input_gen_s": round(gen_s, 2), "parity_vs_dlog": bool(ok),
                   "steps_requested": args.steps, "arm_wall_s": round(time.perf_counter() - t_arm, 1)},
        "cpu_baseline": {"value": ms, "unit": "ms", "cores": cores, "kind": "port",
                         "sample": f"the full workload: all {n_total} points per step, {len(times)} timed steps", "what": cpu_what(cref)},
        "c0_2p16": {"workload": "G1 MSM 2^16 (benches/group.rs:18-26 shape) on the host cores", "cpu_ms": min(t16[1:]), "cores": cores},
        "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
class Ctx:
    """per-process state of our arm: device, stream, ranks, engine"""

    def __init__(self, args):
        import torch

        import ark_blst_b200 as eng

        self.torch, self.eng = torch, eng
        self.rank, self.world, self.local = dist_env()
        if self.world != args.gpus and self.world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — the MSM engine has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist

            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
            # a CPU-side group for waits that must leave the GPUs idle (an NCCL barrier parks a spinning kernel on every
            # waiting rank's GPU, and kernels of another process then time-slice against it)
            self.cpu_group = dist.new_group(backend="gloo")
        self.L = eng._lib.lib
        eng._lib.check(self.L.b200msm_init(self.local, 1), "init")
        self.stream = torch.cuda.current_stream().cuda_stream
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)  # > 126 MB L2
        self.peak = eng.imad_peak()

    def barrier(self):
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def cpu_barrier(self):
        self.torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier(group=self.cpu_group)

    def allmax(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.dist:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.dist:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def allok(self, ok):
        t = self.torch.tensor([1 if ok else 0], dtype=self.torch.int32, device=self.dev)
        if self.dist:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item())


class Case:
    """One MSM over the ranks: every rank holds `n` synthetic points (seeded per rank) in HBM, computes
    its partial, the partials are all-gathered over NCCL and rank 0 adds them.  `n` per rank is the
    caller's choice: fixed per GPU (weak scaling) or total/world (strong scaling)."""

    def __init__(self, cx, g2, n, tag, canon=None):
        import numpy as np

        torch, eng = cx.torch, cx.eng
        self.cx, self.g2, self.n, self.np = cx, int(g2), n, np
        self.G = eng.G2 if g2 else eng.G1
        self.grp = eng.G2Projective if g2 else eng.G1Projective
        self.aw, self.jw = (24, 36) if g2 else (12, 18)
        self.seed_b = SEED_BASES + 7919 * tag + 1000003 * cx.rank
        self.seed_s = SEED_SCALARS + 7919 * tag + 1000003 * cx.rank
        self.bases = torch.empty((n, self.aw), dtype=torch.int64, device=cx.dev)
        self.scalars = torch.empty((n, 4), dtype=torch.int64, device=cx.dev)
        eng.synth_bases_device(self.G, self.seed_b, n, self.bases.data_ptr(), cx.stream)
        self.mont = canon is None
        self.canon = canon
        if canon is None:
            eng.synth_scalars_device(self.seed_s, n, True, self.scalars.data_ptr(), cx.stream)  # Montgomery: what `msm` receives
        else:
            self.scalars.copy_(torch.from_numpy(canon.view(np.int64)))
        self.partial = torch.zeros(self.jw, dtype=torch.int64, device=cx.dev)
        self.gathered = torch.zeros((cx.world, self.jw), dtype=torch.int64, device=cx.dev)
        self.result = torch.zeros(self.jw, dtype=torch.int64, device=cx.dev)
        self.table = None
        torch.cuda.synchronize()

    def combine(self):
        """the multi-GPU exchange of the path (ark_blst_b200/dist.py): one Jacobian partial per rank all-gathered over
        NCCL, final addition on rank 0 by k_sum_partials"""
        from ark_blst_b200 import dist as msm_dist

        cx = self.cx
        if cx.world > 1:
            msm_dist.gather_partials(self.partial, cx.world, cx.dist, out=self.gathered)
            if cx.rank == 0:
                msm_dist.combine_on_device(self.G, self.gathered, cx.stream, out=self.result)
        else:
            self.result.copy_(self.partial)

    def step_device(self, table=False):
        cx = self.cx
        if table:
            cx.eng.run_table_device(self.G, self.table["t"].data_ptr(), self.n, self.table["c"], self.scalars.data_ptr(), self.n, self.mont,
                                    self.partial.data_ptr(), cx.stream)
        else:
            cx.eng.run_device(self.G, self.bases.data_ptr(), self.scalars.data_ptr(), self.n, self.mont, self.partial.data_ptr(), cx.stream)
        self.combine()

    def time_device(self, steps, warmup, table=False):
        """CUDA events around every step on torch's current stream (the stream the MSM is launched on), L2 flushed
        between steps outside the events, max over ranks. Returns (ms, phases of this rank, launches per step)."""
        cx, torch = self.cx, self.cx.torch
        for _ in range(warmup):
            cx.flush.zero_()
            self.step_device(table)
        cx.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        launches0 = cx.L.b200msm_launch_count()
        phases = []
        for k in range(steps):
            cx.flush.zero_()                   # L2 flush between steps, outside the per-step events
            ev[k][0].record()
            self.step_device(table)
            ev[k][1].record()
            ev[k][1].synchronize()
            phases.append(cx.eng.last_phase_ms())
        cx.barrier()
        launches = (cx.L.b200msm_launch_count() - launches0) / steps
        ms = cx.allmax(sum(a.elapsed_time(b) for a, b in ev) / steps)
        ph = {k: sum(p[k] for p in phases) / len(phases) for k in phases[0] if k != "valid"}
        return ms, ph, launches

    def expected_partial(self):
        from oracle import cref

        sc = self.canon if self.canon is not None else cref.synth_scalars(self.seed_s, self.n, False)
        return cref.msm_by_dlog(self.g2, self.seed_b, sc)

    def check(self, extra_totals=()):
        """this rank's partial against the known-discrete-log closed form; on rank 0 the combined result against the
        sum of all ranks' expected partials, and every array in `extra_totals` against the combined result"""
        from oracle import cref

        cx, np, torch = self.cx, self.np, self.cx.torch
        exp = self.expected_partial()
        ok = cref.affine_equal(self.g2, self.partial.cpu().numpy().view(np.uint64), exp)
        if cx.world > 1:
            exp_all = [torch.zeros(self.jw, dtype=torch.int64, device=cx.dev) for _ in range(cx.world)]
            cx.dist.all_gather(exp_all, torch.from_numpy(exp.view(np.int64)).to(cx.dev))
        if cx.rank == 0:
            total = self.result.cpu().numpy().view(np.uint64)
            exp_total = exp
            if cx.world > 1:
                exp_total = np.zeros(self.jw, dtype=np.uint64)
                for e in exp_all:
                    exp_total = cref.add(self.g2, exp_total, e.cpu().numpy().view(np.uint64))
            ok = ok and cref.affine_equal(self.g2, total, exp_total)
            for x in extra_totals:
                ok = ok and x is not None and cref.affine_equal(self.g2, x, exp_total)
        return cx.allok(ok)

    # ---- end to end: the reference-facing call with HOST buffers, H2D + D2H inside ----
    def host_buffers(self, pinned):
        torch, np = self.cx.torch, self.np
        if pinned:
            hb = torch.empty((self.n, self.aw), dtype=torch.int64, pin_memory=True)
            hs = torch.empty((self.n, 4), dtype=torch.int64, pin_memory=True)
            hb.copy_(self.bases); hs.copy_(self.scalars)
            torch.cuda.synchronize()
            return hb, hs, hb.numpy().view(np.uint64), hs.numpy().view(np.uint64)
        hb_np = self.bases.cpu().numpy().view(np.uint64).copy()      # ordinary pageable memory, as a Rust Vec is
        hs_np = self.scalars.cpu().numpy().view(np.uint64).copy()
        return None, None, hb_np, hs_np

    def time_e2e(self, steps, warmup, call):
        """`call()` → this rank's partial as host limbs (through the host-buffer C-ABI); partials → device,
        NCCL all-gather, final addition on rank 0, result read back. Wall clock bracketed by barriers, max over ranks."""
        from ark_blst_b200 import dist as msm_dist

        cx, torch, np = self.cx, self.cx.torch, self.np
        hpart = torch.zeros(self.jw, dtype=torch.int64, pin_memory=True)

        def step():
            out = call()
            if cx.world > 1:
                hpart.numpy().view(np.uint64)[:] = out
                self.partial.copy_(hpart, non_blocking=True)
                msm_dist.gather_partials(self.partial, cx.world, cx.dist, out=self.gathered)
                if cx.rank == 0:
                    msm_dist.combine_on_device(self.G, self.gathered, cx.stream, out=self.result)
                    return self.result.cpu().numpy().view(np.uint64)
                torch.cuda.synchronize()
            return out

        for _ in range(max(1, warmup)):
            out = step()
        cx.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            out = step()
        cx.barrier()
        return cx.allmax((time.perf_counter() - t0) * 1e3 / steps), (out.copy() if cx.rank == 0 else None)


def side_steps(args):
    return min(args.steps, 5), 3


def block_msm(cx, args, g2, n_total, tag, name):
    """A BASELINE config as a strong-scaling block: an n_total-point MSM sharded evenly over the ranks."""
    from ark_blst_b200.dist import shard_range

    steps, warmup = side_steps(args)
    lo, hi = shard_range(n_total, cx.rank, cx.world)
    case = Case(cx, g2, hi - lo, tag)
    cx.L.b200msm_set_profiling(1)
    ms, ph, _ = case.time_device(steps, warmup)
    cx.L.b200msm_set_profiling(0)
    plan = cx.eng.last_plan()
    total_dev = case.result.cpu().numpy().view(case.np.uint64).copy() if cx.rank == 0 else None
    _, _, hb_np, hs_np = keep = case.host_buffers(True)
    e2e_ms, e2e_out = case.time_e2e(steps, warmup, lambda: case.grp.msm(hb_np, hs_np))
    case.step_device()
    cx.barrier()
    ok = case.check((total_dev, e2e_out))
    del keep
    c, W, fp_total, _ = work_model(n_total, g2)
    out = {"workload": f"{'G2' if g2 else 'G1'} MSM 2^{int(math.log2(n_total))} points sharded over {cx.world} GPU(s) ({hi - lo} per GPU)",
           "scaling": "strong", "ms": ms, "points_per_s": n_total / (ms * 1e-3), "e2e_ms": e2e_ms,
           "e2e_h2d_bytes_per_step": n_total * (case.aw * 8 + 32), "parity_ok": ok, "steps": steps, "warmup": warmup,
           "engine_plan_per_gpu": plan, "phases_ms_rank0": {k: round(v, 4) for k, v in ph.items()},
           "whole_msm_frac_of_imad_peak": fp_total * FPMUL_IMAD / (ms * 1e-3) / (cx.peak["imad_per_s"] * cx.world),
           "canonical_window_bits": c, "canonical_windows": W, "algorithmic_imad": fp_total * FPMUL_IMAD}
    del case
    cx.torch.cuda.empty_cache()
    return out


def witness_scalars(seed, n):
    """≈40 % zeros, ≈20 % ones, ≈10 % below 2^32, rest uniform (canonical form)"""
    import numpy as np

    from oracle import cref

    rng = np.random.default_rng(seed & 0xffffffff)
    u = rng.random(n)
    canon = cref.synth_scalars(seed, n, False)
    small = rng.integers(0, 1 << 32, size=n, dtype=np.uint64)
    canon[u < 0.7] = 0
    canon[(u >= 0.4) & (u < 0.6), 0] = 1
    mid = (u >= 0.6) & (u < 0.7)
    canon[mid, 0] = small[mid]
    return canon


def block_groth16(cx, args, logn, scalars_kind, lanes=4, table=False, steps=None, warmup=3):
    """BASELINE.json configs[4]: a Groth16-prover-shaped batch — three G1 MSMs and one G2 MSM of
    2^logn points each, issued back to back on `lanes` engine lanes, every MSM sharded over the ranks."""
    torch, eng, L = cx.torch, cx.eng, cx.L
    from oracle import cref
    import numpy as np

    if steps is None:
        steps, warmup = side_steps(args)
    from ark_blst_b200.dist import shard_range

    n_total = 1 << logn
    lo, hi = shard_range(n_total, cx.rank, cx.world)
    n = hi - lo
    cases, nonzero = [], 0.0
    for k, g2 in enumerate((1, 0, 0, 0)):  # G2 first: the tail left exposed at the end is a G1 one
        canon = None
        if scalars_kind == "witness":
            canon = witness_scalars(SEED_SCALARS + 7919 * (40 + k) + 1000003 * cx.rank, n)
            nonzero_k = cx.allsum(count_nonzero_booth_digits(canon, work_model(n_total, g2)[0]))
        else:
            c, W, _, _ = work_model(n_total, g2)
            nonzero_k = n_total * W * (1 - 2.0 ** -c)
        case = Case(cx, g2, n, 40 + k, canon)
        if table:  # resident proving key: fixed-base window table per MSM, built once outside the timed region
            tc, tw = eng.table_plan(g2, n)
            t = torch.empty((tw, n, case.aw), dtype=torch.int64, device=cx.dev)
            t[0].copy_(case.bases)
            eng.table_build_device(g2, t.data_ptr(), n, tc, t.data_ptr(), cx.stream)
            case.table = {"c": tc, "t": t}
        cases.append((case, nonzero_k))
    torch.cuda.synchronize()
    lanes = max(1, min(lanes, len(cases)))
    lane_streams = [torch.cuda.Stream(device=cx.dev) for _ in range(lanes)] if lanes > 1 else []

    def step():
        # the four MSMs go out on `lanes` engine lanes / streams (b200msm_set_lane): the latency-bound tail of one
        # overlaps the accumulation of the next; the partials are gathered once all four are in
        cur = torch.cuda.current_stream()
        for k, (case, _) in enumerate(cases):
            st = cx.stream
            if lanes > 1:
                L.b200msm_set_lane(k % lanes)
                lane_streams[k % lanes].wait_stream(cur)
                st = lane_streams[k % lanes].cuda_stream
            if case.table:
                eng.run_table_device(case.G, case.table["t"].data_ptr(), n, case.table["c"], case.scalars.data_ptr(), n, case.mont,
                                     case.partial.data_ptr(), st)
            else:
                eng.run_device(case.G, case.bases.data_ptr(), case.scalars.data_ptr(), n, case.mont, case.partial.data_ptr(), st)
        if lanes > 1:
            L.b200msm_set_lane(0)
            for ls in lane_streams:
                cur.wait_stream(ls)
        for case, _ in cases:
            case.combine()

    for _ in range(warmup):
        step()
    cx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    e1.synchronize()
    ms = cx.allmax(e0.elapsed_time(e1) / steps)
    ok = True
    for case, _ in cases:
        ok = case.check() and ok
    imad_uniform = (3 * work_model(n_total, False)[2] + work_model(n_total, True)[2]) * FPMUL_IMAD
    imad = sum(work_model_counted(nz, n_total, case.g2) for case, nz in cases) * FPMUL_IMAD
    peak = cx.peak["imad_per_s"] * cx.world
    out = {"workload": f"3xG1 + 1xG2 MSM at 2^{logn} points each, sharded over {cx.world} GPU(s), {scalars_kind} scalars",
           "ms_per_batch": ms, "steps": steps, "warmup": warmup, "lanes": lanes, "parity_ok": ok,
           "bases": "resident fixed-base window tables (built once)" if table else "resident affine bases",
           "algorithmic_imad": imad,
           "numerator": "SURVEY 8d work model at the canonical c*, accumulate term = the non-zero Booth digits of THESE scalars (counted)"
           if scalars_kind == "witness" else "SURVEY 8d work model at the canonical c*, uniform scalars",
           "frac_of_imad_peak": imad / (ms * 1e-3) / peak,
           "uniform_model_imad": imad_uniform}
    del cases
    torch.cuda.empty_cache()
    return out


def block_single_process(args, n_gpus):
    """The path a Rust `msm` takes on an 8-GPU box: ONE process, b200msm_init(0, N), host buffers in,
    the library shards over the N GPUs (one host worker thread per device, partials to device 0 by peer
    copy).  Runs as a child process of rank 0 while the other ranks wait at a barrier (their GPUs idle)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--workload", "single_process", "--gpus", str(n_gpus),
           "--steps", str(min(args.steps, 5)), "--warmup", "3"]
    drop = ("RANK", "WORLD_SIZE", "LOCAL_RANK", "LOCAL_WORLD_SIZE", "GROUP_RANK", "ROLE_RANK", "MASTER_ADDR", "MASTER_PORT",
            "TORCHELASTIC_RUN_ID")
    env = {k: v for k, v in os.environ.items() if k not in drop}
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
        for line in reversed(r.stdout.strip().splitlines()):
            if line.startswith("{"):
                return json.loads(line)
        return {"error": (r.stderr or r.stdout)[-400:]}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)}


def run_single_process(args):
    """child of block_single_process (also usable alone: --workload single_process --gpus N)"""
    import numpy as np
    import torch

    import ark_blst_b200 as eng
    from oracle import cref

    L = eng._lib.lib
    N = args.gpus
    eng._lib.check(L.b200msm_init(0, N), "init")
    out = {"api": f"one process: b200msm_init(0, {N}) then b200msm_g1 / b200msm_run(host buffers); sharding, peer gather and final addition inside the library",
           "devices_bound": L.b200msm_device_count()}
    torch.cuda.set_device(0)
    for name, g2, logn in (("g1_2p24", 0, args.logn or 24), ("g2_2p20", 1, 20)):
        n = 1 << logn
        aw = 24 if g2 else 12
        grp = eng.G2Projective if g2 else eng.G1Projective
        db = torch.empty((n, aw), dtype=torch.int64, device="cuda:0")
        ds = torch.empty((n, 4), dtype=torch.int64, device="cuda:0")
        eng.synth_bases_device(g2, SEED_BASES + 99, n, db.data_ptr(), 0)
        eng.synth_scalars_device(SEED_SCALARS + 99, n, True, ds.data_ptr(), 0)
        torch.cuda.synchronize()
        hb = torch.empty((n, aw), dtype=torch.int64, pin_memory=True); hb.copy_(db)
        hs = torch.empty((n, 4), dtype=torch.int64, pin_memory=True); hs.copy_(ds)
        torch.cuda.synchronize()
        del db, ds
        torch.cuda.empty_cache()
        exp = cref.msm_by_dlog(g2, SEED_BASES + 99, cref.synth_scalars(SEED_SCALARS + 99, n, False))
        hb_np, hs_np = hb.numpy().view(np.uint64), hs.numpy().view(np.uint64)

        def timeit(call):
            for _ in range(args.warmup):
                r = call()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                r = call()
            return (time.perf_counter() - t0) * 1e3 / args.steps, bool(cref.affine_equal(g2, r, exp))

        res = {"points": n}
        L.b200msm_set_graphs(0)
        res["e2e_pinned_ms_kernel_by_kernel"], ok0 = timeit(lambda: grp.msm(hb_np, hs_np))
        L.b200msm_set_graphs(1)
        res["e2e_pinned_ms"], ok1 = timeit(lambda: grp.msm(hb_np, hs_np))
        ok1 = ok1 and ok0
        pb, ps = hb_np.copy(), hs_np.copy()
        res["e2e_pageable_ms"], ok2 = timeit(lambda: grp.msm(pb, ps))
        del pb
        rb = eng.ResidentBases(grp, hb_np)
        res["resident_bases_e2e_ms"], ok3 = timeit(lambda: rb.msm(hs_np))
        res["resident_bases_pageable_scalars_ms"], ok3b = timeit(lambda: rb.msm(ps))
        ok4 = True
        if not g2:
            rb.precompute()
            res["resident_table_e2e_ms"], ok4 = timeit(lambda: rb.msm(hs_np))
        rb.close()
        res["parity_ok"] = ok1 and ok2 and ok3 and ok3b and ok4
        res["h2d_bytes_per_step"] = n * (aw * 8 + 32)
        out[name] = res
        del hb, hs
    L.b200msm_shutdown()
    print(json.dumps(out))
    return 0


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np

    cx = Ctx(args)
    torch, eng, L = cx.torch, cx.eng, cx.L
    rank, world = cx.rank, cx.world
    blocks = set(ALL_BLOCKS if args.blocks == "all" else [b for b in args.blocks.split(",") if b and b != "headline"])

    g2 = args.group == "g2"
    n = 1 << args.logn
    n_total = n * world
    L.b200msm_set_profiling(1)
    head = Case(cx, g2, n, 0)
    aw, jw = head.aw, head.jw

    sampler = ClockSampler(cx.local)
    if rank == 0:
        sampler.start()
    cx.barrier()
    t_wall0 = time.perf_counter()
    dev_ms, ph, launches = head.time_device(args.steps, args.warmup)
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    plan = eng.last_plan()
    total_dev = head.result.cpu().numpy().view(np.uint64).copy() if rank == 0 else None
    my_partial_ok_later = head.partial.clone()

    # ---- e2e: the reference-facing call with HOST buffers, H2D + D2H inside ----
    hb, hs, hb_np, hs_np = head.host_buffers(True)
    e2e_ms, e2e_out = head.time_e2e(args.steps, args.warmup, lambda: head.grp.msm(hb_np, hs_np))
    clocks = sampler.stop() if rank == 0 else None
    _, _, pb_np, ps_np = head.host_buffers(False)
    e2e_pg_ms, e2e_pg_out = head.time_e2e(min(args.steps, 10), 3, lambda: head.grp.msm(pb_np, ps_np))
    del pb_np

    # ---- resident bases with a fixed-base window table (b200msm_bases_precompute): the prover
    # shape — the proving key's bases stay on the device, only scalars arrive per MSM.  Reported
    # beside the headline, never instead of it: `value` and `e2e` above take fresh bases per call.
    table_info, extra = None, [total_dev, e2e_out, e2e_pg_out]
    if not args.no_table:
        c_t, W_t = eng.table_plan(head.G, n)
        t = torch.empty((W_t, n, aw), dtype=torch.int64, device=cx.dev)
        t[0].copy_(head.bases)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.table_build_device(head.G, t.data_ptr(), n, c_t, t.data_ptr(), cx.stream)
        torch.cuda.synchronize()
        build_ms = (time.perf_counter() - t0) * 1e3
        head.table = {"c": c_t, "t": t}
        tbl_ms, tph, _ = head.time_device(args.steps, args.warmup, table=True)
        extra.append(head.result.cpu().numpy().view(np.uint64).copy() if rank == 0 else None)
        # end to end: b200msm_bases_upload + b200msm_bases_precompute once, then b200msm_run(host scalars) per MSM
        rb = eng.ResidentBases(head.grp, hb_np)
        res_e2e_ms, res_out = head.time_e2e(args.steps, args.warmup, lambda: rb.msm(hs_np, montgomery=True))
        res_pg_ms, res_pg_out = head.time_e2e(min(args.steps, 10), 3, lambda: rb.msm(ps_np, montgomery=True))
        rb.precompute(c_t)
        tbl_e2e_ms, tbl_out = head.time_e2e(args.steps, args.warmup, lambda: rb.msm(hs_np, montgomery=True))
        rb.close()
        extra += [res_out, res_pg_out, tbl_out]
        table_info = {"ms_per_msm": tbl_ms, "points_per_s": n_total / (tbl_ms * 1e-3), "window_bits": c_t, "windows": W_t,
                      "table_bytes_per_gpu": W_t * n * aw * 8, "build_ms_once": build_ms,
                      "phases_ms": {k: round(v, 4) for k, v in tph.items()},
                      "e2e_ms": tbl_e2e_ms, "e2e_h2d_bytes_per_step": n * 32 * world,
                      "e2e_ms_resident_without_table": res_e2e_ms, "e2e_ms_resident_without_table_pageable_scalars": res_pg_ms,
                      "api": "b200msm_bases_upload + b200msm_bases_precompute once; per MSM b200msm_run(handle, host scalars) "
                             "(device figure: b200msm_run_table_device)"}
        head.table = None
        del t
        torch.cuda.empty_cache()

    # ---- parity of what was timed (the oracle is the checker only) ----
    from oracle import cref

    head.step_device()
    cx.barrier()
    parity = head.check(extra)
    parity = parity and cx.allok(cref.affine_equal(int(g2), my_partial_ok_later.cpu().numpy().view(np.uint64), head.expected_partial()))

    # ---- CPU baseline: rank 0 at N=1 only, the full headline workload on every host core ----
    cpu = None
    c0_cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = cref.ncores()
        best = None
        reps = 2 if args.logn <= 20 else 1
        for _ in range(reps):
            t0 = time.perf_counter()
            cpu_out = cref.msm(int(g2), hb_np, hs_np, 1, nthreads=cores)
            dt = (time.perf_counter() - t0) * 1e3
            best = dt if best is None else min(best, dt)
        parity = parity and cref.affine_equal(int(g2), cpu_out, total_dev)
        cpu = {"value": best, "unit": "ms", "cores": cores, "kind": "port",
               "sample": f"the full 2^{args.logn}-point workload, best of {reps} runs, same inputs as the GPU step",
               "what": cpu_what(cref)}
    del hb, hs
    extra_blocks = {}
    L.b200msm_set_profiling(0)

    # ---- the other BASELINE configs, each a block of the same line ----
    if "c0_2p16" in blocks:
        steps, warmup = max(args.steps, 10), 3
        case = Case(cx, 0, 1 << 16, 16)
        # configs[0] is a single-GPU shape: every rank runs its own copy, rank 0's time is reported
        world_save, cx.world, dist_save, cx.dist = cx.world, 1, cx.dist, None
        L.b200msm_set_profiling(1)
        ms16, ph16, _ = case.time_device(steps, warmup)
        L.b200msm_set_profiling(0)
        plan16 = eng.last_plan()
        _, _, b16, s16 = keep = case.host_buffers(True)
        e2e16, out16 = case.time_e2e(steps, warmup, lambda: case.grp.msm(b16, s16))
        case.step_device()
        ok16 = case.check((out16,))
        blk = {"workload": "G1 MSM 2^16 random points/scalars (benches/group.rs:18-26 shape), one GPU", "ms": ms16, "e2e_ms": e2e16,
               "parity_ok": ok16, "engine_plan": plan16, "phases_ms": {k: round(v, 4) for k, v in ph16.items()},
               "whole_msm_frac_of_imad_peak": work_model(1 << 16, False)[2] * FPMUL_IMAD / (ms16 * 1e-3) / cx.peak["imad_per_s"]}
        if rank == 0 and world_save == 1 and not args.no_cpu:
            cores = cref.ncores()
            ts = []
            for _ in range(4):
                t0 = time.perf_counter()
                cpu16 = cref.msm(0, b16, s16, 1, nthreads=cores)
                ts.append((time.perf_counter() - t0) * 1e3)
            blk["cpu_ms"] = min(ts[1:])
            blk["cpu_cores"] = cores
            blk["cpu_what"] = "the same C port as cpu_baseline, full 2^16-point workload, best of 3 after one warm-up"
            blk["parity_ok"] = blk["parity_ok"] and bool(cref.affine_equal(0, cpu16, out16))
        cx.world, cx.dist = world_save, dist_save
        blk["parity_ok"] = cx.allok(blk["parity_ok"])
        extra_blocks["c0_2p16"] = blk
        del case, keep
        torch.cuda.empty_cache()
    if "g2_2p20" in blocks:
        extra_blocks["g2_2p20"] = block_msm(cx, args, True, 1 << 20, 20, "g2_2p20")
    if "strong_2p24" in blocks:
        extra_blocks["strong_2p24"] = block_msm(cx, args, False, 1 << 24, 24, "strong_2p24")
    if "groth16_2p22" in blocks:
        extra_blocks["groth16_2p22"] = {"uniform": block_groth16(cx, args, 22, "uniform"),
                                        "witness_like": block_groth16(cx, args, 22, "witness")}
    if "single_process" in blocks and world > 1:
        # the other ranks wait on the CPU (gloo), GPUs idle: the child process drives all N GPUs itself
        cx.barrier()
        cx.cpu_barrier()
        sp = block_single_process(args, world) if rank == 0 else None
        cx.cpu_barrier()
        if rank == 0:
            extra_blocks["single_process"] = sp

    if rank == 0:
        for b in extra_blocks.values():
            for v in (b.values() if "parity_ok" not in b else [b]):
                if isinstance(v, dict) and "parity_ok" in v:
                    parity = parity and bool(v["parity_ok"])
        c, W, _, fpmul_acc = work_model(n, g2)            # what each GPU's accumulate kernel runs
        _, _, fpmul_total, _ = work_model(n_total, g2)     # single-problem numerator
        acc_s = ph["accumulate"] * 1e-3
        imad_peak = cx.peak["imad_per_s"]
        entries_per_bucket = n * (1 if plan["glv"] == 0 else (2 if plan["glv"] <= 2 else 4)) / 2.0 ** (plan["window_bits"] - 1)
        ba_rounds = 3 if entries_per_bucket >= 24 else (1 if entries_per_bucket >= 12 else 0)   # csrc/plan.h ba_rounds_for
        exec_fpmul = executed_accumulate_fpmul(n, g2, plan, ba_rounds)
        achieved = fpmul_acc * FPMUL_IMAD / acc_s
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                tj = json.load(open(tpath))
                traffic = tj.get(f"k_accumulate_{args.group}_2^{args.logn}")
                traffic_src = tj.get("source")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": dev_ms, "unit": "ms", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": False, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32x12 Montgomery limbs (integer)", "data": "synthetic",
            "points_per_s": n_total / (dev_ms * 1e-3),
            "config": {"workload": f"{args.group.upper()} MSM 2^{args.logn} points per GPU x {world} GPU(s) = {n_total} points",
                       "window_bits": c, "windows": W,
                       "engine_plan": {**plan, "note": "glv 2 = scalars split k1 + k2*lambda over (P, phi(P)), unsigned top digit (4: four parts over the psi images, G2); "
                                                       "window_bits/windows above are the canonical c* of the work model (roofline numerator)"},
                       "scalars": "uniform mod r, Montgomery form (VariableBaseMSM::msm)",
                       "bases": "random subgroup points k_i*G, affine, resident in HBM",
                       "l2": "flushed between steps (256 MiB memset, outside the per-step events)",
                       "parity": "every timed result == (sum s_i k_i)*G (known-discrete-log closed form) and == the CPU port's result" if parity else "MISMATCH"},
            "parity_ok": parity,
            "wall_ms_per_step_incl_flush": wall_ms / (args.steps + args.warmup),
            "phases_ms": {k: round(v, 4) for k, v in ph.items()},
            "roofline": {
                "bound": "imad", "kernel": f"bucket accumulation: k_accumulate<{'fp2' if g2 else 'fp'}> (+ heavy-bucket and batched-affine kernels of the same phase)",
                "achieved": achieved / 1e12, "peak": imad_peak / 1e12, "unit": "TIMAD/s",
                "frac": achieved / imad_peak, "traffic": traffic,
                "traffic_source": traffic_src or "profiles/traffic.json: dram bytes of one ncu --set full capture of this kernel (a recorded constant, not measured in this run)",
                "peak_source": "measured live by b200msm_imad_peak (mad.lo.u32 issue rate; MEASURED_PEAKS.json has no integer figure)",
                "algorithmic_imad_per_launch": fpmul_acc * FPMUL_IMAD,
                "kernel_ms": ph["accumulate"],
                "whole_msm": {"algorithmic_imad": fpmul_total * FPMUL_IMAD,
                              "frac": fpmul_total * FPMUL_IMAD / (dev_ms * 1e-3) / (imad_peak * world)},
                "executed": {"ba_rounds": ba_rounds, "imad_per_launch": exec_fpmul * FPMUL_IMAD, "frac": exec_fpmul * FPMUL_IMAD / acc_s / imad_peak,
                             "note": "IMAD the phase really executes under the engine's plan: with batched-affine pairing rounds an addition costs 6 "
                                     "products instead of the 10 the SURVEY 8d numerator counts, so `frac` (fixed 8d numerator = algorithmic work of the "
                                     "XYZZ formulation per second, over the pipe peak) can exceed 1 while the pipe utilisation is this figure"},
                "gather_gbs": n * W * (192 if g2 else 96) / acc_s / 1e9,
                "hbm_peak_gbs": _measured_hbm(),
            },
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": n * (aw * 8 + 32) * world,
                    "d2h_bytes_per_step": jw * 8 * world, "points_per_s": n_total / (e2e_ms * 1e-3),
                    "api": "b200msm_g1/b200msm_g2(host bases, host scalars) via ark_blst_b200.G?Projective.msm, pinned host memory"},
            "e2e_pageable": {"value": e2e_pg_ms, "unit": "ms", "steps": min(args.steps, 10),
                             "api": "the same call from ordinary pageable host memory (numpy arrays; what a Rust Vec is)"},
            "resident_table": table_info,
            **extra_blocks,
            "gpu_launches": launches,
            "clocks": clocks,
            "imad_peak": cx.peak,
        }
        if table_info is not None:  # same numerator as the headline (the plain MSM's algorithmic work at c*)
            table_info["frac_of_plain_msm_imad"] = fpmul_total * FPMUL_IMAD / (table_info["ms_per_msm"] * 1e-3) / (imad_peak * world)
        print(json.dumps(line))
    if world > 1:
        cx.dist.barrier()
        cx.dist.destroy_process_group()
    return 0


def run_groth16(args):
    """`--workload groth16`: the configs[4] block alone, with its knobs (--logn, --scalars, --lanes, --table)."""
    cx = Ctx(args)
    blk = block_groth16(cx, args, args.logn, args.scalars, lanes=args.lanes, table=args.table, steps=args.steps, warmup=args.warmup)
    if cx.rank == 0:
        print(json.dumps({
            "metric": "Groth16-shaped batch: 3xG1 + 1xG2 MSM, ms per batch", "value": blk["ms_per_batch"], "unit": "ms", "n_gpus": cx.world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": blk["ms_per_batch"], "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32x12 Montgomery limbs (integer)", "data": "synthetic",
            "config": {"workload": blk["workload"], "lanes": blk["lanes"], "bases": blk["bases"]},
            "parity_ok": blk["parity_ok"],
            "roofline": {"bound": "imad", "achieved": blk["algorithmic_imad"] / (blk["ms_per_batch"] * 1e-3) / 1e12,
                         "peak": cx.peak["imad_per_s"] * cx.world / 1e12, "unit": "TIMAD/s",
                         "frac": blk["frac_of_imad_peak"], "traffic": None, "algorithmic_imad": blk["algorithmic_imad"],
                         "numerator": blk["numerator"]}}))
    if cx.world > 1:
        cx.dist.barrier()
        cx.dist.destroy_process_group()
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
    ap.add_argument("--workload", default="msm", choices=["msm", "groth16", "single_process"])
    ap.add_argument("--blocks", default="all", help="comma list of the extra per-config blocks to run beside the headline "
                                                    f"({', '.join(ALL_BLOCKS)}), 'all' or 'headline' (none)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--scalars", default="uniform", choices=["uniform", "witness"], help="groth16 workload: scalar distribution")
    ap.add_argument("--lanes", type=int, default=4, help="groth16 workload: engine lanes / streams the four MSMs are spread over (1 = back to back on one stream)")
    ap.add_argument("--table", action="store_true", help="groth16 workload: run every MSM against a resident fixed-base window table")
    ap.add_argument("--no-table", action="store_true", help="skip the resident-bases fixed-base-table leg")
    args = ap.parse_args()
    if args.workload == "single_process":
        if args.warmup < 3:
            args.warmup = 3
        return run_single_process(args)
    if args.logn is None:
        args.logn = 22 if args.workload == "groth16" else 20
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.workload == "groth16" and args.impl == "b200":
        return run_groth16(args)
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
