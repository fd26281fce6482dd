"""BASELINE.json's full sizes, through size-independent properties (no EC oracle at this n):
known-discrete-log closed form, shard-sum invariance and linearity. Inputs live in HBM (torch is
used only as the device allocator)."""
import numpy as np
import pytest

from oracle import bls12381 as o

pytestmark = pytest.mark.gpu


def _dev_inputs(eng, torch, g2, seed_b, seed_s, n, mont):
    aw = 24 if g2 else 12
    bases = torch.empty((n, aw), dtype=torch.int64, device="cuda")
    scalars = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    eng.synth_bases_device(g2, seed_b, n, bases.data_ptr())
    eng.synth_scalars_device(seed_s, n, mont, scalars.data_ptr())
    torch.cuda.synchronize()
    return bases, scalars


def _run(eng, torch, g2, bases, scalars, n, mont):
    out = torch.zeros(36 if g2 else 18, dtype=torch.int64, device="cuda")
    eng.run_device(g2, bases.data_ptr(), scalars.data_ptr(), n, mont, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return out.cpu().numpy().view(np.uint64)


def test_synth_matches_oracle(eng, cref):
    import torch

    for g2 in (0, 1):
        n = 96
        bases, scalars = _dev_inputs(eng, torch, g2, 123, 456, n, True)
        assert np.array_equal(bases.cpu().numpy().view(np.uint64), cref.synth_bases(g2, 123, n))
        assert np.array_equal(scalars.cpu().numpy().view(np.uint64), cref.synth_scalars(456, n, True))
        _, sc = _dev_inputs(eng, torch, g2, 123, 456, n, False)
        assert np.array_equal(sc.cpu().numpy().view(np.uint64), cref.synth_scalars(456, n, False))


@pytest.mark.parametrize("g2,logn", [(0, 16), (0, 20), (0, 22), (0, 23), (0, 24), (1, 18), (1, 20), (1, 21), (1, 22)])
def test_dlog_closed_form(eng, cref, g2, logn):
    """Σ sᵢ·(kᵢ·G) = (Σ sᵢkᵢ mod r)·G at BASELINE sizes"""
    import torch

    n = 1 << logn
    sb, ss = 0xB2000381_00000000 + logn, 77 + logn
    bases, scalars = _dev_inputs(eng, torch, g2, sb, ss, n, True)
    got = _run(eng, torch, g2, bases, scalars, n, True)
    exp = cref.msm_by_dlog(g2, sb, cref.synth_scalars(ss, n, False))
    assert cref.affine_equal(g2, got, exp)
    if logn >= 22:   # the largest sizes also through the canonical-BigInt entry (msm_bigint) and with GLV forced the other way
        _, canon = _dev_inputs(eng, torch, g2, sb, ss, n, False)
        L = eng._lib.lib
        try:
            assert L.b200msm_set_glv((1 if logn > 22 else 0) if not g2 else 2) == 0   # (G2: four parts at full size too)
            assert cref.affine_equal(g2, _run(eng, torch, g2, bases, canon, n, False), exp)
        finally:
            L.b200msm_set_glv(-1)
    del bases, scalars
    torch.cuda.empty_cache()


def test_shard_sum_invariance_and_linearity(eng, cref):
    """1 shard vs 2/4/8 shards combined by the final-addition kernel (the multi-GPU combine, on
    one GPU); and msm(P,s) + msm(P,t) = msm(P,s+t)."""
    import torch

    n = 1 << 18
    bases, scalars = _dev_inputs(eng, torch, 0, 9001, 9002, n, False)
    whole = _run(eng, torch, 0, bases, scalars, n, False)
    for k in (2, 4, 8):
        parts = torch.zeros((k, 18), dtype=torch.int64, device="cuda")
        for j in range(k):
            lo, hi = n * j // k, n * (j + 1) // k
            eng.run_device(0, bases[lo:].data_ptr(), scalars[lo:].data_ptr(), hi - lo, False, parts[j].data_ptr(),
                           torch.cuda.current_stream().cuda_stream)
        out = torch.zeros(18, dtype=torch.int64, device="cuda")
        eng.sum_partials_device(0, parts.data_ptr(), k, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert cref.affine_equal(0, out.cpu().numpy().view(np.uint64), whole)
    # linearity with t = second stream
    _, t = _dev_inputs(eng, torch, 0, 9001, 9003, n, False)
    rt = _run(eng, torch, 0, bases, t, n, False)
    s_h = cref.synth_scalars(9002, n, False)
    t_h = cref.synth_scalars(9003, n, False)
    exp = cref.add(0, cref.msm_by_dlog(0, 9001, s_h), cref.msm_by_dlog(0, 9001, t_h))
    assert cref.affine_equal(0, cref.add(0, whole, rt), exp)


def test_lanes_overlap_on_streams(eng, cref):
    """b200msm_set_lane: MSMs issued on different lanes and streams (a Groth16-shaped batch: G2 + 3×G1)
    run concurrently on separate scratch arenas; every result equals the dlog closed form and the
    lane-0, one-stream result"""
    import torch

    L = eng._lib.lib
    n = 1 << 16
    jobs = []
    for k, g2 in enumerate((1, 0, 0, 0)):
        bases, scalars = _dev_inputs(eng, torch, g2, 7000 + k, 7100 + k, n, True)
        jobs.append((g2, bases, scalars, 7000 + k, 7100 + k))
    serial = [_run(eng, torch, g2, b, s, n, True) for g2, b, s, _, _ in jobs]
    streams = [torch.cuda.Stream() for _ in range(2)]
    cur = torch.cuda.current_stream()
    for rep in range(3):
        outs = [torch.zeros(36 if g2 else 18, dtype=torch.int64, device="cuda") for g2, *_ in jobs]
        try:
            for k, (g2, b, s, _, _) in enumerate(jobs):
                assert L.b200msm_set_lane(k % 2) == 0
                streams[k % 2].wait_stream(cur)
                eng.run_device(g2, b.data_ptr(), s.data_ptr(), n, True, outs[k].data_ptr(), streams[k % 2].cuda_stream)
        finally:
            L.b200msm_set_lane(0)
        for st in streams:
            cur.wait_stream(st)
        torch.cuda.synchronize()
        for k, (g2, b, s, sb, ss) in enumerate(jobs):
            got = outs[k].cpu().numpy().view(np.uint64)
            assert cref.affine_equal(g2, got, serial[k]), (rep, k)
            assert cref.affine_equal(g2, got, cref.msm_by_dlog(g2, sb, cref.synth_scalars(ss, n, False))), (rep, k)

