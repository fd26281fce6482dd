"""Batched-affine pairing rounds (csrc/batch_affine.cuh) and CUDA-graph replay (run_pass): every forced
configuration must give the same group element as the oracle, incl. the exceptional cases INSIDE a
batch — P + P (doubling), P + (−P) (cancellation), identity bases, padding slots."""
import random

import numpy as np
import pytest

from helpers import curve, jac_to_point, points_to_limbs, rand_fp2_limbs, rand_fp_limbs, scalars_to_limbs
from oracle import bls12381 as o

pytestmark = pytest.mark.gpu


def _grp(eng, g2):
    return eng.G2Projective if g2 else eng.G1Projective


@pytest.fixture
def knobs(eng):
    L = eng._lib.lib
    yield L
    L.b200msm_set_batch_affine(-1)
    L.b200msm_set_graphs(1)
    L.b200msm_set_glv(-1)
    L.b200msm_set_window_bits(0)
    L.b200msm_set_stream_slices(8, 0)
    L.b200msm_set_max_chunk(0)
    L.b200msm_set_heavy_factor(0)


def test_divsteps_inverse_matches_bigint(eng):
    """op 6 of the field hook = csrc/modinv.cuh on the device, Fp and Fp2, incl. 1, p−1 and small values"""
    u = eng._lib.u64p
    rng = random.Random(61)
    vals = [1, 2, o.P - 1, o.P - 2, (o.P - 1) // 2, 3, 1 << 30, (1 << 30) - 1, 1 << 64, o.MONT_RINV] + [rng.randrange(1, o.P) for _ in range(1000)]
    a = np.array([o.FpOps.to_limbs(v) for v in vals], dtype=np.uint64)
    out = np.zeros_like(a)
    assert eng._lib.lib.b200msm_dbg_field_op(0, 6, a.ctypes.data_as(u), a.ctypes.data_as(u), out.ctypes.data_as(u), a.shape[0]) == 0
    assert np.array_equal(out, np.array([o.FpOps.to_limbs(o.FpOps.inv(v)) for v in vals], dtype=np.uint64))
    av, a2 = rand_fp2_limbs(rng, 300, edge=False)
    av[:3] = [(1, 0), (0, 1), (o.P - 1, o.P - 1)]
    a2 = np.array([o.Fp2Ops.to_limbs(v) for v in av], dtype=np.uint64)
    out2 = np.zeros_like(a2)
    assert eng._lib.lib.b200msm_dbg_field_op(1, 6, a2.ctypes.data_as(u), a2.ctypes.data_as(u), out2.ctypes.data_as(u), a2.shape[0]) == 0
    assert np.array_equal(out2, np.array([o.Fp2Ops.to_limbs(o.Fp2Ops.inv(v)) for v in av], dtype=np.uint64))


@pytest.mark.parametrize("g2", [0, 1])
@pytest.mark.parametrize("rounds", [1, 2, 3])
def test_exceptional_cases_inside_a_batch(eng, knobs, g2, rounds):
    """few distinct points and tiny scalars: every bucket is full of repeated points, opposite points and
    identities, so the rounds hit doubling, cancellation, copies and empty slots at every level"""
    C = curve(g2)
    rng = random.Random(500 + 10 * g2 + rounds)
    base = [C.mul(C.gen, rng.randrange(1, o.R_ORDER)) for _ in range(3)]
    pool = base + [C.neg(p) for p in base] + [None]
    assert knobs.b200msm_set_batch_affine(rounds) == 0
    for n, smax in ((1, 3), (2, 3), (5, 2), (64, 4), (257, 3), (700, 6)):
        pts = [pool[rng.randrange(len(pool))] for _ in range(n)]
        sc = [rng.randrange(smax) for _ in range(n)]
        exp = C.msm_naive(pts, sc)
        for glv in (0, 1):
            assert knobs.b200msm_set_glv(glv) == 0
            got = _grp(eng, g2).msm(points_to_limbs(C, pts), scalars_to_limbs(sc, True))
            assert C.eq(jac_to_point(C, got), exp), (n, smax, glv)
    # the same point n times with the same scalar: log2(n) levels of pure doublings
    P = base[0]
    for n in (2, 4, 7, 8, 33):
        exp = C.mul(P, 5 * n % o.R_ORDER)
        got = _grp(eng, g2).msm(points_to_limbs(C, [P] * n), scalars_to_limbs([5] * n, True))
        assert C.eq(jac_to_point(C, got), exp), n


@pytest.mark.parametrize("g2,n", [(0, 3000), (0, (1 << 15) + 11), (1, 2500)])
@pytest.mark.parametrize("rounds", [0, 1, 2, 3])
def test_rounds_vs_c_oracle(eng, cref, knobs, g2, n, rounds):
    bases = cref.synth_bases(g2, 3100 + n, n)
    sm = cref.synth_scalars(3200 + n, n, True)
    sc = cref.synth_scalars(3200 + n, n, False)
    exp = cref.msm(g2, bases, sc, 0)
    assert knobs.b200msm_set_batch_affine(rounds) == 0
    for c, glv in ((0, -1), (8, 0), (11, 1), (5, 0)):   # small widths make full buckets, where the rounds do real additions
        assert knobs.b200msm_set_window_bits(c) == 0 and knobs.b200msm_set_glv(glv) == 0
        assert cref.affine_equal(g2, _grp(eng, g2).msm(bases, sm), exp), (c, glv)
        assert cref.affine_equal(g2, _grp(eng, g2).msm_bigint(bases, sc), exp), (c, glv)


@pytest.mark.parametrize("rounds", [1, 2])
def test_rounds_with_every_host_path(eng, cref, knobs, rounds):
    """streamed slices (accumulate INTO shared buckets), chunked passes, resident bases, fixed-base table,
    witness-like scalars (one huge bucket: the heavy path reads what the rounds left)"""
    g2, n = 0, 40000
    bases = cref.synth_bases(g2, 4100, n)
    sc = cref.synth_scalars(4200, n, False)
    sc[: n // 2] = 0
    sc[n // 2: n // 2 + n // 4, 0] = 1
    sc[n // 2: n // 2 + n // 4, 1:] = 0
    sm = np.array([o.scalar_to_limbs(o.limbs_to_int(r), True) for r in sc.tolist()], dtype=np.uint64)
    exp = cref.msm(g2, bases, sc, 0)
    assert knobs.b200msm_set_batch_affine(rounds) == 0
    assert knobs.b200msm_set_window_bits(9) == 0
    G = eng.G1Projective
    assert cref.affine_equal(g2, G.msm(bases, sm), exp)
    assert knobs.b200msm_set_stream_slices(8, 1024) == 0          # streamed
    assert cref.affine_equal(g2, G.msm(bases, sm), exp)
    assert knobs.b200msm_set_stream_slices(1, 0) == 0
    assert knobs.b200msm_set_max_chunk(7000) == 0                 # chunked
    assert cref.affine_equal(g2, G.msm(bases, sm), exp)
    assert knobs.b200msm_set_max_chunk(0) == 0
    assert knobs.b200msm_set_heavy_factor(1) == 0                 # many heavy buckets
    assert cref.affine_equal(g2, G.msm(bases, sm), exp)
    assert knobs.b200msm_set_heavy_factor(0) == 0
    assert knobs.b200msm_set_window_bits(0) == 0
    rb = eng.ResidentBases(G, bases)
    try:
        assert cref.affine_equal(g2, rb.msm(sm), exp)
        rb.precompute(10)
        assert cref.affine_equal(g2, rb.msm(sm), exp)
        assert cref.affine_equal(g2, rb.msm(sm[:12345]), cref.msm(g2, bases[:12345], sc[:12345], 0))
    finally:
        rb.close()


@pytest.mark.parametrize("g2", [0, 1])
def test_graph_replay_is_bit_exact(eng, cref, knobs, g2):
    """the same call four times: issued directly, recorded as CUDA graphs on the second sight, replayed after;
    then with graphs off; then new scalars in the same buffers (the graphs hold pointers, not data)"""
    import torch

    n = 20000
    aw, jw = (24, 36) if g2 else (12, 18)
    G = eng.G2 if g2 else eng.G1
    bases = torch.empty((n, aw), dtype=torch.int64, device="cuda")
    scalars = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    out = torch.zeros(jw, dtype=torch.int64, device="cuda")
    eng.synth_bases_device(G, 5100, n, bases.data_ptr())
    st = torch.cuda.Stream()
    for seed in (5200, 5201):
        eng.synth_scalars_device(seed, n, True, scalars.data_ptr())
        torch.cuda.synchronize()
        exp = cref.msm_by_dlog(g2, 5100, cref.synth_scalars(seed, n, False))
        for rep in range(4):
            out.zero_()
            eng.run_device(G, bases.data_ptr(), scalars.data_ptr(), n, True, out.data_ptr(), st.cuda_stream)
            st.synchronize()
            assert cref.affine_equal(g2, out.cpu().numpy().view(np.uint64), exp), (seed, rep)
    launches0 = knobs.b200msm_launch_count()
    eng.run_device(G, bases.data_ptr(), scalars.data_ptr(), n, True, out.data_ptr(), st.cuda_stream)
    st.synchronize()
    replay = knobs.b200msm_launch_count() - launches0
    assert knobs.b200msm_set_graphs(0) == 0
    launches0 = knobs.b200msm_launch_count()
    eng.run_device(G, bases.data_ptr(), scalars.data_ptr(), n, True, out.data_ptr(), st.cuda_stream)
    st.synchronize()
    assert knobs.b200msm_launch_count() - launches0 == replay > 20      # the graph runs the same kernels, and says so
    assert cref.affine_equal(g2, out.cpu().numpy().view(np.uint64), exp)
    # the legacy default stream cannot be captured: the library records on a stream of its own
    assert knobs.b200msm_set_graphs(1) == 0
    for rep in range(3):
        eng.run_device(G, bases.data_ptr(), scalars.data_ptr(), n, True, out.data_ptr(), 0)
        torch.cuda.synchronize()
        assert cref.affine_equal(g2, out.cpu().numpy().view(np.uint64), exp)


def test_two_pipeline_rounds_repeated(eng, cref):
    """Large rounds run as two half-range pipelines on two streams, which may drift a round apart: every array a
    pipeline owns must keep its place across rounds (a first version moved part 1's scratch with the round's plan and
    lost ≈1 result in 25).  Back-to-back table + plain MSMs at 2^20, repeated, results checked every time."""
    import torch

    g2, n = 0, 1 << 20
    st = torch.cuda.current_stream().cuda_stream
    c, W = eng.table_plan(g2, n)
    table = torch.empty((W, n, 12), dtype=torch.int64, device="cuda")
    scalars = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    eng.synth_bases_device(g2, 6100, n, table.data_ptr())
    eng.table_build_device(g2, table.data_ptr(), n, c, table.data_ptr(), st)
    out = torch.zeros((2, 18), dtype=torch.int64, device="cuda")
    for it in range(10):
        eng.synth_scalars_device(6200 + it, n, True, scalars.data_ptr())
        eng.run_table_device(g2, table.data_ptr(), n, c, scalars.data_ptr(), n, True, out[0].data_ptr(), st)
        eng.run_device(g2, table.data_ptr(), scalars.data_ptr(), n, True, out[1].data_ptr(), st)
        torch.cuda.synchronize()
        got = out.cpu().numpy().view(np.uint64)
        exp = cref.msm_by_dlog(g2, 6100, cref.synth_scalars(6200 + it, n, False))
        assert cref.affine_equal(g2, got[0], exp), ("table", it)
        assert cref.affine_equal(g2, got[1], exp), ("plain", it)


@pytest.mark.parametrize("n", [1, 5, 700, 20011, (1 << 16) + 3])
def test_g2_four_part_decomposition(eng, cref, knobs, n):
    """b200msm_set_glv(2): G2 scalars as four base-|z| digits over Q, −ψ(Q), ψ²(Q), −ψ³(Q) (csrc/gls4.cuh) — against
    the two-part split, no split, and the C oracle; with and without batched-affine rounds; identity bases included"""
    g2 = 1
    bases = cref.synth_bases(g2, 7100 + n, n)
    if n > 4:
        bases[3] = 0                                            # an identity base: all its images are the identity too
    sm = cref.synth_scalars(7200 + n, n, True)
    sc = cref.synth_scalars(7200 + n, n, False)
    if n > 4:
        sc[0] = 0
        sc[1] = np.array(o.int_to_limbs(o.R_ORDER - 1, 4), dtype=np.uint64)
        sc[2] = np.array(o.int_to_limbs((-o.BLS_X) ** 3, 4), dtype=np.uint64)     # digits (0, 0, 0, 1)
        sm = np.array([o.scalar_to_limbs(o.limbs_to_int(r), True) for r in sc.tolist()], dtype=np.uint64)
    exp = cref.msm(g2, bases, sc, 0)
    for mode in (2, 1, 0):
        assert knobs.b200msm_set_glv(mode) == 0
        for rounds, c in ((-1, 0), (0, 0), (2, 8), (3, 13)):
            assert knobs.b200msm_set_batch_affine(rounds) == 0 and knobs.b200msm_set_window_bits(c) == 0
            assert cref.affine_equal(g2, eng.G2Projective.msm(bases, sm), exp), (mode, rounds, c)
            assert cref.affine_equal(g2, eng.G2Projective.msm_bigint(bases, sc), exp), (mode, rounds, c)
    assert knobs.b200msm_set_glv(2) == 0 and knobs.b200msm_set_window_bits(0) == 0 and knobs.b200msm_set_batch_affine(-1) == 0
    eng.G2Projective.msm(bases, sm)
    assert eng.last_plan()["glv"] in (3, 4)
    # G1 has no ψ: mode 2 means the two-part split there
    b1 = cref.synth_bases(0, 7300 + n, n)
    assert cref.affine_equal(0, eng.G1Projective.msm(b1, sm), cref.msm(0, b1, sc, 0))
    assert eng.last_plan()["glv"] in (1, 2)


@pytest.mark.parametrize("g2,n", [(0, 1 << 19), (0, 3 << 18), (0, (1 << 19) + 12345), (1, 1 << 19), (1, 3 << 17)])
def test_two_pipeline_rounds_sizes_between_the_clamps(eng, cref, g2, n):
    """sizes whose per-pipeline plans are NOT at the K = 32 clamp: there the thread count of a later round can exceed
    the first round's (K halves with the slot count), so a pipeline's scratch must be sized by the maximum over the
    rounds — the 8-GPU Groth16 shape (2^19 points per GPU) found a layout that used the first round's sizes"""
    import torch

    aw, jw = (24, 36) if g2 else (12, 18)
    bases = torch.empty((n, aw), dtype=torch.int64, device="cuda")
    scalars = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    out = torch.zeros(jw, dtype=torch.int64, device="cuda")
    eng.synth_bases_device(g2, 8100 + g2, n, bases.data_ptr())
    st = torch.cuda.current_stream().cuda_stream
    for it in range(4):
        eng.synth_scalars_device(8200 + it, n, True, scalars.data_ptr())
        eng.run_device(g2, bases.data_ptr(), scalars.data_ptr(), n, True, out.data_ptr(), st)
        torch.cuda.synchronize()
        exp = cref.msm_by_dlog(g2, 8100 + g2, cref.synth_scalars(8200 + it, n, False))
        assert cref.affine_equal(g2, out.cpu().numpy().view(np.uint64), exp), it
