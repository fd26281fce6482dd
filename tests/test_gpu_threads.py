"""Concurrent callers: arkworks provers call msm from rayon worker threads (SURVEY §8b
"Threading"), so the C-ABI must be safe under concurrent calls from several host threads.
ctypes releases the GIL during the foreign call, so these threads really overlap inside the
library (per-device mutex + GPU-side serialisation of the scratch arena)."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_concurrent_host_threads(eng, cref):
    jobs = []
    for t in range(6):
        g2 = t % 3 == 2
        n = 3000 + 517 * t
        bases = cref.synth_bases(int(g2), 300 + t, n)
        sc = cref.synth_scalars(400 + t, n, True)
        jobs.append((g2, bases, sc, cref.msm(int(g2), bases, sc, 1)))
    errors = []

    def work(j, reps):
        g2, bases, sc, exp = jobs[j]
        grp = eng.G2Projective if g2 else eng.G1Projective
        try:
            for _ in range(reps):
                got = grp.msm(bases, sc)
                if not cref.affine_equal(int(g2), got, exp):
                    errors.append(("mismatch", j))
        except Exception as e:  # noqa: BLE001
            errors.append((repr(e), j))

    threads = [threading.Thread(target=work, args=(j, 5)) for j in range(len(jobs))]
    for th in threads:
        th.start()
    for th in threads:
        th.join(180)
    assert not errors, errors


def test_concurrent_resident_and_oneshot(eng, cref):
    n = 4000
    bases = cref.synth_bases(0, 777, n)
    rb = eng.ResidentBases(eng.G1Projective, bases)
    scs = [cref.synth_scalars(800 + k, n, True) for k in range(4)]
    exps = [cref.msm(0, bases, s, 1) for s in scs]
    errors = []

    def work(k):
        for _ in range(4):
            a = rb.msm(scs[k])
            b = eng.G1Projective.msm(bases, scs[k])
            if not (cref.affine_equal(0, a, exps[k]) and cref.affine_equal(0, b, exps[k])):
                errors.append(k)

    ths = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    [t.start() for t in ths]
    [t.join(180) for t in ths]
    rb.close()
    assert not errors, errors


def test_setters_flipped_under_load(eng, cref):
    """A thread flips the plan knobs (GLV mode, window width, slicing) while others run MSMs: every call
    plans from ONE snapshot of the tunables, so results stay bit-exact whatever it catches."""
    L = eng._lib.lib
    n = 6000
    bases = cref.synth_bases(0, 911, n)
    sc = cref.synth_scalars(912, n, True)
    exp = cref.msm(0, bases, sc, 1)
    rb = eng.ResidentBases(eng.G1Projective, bases)
    stop = threading.Event()
    errors = []

    def flip():
        k = 0
        while not stop.is_set():
            L.b200msm_set_glv((-1, 0, 1)[k % 3])
            L.b200msm_set_window_bits((0, 9, 12, 0, 14)[k % 5])
            L.b200msm_set_stream_slices(1 + k % 8, 1024)
            L.b200msm_set_heavy_factor((0, 2, 5)[k % 3])
            k += 1

    def work(resident):
        try:
            for _ in range(12):
                got = rb.msm(sc) if resident else eng.G1Projective.msm(bases, sc)
                if not cref.affine_equal(0, got, exp):
                    errors.append("mismatch")
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    fl = threading.Thread(target=flip)
    ws = [threading.Thread(target=work, args=(k % 2 == 0,)) for k in range(4)]
    fl.start()
    [t.start() for t in ws]
    [t.join(300) for t in ws]
    stop.set()
    fl.join(10)
    L.b200msm_set_glv(-1); L.b200msm_set_window_bits(0); L.b200msm_set_stream_slices(8, 0); L.b200msm_set_heavy_factor(0)
    rb.close()
    assert not errors, errors
