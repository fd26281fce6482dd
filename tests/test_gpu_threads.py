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
