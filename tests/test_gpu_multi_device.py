"""Single-process multi-device path of the C-ABI (b200msm_init(first, N) — the shape a Rust caller
uses): points sharded by index range over N GPUs, partials gathered on device 0 by peer copies,
final addition on device. Skipped when fewer than 2 GPUs are visible."""
import ctypes
import subprocess
import sys
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys, ctypes
import numpy as np
sys.path.insert(0, %r)
import ark_blst_b200 as eng
from oracle import cref
L = eng._lib.lib
ndev = int(sys.argv[1])
assert L.b200msm_init(0, ndev) == 0, L.b200msm_last_error()
assert L.b200msm_device_count() == ndev
for g2, n in ((0, 7), (0, 20011), (1, 3001)):
    bases = cref.synth_bases(g2, 5 + n, n)
    sm = cref.synth_scalars(6 + n, n, True)
    exp = cref.msm(g2, bases, sm, 1)
    grp = eng.G2Projective if g2 else eng.G1Projective
    assert cref.affine_equal(g2, grp.msm(bases, sm), exp), ("one-shot", g2, n)
    rb = eng.ResidentBases(grp, bases)
    assert cref.affine_equal(g2, rb.msm(sm), exp), ("resident", g2, n)
    half = n // 2
    assert cref.affine_equal(g2, rb.msm(sm[:half]), cref.msm(g2, bases[:half], sm[:half], 1)), ("prefix", g2, n)
    rb.close()
L.b200msm_shutdown()
print("ok", ndev)
"""


def _ngpu():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.parametrize("ndev", [2, 4, 8])
def test_single_process_multi_device(ndev):
    if _ngpu() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    r = subprocess.run([sys.executable, "-c", CHILD % ROOT, str(ndev)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and f"ok {ndev}" in r.stdout, r.stdout + r.stderr
