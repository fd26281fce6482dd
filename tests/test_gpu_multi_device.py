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
assert L.b200msm_init(0, ndev) == 0                      # same binding again: fine
assert L.b200msm_init(0, 1) != 0 and b"already bound" in L.b200msm_last_error()   # another range: an error, not a silent no-op

def check(g2, n, seed, tag):
    bases = cref.synth_bases(g2, seed, n)
    sm = cref.synth_scalars(seed + 1, n, True)
    sc = cref.synth_scalars(seed + 1, n, False)
    exp = cref.msm_by_dlog(g2, seed, sc) if n > 30000 else cref.msm(g2, bases, sm, 1)
    grp = eng.G2Projective if g2 else eng.G1Projective
    assert cref.affine_equal(g2, grp.msm(bases, sm), exp), ("one-shot", tag, g2, n)
    assert cref.affine_equal(g2, grp.msm_bigint(bases, sc), exp), ("one-shot bigint", tag, g2, n)
    rb = eng.ResidentBases(grp, bases)
    assert cref.affine_equal(g2, rb.msm(sm), exp), ("resident", tag, g2, n)
    half = n // 2
    exp_half = cref.msm_by_dlog(g2, seed, sc[:half]) if n > 30000 else cref.msm(g2, bases[:half], sm[:half], 1)
    assert cref.affine_equal(g2, rb.msm(sm[:half]), exp_half), ("prefix", tag, g2, n)
    rb.precompute()                                          # the same handle as per-device fixed-base tables
    assert cref.affine_equal(g2, rb.msm(sm), exp), ("table", tag, g2, n)
    assert cref.affine_equal(g2, rb.msm(sm[:half]), exp_half), ("table prefix", tag, g2, n)
    rb.close()

for g2, n in ((0, 7), (0, 20011), (1, 3001)):
    check(g2, n, 5 + n, "plain")
# every device's share goes down the streamed branch (uploads in slices accumulated into shared buckets)
assert L.b200msm_set_stream_slices(8, 1024) == 0
for g2, n in ((0, 20011 * ndev // 2), (1, 4099 * ndev)):
    check(g2, n, 9 + n, "streamed-forced")
assert L.b200msm_set_stream_slices(8, 0) == 0
# real size: 2^18 points per device is the default streaming threshold (G1), G2 at 2^16 per device
check(0, (1 << 18) * ndev, 77, "streamed")
check(1, (1 << 16) * ndev, 78, "g2")
L.b200msm_shutdown()
assert L.b200msm_device_count() == 0
assert L.b200msm_init(0, 1) == 0                         # re-binding after shutdown works
bases = cref.synth_bases(0, 3, 100); sm = cref.synth_scalars(4, 100, True)
assert cref.affine_equal(0, eng.G1Projective.msm(bases, sm), cref.msm(0, bases, sm, 1))
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
    r = subprocess.run([sys.executable, "-c", CHILD % ROOT, str(ndev)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and f"ok {ndev}" in r.stdout, r.stdout + r.stderr
