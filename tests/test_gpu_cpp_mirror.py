"""Builds and runs tests/cpp/test_msm_mirror.cpp — the C++ twin of the reference's own MSM tests
(src/tests.rs:50-67, src/g1.rs:695-709) — against libb200msm.so on the GPU."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_mirror_on_gpu(tmp_path, cref):
    exe = str(tmp_path / "test_msm_mirror")
    libdir = os.path.join(ROOT, "ark_blst_b200")
    odir = os.path.join(ROOT, "oracle")
    subprocess.run(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_msm_mirror.cpp"),
                    "-L", libdir, "-lb200msm", "-L", odir, "-lmsm_ref", f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{odir}"], check=True)
    g1 = "".join("%016x" % int(v) for v in cref.generator_limbs(0))
    g2 = "".join("%016x" % int(v) for v in cref.generator_limbs(1))
    r = subprocess.run([exe, g1, g2], capture_output=True, text=True)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr
