"""The C++ drop-in of the VED filter (include/itkVEDMultigridImageFilter.h over include/madved.h and include/madgpu.h) compiled
against the stand-in ITK of oracle/shim and driven by tests/cxx/ved_dropin_test.cxx the way the reference's test/itkVEDTest_GS.cxx
drives the original.  CPU: it compiles, links libmadgpu.so and fails loudly without a GPU.  GPU: its output equals the vectors
recorded from the reference's own VED code (tests/golden/make_golden_ved.py)."""
import os
import subprocess

import numpy as np
import pytest

from util import GOLDEN, ROOT, random_image, rel_l2

PKG = os.path.join(ROOT, "multigridanisotropicdiffusion_b200")


@pytest.fixture(scope="module")
def ved_exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("dropin_ved") / "ved_dropin_test")
    cmd = ["g++", "-O1", "-std=c++14", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle", "shim"),
           "-o", out, os.path.join(ROOT, "tests", "cxx", "ved_dropin_test.cxx"), "-L" + PKG, "-lmadgpu", "-Wl,-rpath," + PKG]
    subprocess.check_call(cmd)
    return out


def _run_ved(exe, tmp, cycle, pixel, img, spacing):
    dt = {"i16": np.int16, "f64": np.float64}[pixel]
    a, b = str(tmp / "in.raw"), str(tmp / "out.raw")
    img.astype(dt).tofile(a)
    cmd = [exe, cycle, pixel, a, b] + [str(s) for s in img.shape[::-1]] + [repr(float(s)) for s in spacing]
    r = subprocess.run(cmd, capture_output=True, text=True)
    out = np.fromfile(b, dtype=dt).reshape(img.shape) if r.returncode == 0 else None
    return r, out


def test_ved_dropin_compiles_and_refuses_to_run_without_a_gpu(ved_exe, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r, out = _run_ved(ved_exe, tmp_path, "v", "i16", random_image((16, 16, 16), seed=1), (0.3125, 0.3125, 0.5))
    assert r.returncode == 2 and out is None
    assert "madgpu_create" in r.stderr and "CUDA" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("pixel", ["i16", "f64"])
def test_ved_dropin_reference_test(ved_exe, tmp_path, pixel):
    """test/itkVEDTest_GS.cxx ("v") through the drop-in header, against the vectors of the reference's own VED code."""
    from util import load_ved_test
    g = np.load(os.path.join(GOLDEN, "ref_vedfilter_gs_v.npz"))
    img, sp = load_ved_test()
    r, out = _run_ved(ved_exe, tmp_path, "v", pixel, img, sp)
    assert r.returncode == 0, r.stderr
    assert "steps 4" in r.stdout and "scales 5" in r.stdout
    if pixel == "f64":
        sub = int(g["sub"])
        assert rel_l2(out[::sub, ::sub, ::sub], g["sample"]) < 1e-4
    else:
        d = np.abs(out.astype(int) - g["out_short"].astype(int))
        assert d.max() <= 1 and (d != 0).mean() < 1e-3
