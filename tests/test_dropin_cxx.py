"""The C++ drop-in (include/itkMultigridAnisotropicDiffusionImageFilter.h + include/mad/itkMultigridSmootherTags.h over
include/madgpu.h) compiled against the stand-in ITK of oracle/shim and driven by tests/cxx/dropin_test.cxx the way
the reference's test programs drive the original filter.  CPU: it compiles, links libmadgpu.so, and -- there being no
CPU fallback -- fails loudly without a GPU.  GPU: its output equals the golden vectors recorded from the reference."""
import os
import subprocess
import sys

import numpy as np
import pytest

from util import GOLDEN, ROOT, load_lena, random_image, random_spd_tensor, rel_l2

PKG = os.path.join(ROOT, "multigridanisotropicdiffusion_b200")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("dropin") / "dropin_test")
    cmd = ["g++", "-O1", "-std=c++14", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle", "shim"),
           "-o", out, os.path.join(ROOT, "tests", "cxx", "dropin_test.cxx"), "-L" + PKG, "-lmadgpu", "-Wl,-rpath," + PKG]
    subprocess.check_call(cmd)
    return out


def _run(exe, tmp, dim, smoother, cycle, nu, dt, tol, steps, pixel, img, tensor, spacing):
    dt_map = {"f32": np.float32, "f64": np.float64, "i16": np.int16, "u8": np.uint8}
    a, b, c = str(tmp / "in.raw"), str(tmp / "tensor.raw"), str(tmp / "out.raw")
    img.astype(dt_map[pixel]).tofile(a)
    tensor.astype(dt_map[pixel] if pixel in ("f32", "f64") else np.float64).tofile(b)
    n = [str(s) for s in img.shape[::-1]]
    cmd = [exe, str(dim), smoother, cycle, str(nu), repr(dt), repr(tol), str(steps), pixel, a, b, c] + n + [repr(float(s)) for s in spacing]
    r = subprocess.run(cmd, capture_output=True, text=True)
    out = np.fromfile(c, dtype=dt_map[pixel]).reshape(img.shape) if r.returncode == 0 else None
    return r, out


def test_dropin_compiles_and_refuses_to_run_without_a_gpu(exe, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    img = random_image((16, 16), seed=1)
    T = random_spd_tensor((16, 16), seed=1)
    r, out = _run(exe, tmp_path, 2, "wj", "v", 2, 0.1, 1e-6, 1, "f32", img, T, (1.0, 1.0))
    assert r.returncode == 2 and out is None
    assert "madgpu_create" in r.stderr and "CUDA" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("smoother,cycle", [("wj", "v"), ("gs", "fmg")])
def test_dropin_reference_2d_test(exe, tmp_path, smoother, cycle):
    """test/itk2DDiffusionTest_{WJ,GS}.cxx through the drop-in header: float image, float tensor."""
    g = np.load(os.path.join(GOLDEN, f"ref_lena_{smoother}_{cycle}.npz"))
    img = load_lena().astype(np.float32)
    T = np.zeros(img.shape + (3,), dtype=np.float32)
    T[..., 0] = 50.0
    T[..., 2] = 30.0
    r, out = _run(exe, tmp_path, 2, smoother, cycle, 2, 0.1, 1e-10, 1, "f32", img, T, (1.0, 1.0))
    assert r.returncode == 0, r.stderr
    sub = int(g["sub"])
    assert rel_l2(out[::sub, ::sub], g["sample"]) < 5e-7
    cycles = int(r.stdout.split("cycles")[1].split()[0])
    assert abs(cycles - int(g["cycles"][0])) <= (1 if smoother == "wj" else 3)


@pytest.mark.gpu
@pytest.mark.parametrize("pixel", ["f64", "i16"])
def test_dropin_3d_gauss_seidel(exe, tmp_path, pixel):
    g = np.load(os.path.join(GOLDEN, "ref_small3d_gs_v.npz"))
    shape, sp = (23, 25, 27), (0.3125, 0.3125, 0.5)
    T = random_spd_tensor(shape, seed=2).astype(np.float64)
    img = random_image(shape, seed=5).astype(np.float64)
    if pixel == "f64":
        r, out = _run(exe, tmp_path, 3, "gs", "v", 2, 0.1, 1e-10, 2, "f64", img, T, sp)
        assert r.returncode == 0, r.stderr
        assert rel_l2(out, g["sample"]) < 1e-6
    else:
        # short pixels: the tensor image has the pixel type too, use integer-valued entries
        Ti = np.round(8.0 * T)
        Ti[..., [0, 3, 5]] += 4.0
        imgi = np.round(img)
        r, out = _run(exe, tmp_path, 3, "gs", "v", 2, 0.1, 1e-10, 1, "i16", imgi, Ti, sp)
        assert r.returncode == 0, r.stderr
        from oracle import oracle as O
        o = O.Oracle(shape, sp, Ti, 0.1, smoother=0, nu=2)
        ref, _, _ = o.solve(imgi, tolerance=1e-10)
        d = np.abs(out.astype(np.int32) - np.trunc(ref).astype(np.int32))
        assert d.max() <= 1 and (d != 0).mean() < 1e-3
