"""The reference's OWN test programs -- /root/reference/test/itk2DDiffusionTest_GS.cxx, itk2DDiffusionTest_WJ.cxx and
itkVEDTest_GS.cxx, unmodified and compiled from where they lie -- built against this repo's drop-in headers (include/itk*.h,
include/mad/*.h) and the stand-in ITK (oracle/shim), linked with libmadgpu.so, and run as the reference's test/CMakeLists.txt
registers them: 3 programs x {v, fmg, s} = its nine CTest entries.  Nothing of the reference is copied; the binary
(tests/_build/ref_tests_dropin, tests/cxx/build_ref_tests.sh) is built where /root/reference exists and travels as a built artefact.

The programs read test_data/lena.jpg and test_data/ved_test.mhd relative to the working directory and assert nothing themselves
(SURVEY section 4); here their outputs are compared with the vectors recorded from the reference's own code.  The stand-in reader
serves lena.jpg through a MetaImage side-car (no JPEG library in this image).
CPU: the programs compile, link, read their inputs and -- no CPU fallback -- stop at madgpu_create; and the same binary is run
against the EMULATED device (LD_LIBRARY_PATH pointing at the host build of the CUDA source, tests/mad_host/), which checks the whole
chain -- reader stand-in, drop-in header, C-ABI, kernels, output cast, writer -- on the CPU.  GPU: all nine entries."""
import os
import subprocess

import numpy as np
import pytest

from multigridanisotropicdiffusion_b200 import metaimage
from util import GOLDEN, ROOT, load_lena, load_ved_test

EXE = os.path.join(ROOT, "tests", "_build", "ref_tests_dropin")


@pytest.fixture(scope="module")
def exe():
    if os.path.isdir("/root/reference/test"):
        subprocess.check_call([os.path.join(ROOT, "tests", "cxx", "build_ref_tests.sh")], stdout=subprocess.DEVNULL)
    if not os.path.exists(EXE):
        pytest.skip("tests/_build/ref_tests_dropin not built (needs /root/reference)")
    return EXE


@pytest.fixture
def workdir(tmp_path):
    """cwd of the reference's tests: test_data/ with their two inputs."""
    d = tmp_path / "test_data"
    d.mkdir()
    vol, sp = load_ved_test()
    metaimage.write(str(d / "ved_test.mhd"), vol, {"spacing": sp, "TransformMatrix": "-1 0 0 0 -1 0 0 0 1", "Offset": "-13.9881 -27.1641 -52.1181"})
    metaimage.write(str(d / "lena.jpg.mhd"), load_lena(), {"spacing": (1.0, 1.0)}, compressed=False)
    return tmp_path


def _run(exe, cwd, test, mode, env=None):
    return subprocess.run([exe, test, mode], cwd=str(cwd), capture_output=True, text=True, timeout=2400, env=env)


@pytest.fixture(scope="module")
def emulated_env(tmp_path_factory):
    """Environment in which `libmadgpu.so` resolves to the host build of the CUDA source (the binary's RUNPATH comes after
    LD_LIBRARY_PATH)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "mad_host"))
    import hostlib
    d = tmp_path_factory.mktemp("emulated_device")
    os.symlink(hostlib.build(), str(d / "libmadgpu.so"))
    return dict(os.environ, LD_LIBRARY_PATH=str(d) + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))


def test_reference_test_programs_build_against_the_dropin_and_need_a_gpu(exe, workdir):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    for test in ("itkVEDTest_GS", "itk2DDiffusionTest_WJ", "itk2DDiffusionTest_GS"):
        r = _run(exe, workdir, test, "v")
        assert r.returncode == 2, (r.stdout, r.stderr)
        assert "size [69, 77, 69]" in r.stdout or "size [512, 512]" in r.stdout  # the input was read
        assert "madgpu_create" in r.stderr and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["v", "fmg", "s"])
@pytest.mark.parametrize("smoother", ["GS", "WJ"])
def test_itk2DDiffusionTest(exe, workdir, smoother, mode):
    r = _run(exe, workdir, f"itk2DDiffusionTest_{smoother}", mode)
    assert r.returncode == 0, r.stderr[-2000:]
    out, _ = metaimage.read(str(workdir / "test_data" / "lena_out.jpg.mhd"))  # the float result cast back to unsigned char (:128-138)
    diff, _ = metaimage.read(str(workdir / "test_data" / "lena_diff.jpg.mhd"))
    assert out.dtype == np.uint8 and out.shape == (512, 512) and diff.shape == (512, 512)
    g = np.load(os.path.join(GOLDEN, f"ref_lena_{smoother.lower()}_{mode}.npz"))
    sub = int(g["sub"])
    want = g["sample"].astype(np.float32).astype(np.uint8)  # static_cast< unsigned char >( float )
    d = np.abs(out[::sub, ::sub].astype(int) - want.astype(int))
    assert d.max() <= 1 and (d != 0).mean() < (0.05 if (smoother, mode) == ("GS", "s") else 0.005)
    assert np.abs(out.astype(int) - load_lena().astype(int)).max() > 5  # it did diffuse


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["v", "fmg", "s"])
def test_itkVEDTest_GS(exe, workdir, mode):
    r = _run(exe, workdir, "itkVEDTest_GS", mode)
    assert r.returncode == 0, r.stderr[-2000:]
    out, meta = metaimage.read(str(workdir / "test_data" / "ved_test_out.mhd"))
    vol, sp = load_ved_test()
    assert out.dtype == np.int16 and out.shape == vol.shape and meta["spacing"] == sp
    assert [float(x) for x in meta["TransformMatrix"].split()] == [-1, 0, 0, 0, -1, 0, 0, 0, 1]  # direction restored by the test (:108-117)
    assert np.abs(out.astype(int) - vol.astype(int)).max() > 5
    if mode != "s":  # V-cycles and FMG converge to the same image (tolerance 1e-10); 100 smoother sweeps do not converge
        g = np.load(os.path.join(GOLDEN, "ref_vedfilter_gs_v.npz"))
        d = np.abs(out.astype(int) - g["out_short"].astype(int))
        assert d.max() <= 1 and (d != 0).mean() < 1e-3


# ---- the same binary on the emulated device (CPU) ------------------------------------------------------------------------------
@pytest.mark.parametrize("smoother,mode", [("WJ", "v"), ("GS", "fmg")])
def test_itk2DDiffusionTest_on_the_emulated_device(exe, workdir, emulated_env, smoother, mode):
    r = _run(exe, workdir, f"itk2DDiffusionTest_{smoother}", mode, env=emulated_env)
    assert r.returncode == 0, r.stderr[-2000:]
    out, _ = metaimage.read(str(workdir / "test_data" / "lena_out.jpg.mhd"))
    g = np.load(os.path.join(GOLDEN, f"ref_lena_{smoother.lower()}_{mode}.npz"))
    sub = int(g["sub"])
    d = np.abs(out[::sub, ::sub].astype(int) - g["sample"].astype(np.float32).astype(np.uint8).astype(int))
    assert d.max() <= 1 and (d != 0).mean() < 0.005
    cycles = r.stdout.count("|--- VCycle n.")
    assert abs(cycles - int(g["cycles"][0])) <= (1 if smoother == "WJ" else 3)


def test_itkVEDTest_GS_on_the_emulated_device(exe, tmp_path, emulated_env):
    """itkVEDTest_GS v.  By default on a 32x40x64 crop of ved_test.mhd against the oracle's GenerateData (about 20 s); with
    MADGPU_FULL_EMULATION=1 on the whole volume against the vector recorded from the reference's own code (about 90 s; the result
    of that run is identical to the golden vector in every voxel)."""
    from oracle import ved as V
    full = os.environ.get("MADGPU_FULL_EMULATION") == "1"
    vol, sp = load_ved_test()
    if not full:
        vol = np.ascontiguousarray(vol[18:50, 20:60, 2:66])
    d = tmp_path / "test_data"
    d.mkdir()
    metaimage.write(str(d / "ved_test.mhd"), vol, {"spacing": sp, "TransformMatrix": "-1 0 0 0 -1 0 0 0 1"})
    r = _run(exe, tmp_path, "itkVEDTest_GS", "v", env=emulated_env)
    assert r.returncode == 0, r.stderr[-2000:]
    out, meta = metaimage.read(str(d / "ved_test_out.mhd"))
    assert out.dtype == np.int16 and out.shape == vol.shape and meta["spacing"] == sp
    if full:
        want = np.load(os.path.join(GOLDEN, "ref_vedfilter_gs_v.npz"))["out_short"]
    else:
        want, _ = V.ved_filter(vol, sp, V.DEFAULT_SCALES, alpha=0.5, beta=0.5, gamma=5.0, epsilon=0.01, omega=1.5, sensitivity=10.0, iterations=1,
                               diffusion_iterations=4, smoother=0, cycle=0, time_step=0.1, tolerance=1e-10, iterations_per_grid=3, out_dtype=np.int16)
    dd = np.abs(out.astype(int) - want.astype(int))
    assert dd.max() <= 1 and (dd != 0).mean() < 1e-3
    assert np.abs(out.astype(int) - vol.astype(int)).max() > 5
