"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol include/madgpu.h and
include/madved.h declare, parameter defaults mirror the reference's constructors, and -- there being no CPU fallback --
every compute entry point fails loudly without a GPU.  No CUDA work is done here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    names = set()
    for hdr in ("madgpu.h", "madved.h"):
        src = open(os.path.join(ROOT, "include", hdr)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(mad(?:gpu|ved)_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_header_declares_the_documented_surface():
    names = _header_functions()
    for must in ("madgpu_create", "madgpu_destroy", "madgpu_set_tensor_f32", "madgpu_set_tensor_f64", "madgpu_solve_u8",
                 "madgpu_solve_i16", "madgpu_solve_f32", "madgpu_solve_f64", "madgpu_last_error", "madgpu_op_vcycle",
                 "madved_create", "madved_hessian", "madved_update_vesselness", "madved_update_vesselness_host_f64", "madved_run"):
        assert must in names
    assert len(names) >= 50


def test_library_exports_every_declared_symbol():
    from multigridanisotropicdiffusion_b200 import _lib
    lib = _lib.load()
    missing = [n for n in _header_functions() if not hasattr(lib, n)]
    assert not missing, missing
    # and the Python binding knows every one of them
    assert sorted(_lib.EXPORTS) == _header_functions()


def test_defaults_mirror_the_reference_constructor():
    """itkMultigridAnisotropicDiffusionImageFilter.hxx:38-49 and mad/itkMultigridWeightedJacobiSmoother.hxx:186-191."""
    from multigridanisotropicdiffusion_b200 import _lib
    lib = _lib.load()
    p = _lib.Params()
    lib.madgpu_params_default(C.byref(p))
    assert p.struct_size == C.sizeof(_lib.Params)
    assert p.time_step == 0.01 and p.number_of_steps == 1 and p.cycle == _lib.CYCLE_V
    assert p.iterations_per_grid == 2 and p.tolerance == 1e-6 and p.max_cycles == 100 and p.verbose == 0
    assert p.smoother == _lib.SMOOTHER_GS and abs(p.omega - 2.0 / 3.0) < 1e-16
    assert p.world_size == 1 and p.rank == 0


def test_ved_defaults_mirror_the_reference_constructor():
    """itkVEDMultigridImageFilter.hxx:33-58."""
    from multigridanisotropicdiffusion_b200 import _lib
    from multigridanisotropicdiffusion_b200 import ved
    lib = _lib.load()
    p = _lib.VedParams()
    lib.madved_params_default(C.byref(p))
    assert p.struct_size == C.sizeof(_lib.VedParams)
    assert (p.alpha, p.beta, p.gamma, p.epsilon, p.omega, p.sensitivity) == (0.5, 0.5, 5.0, 0.01, 5.0, 10.0)
    assert ved.DEFAULT_SCALES == (0.300, 0.482, 0.775, 1.245, 2.000)
    import multigridanisotropicdiffusion_b200 as M
    v = M.VEDMultigridImageFilter()
    assert (v._alpha, v._beta, v._gamma, v._epsilon, v._omega, v._sensitivity) == (0.5, 0.5, 5.0, 0.01, 5.0, 10.0)
    assert v._scales == [0.300, 0.482, 0.775, 1.245, 2.000] and v._iterations == 1 and v._cycle == v.VCYCLE


def test_ved_argument_validation_needs_no_gpu():
    from multigridanisotropicdiffusion_b200 import _lib
    lib = _lib.load()
    p = _lib.VedParams()
    lib.madved_params_default(C.byref(p))
    ctx = C.c_void_p()
    p.size[0], p.size[1], p.size[2] = 16, 16, 3  # lines shorter than 4 samples: the recursive filter refuses them
    assert lib.madved_create(C.byref(p), C.byref(ctx)) == _lib.EINVAL and not ctx.value
    assert b"4 samples" in lib.madved_last_error(None)
    p.size[2] = 16
    p.spacing[1] = 0.0
    assert lib.madved_create(C.byref(p), C.byref(ctx)) == _lib.EINVAL
    p.spacing[1] = 1.0
    p.struct_size = 8
    assert lib.madved_create(C.byref(p), C.byref(ctx)) == _lib.EINVAL
    assert lib.madved_hessian(None, 1.0) == _lib.EINVAL
    assert lib.madved_run(None, None, 0, None, 0, None, None, 0, 0, None) == _lib.EINVAL


def test_filter_defaults_and_setters():
    import multigridanisotropicdiffusion_b200 as M
    f = M.MultigridAnisotropicDiffusionImageFilter()
    assert (f._time_step, f._number_of_steps, f._cycle, f._iterations_per_grid, f._tolerance, f._max_cycles) == \
        (0.01, 1, f.VCYCLE, 2, 1e-6, 100)
    assert (f.VCYCLE, f.FMG, f.SMOOTHER) == (0, 1, 2)  # enum CycleType, …Filter.h:123
    v = M.VEDMultigridImageFilter()
    assert (v._time_step, v._tolerance, v._diffusion_iterations, v._diffusion_iterations_per_grid) == (0.1, 1e-6, 5, 2)
    with pytest.raises(M.MadGpuError):
        f.Update()  # no input


def test_argument_validation_needs_no_gpu():
    from multigridanisotropicdiffusion_b200 import _lib
    lib = _lib.load()
    p = _lib.Params()
    lib.madgpu_params_default(C.byref(p))
    ctx = C.c_void_p()
    p.dim = 4
    assert lib.madgpu_create(C.byref(p), C.byref(ctx)) == _lib.EINVAL and not ctx.value
    assert b"dim" in lib.madgpu_last_error(None)
    p.dim = 3
    p.size[0], p.size[1], p.size[2] = 16, 16, 2
    assert lib.madgpu_create(C.byref(p), C.byref(ctx)) == _lib.EINVAL
    p.size[2] = 16
    p.struct_size = 8
    assert lib.madgpu_create(C.byref(p), C.byref(ctx)) == _lib.EINVAL
    assert lib.madgpu_solve_f32(None, None, None, None) == _lib.EINVAL
    assert lib.madgpu_num_levels(None) == _lib.EINVAL


def test_no_cpu_fallback():
    """Without a usable sm_100 device the product path must refuse to run (never route through the oracle)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import multigridanisotropicdiffusion_b200 as M
    with pytest.raises(M.MadGpuError) as e:
        M.MadSolver((16, 16, 16))
    assert "CUDA" in str(e.value) or "device" in str(e.value)
    f = M.MultigridAnisotropicDiffusionImageFilter("wj")
    f.SetInput(np.zeros((16, 16), np.float32))
    f.SetDiffusionTensor(np.zeros((16, 16, 3), np.float32))
    with pytest.raises(M.MadGpuError):
        f.Update()
    with pytest.raises(M.MadGpuError):
        M.MadVed((16, 16, 16))
    v = M.VEDMultigridImageFilter()
    v.SetInput(np.zeros((16, 16, 16), np.int16), (0.3, 0.3, 0.5))
    with pytest.raises(M.MadGpuError):
        v.Update()  # whole VED filter: device only


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "multigridanisotropicdiffusion_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in txt.replace("the oracle", "").replace("CPU oracle", "").lower() or fn == "phantom.py" \
                    or "import oracle" not in txt and "from oracle" not in txt and "libmadoracle" not in txt, fn
                assert "from oracle" not in txt and "import oracle" not in txt and "libmadoracle" not in txt and "mad_oracle" not in txt, fn


def test_cmake_packaging_configures(tmp_path):
    """CMakeLists.txt (standalone mode): configures with the CUDA language for sm_100a and declares the library, the drop-in test
    programs and -- where the reference is present -- its nine CTest entries.  MADGPU_TEST_CMAKE_BUILD=1 also builds everything."""
    import shutil
    import subprocess
    if not shutil.which("cmake") or not shutil.which("nvcc"):
        pytest.skip("cmake / nvcc not available")
    b = str(tmp_path / "build")
    r = subprocess.run(["cmake", "-S", ROOT, "-B", b], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    cache = open(os.path.join(b, "CMakeCache.txt")).read()
    assert "CMAKE_CUDA_COMPILER" in cache
    r = subprocess.run(["cmake", "--build", b, "--target", "help"], capture_output=True, text=True)
    assert "madgpu" in r.stdout and "dropin_test" in r.stdout and "ved_dropin_test" in r.stdout
    if os.path.isdir("/root/reference/test"):
        assert "ref_tests_dropin" in r.stdout
        t = subprocess.run(["ctest", "--test-dir", b, "-N"], capture_output=True, text=True).stdout
        assert "Total Tests: 9" in t and "itkVEDTest_GS_FMG" in t and "itk2DDiffusionTest_WJ_S" in t
    if os.environ.get("MADGPU_TEST_CMAKE_BUILD") == "1":
        assert subprocess.run(["cmake", "--build", b, "-j", "4"], capture_output=True, text=True).returncode == 0
        assert os.path.exists(os.path.join(b, "libmadgpu.so"))


def test_every_environment_hook_is_documented():
    """INTEGRATION.md's table lists every MADGPU_* variable the library reads (csrc/*.cu), and nothing the library does not read."""
    import re
    csrc = os.path.join(ROOT, "multigridanisotropicdiffusion_b200", "csrc")
    src = "".join(open(os.path.join(csrc, f)).read() for f in os.listdir(csrc) if f.endswith((".cu", ".cuh", ".h")))
    read = set(re.findall(r'getenv\("(MADGPU_[A-Z0-9_]+)"\)', src))
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    table = doc[doc.index("### Environment hooks"):]
    documented = set(re.findall(r"`(MADGPU_[A-Z0-9_]+)`", table))
    assert read - documented == set(), f"undocumented: {sorted(read - documented)}"
    test_only = {"MADGPU_EMULATED_DEVICE", "MADGPU_FULL_EMULATION", "MADGPU_LARGE_TEST_DRYRUN", "MADGPU_PROPS_TEST_SIZE", "MADGPU_BENCH_PEER8",
                 "MADGPU_ROOT", "MADGPU_REFERENCE"}
    assert documented - read - test_only == set(), f"documented but never read: {sorted(documented - read - test_only)}"
