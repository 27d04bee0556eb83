"""Builds (when stale) and loads tests/_build/libmadgpu_host.so -- csrc/madgpu.cu compiled unmodified for the host, see
tests/mad_host/madgpu_host.cpp -- and points the product's own Python binding at it.  TEST INFRASTRUCTURE."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(ROOT, "tests", "_build")
HOST_LIB = os.path.join(OUT, "libmadgpu_host.so")
FAKE_NCCL = os.path.join(OUT, "libfakenccl.so")


def build():
    csrc = os.path.join(ROOT, "multigridanisotropicdiffusion_b200", "csrc")
    deps = [os.path.join(HERE, f) for f in ("madgpu_host.cpp", "fiber_shim.h", "fake_nccl.cpp", "build.sh")]
    deps += [os.path.join(csrc, f) for f in ("madgpu.cu", "mad_kernels.cuh", "mad_fast.cuh", "mad_fast2d.cuh", "ved.cu", "ved_kernels.cuh", "ved_math.h")]
    deps += [os.path.join(ROOT, "tests", "fake_cuda", f) for f in ("cuda_runtime.h", "cuda_fp16.h")]
    deps += [os.path.join(ROOT, "include", "madgpu.h"), os.path.join(ROOT, "include", "madved.h")]
    newest = max(os.path.getmtime(d) for d in deps)
    if not (os.path.exists(HOST_LIB) and os.path.exists(FAKE_NCCL)) or min(os.path.getmtime(HOST_LIB), os.path.getmtime(FAKE_NCCL)) < newest:
        subprocess.check_call([os.path.join(HERE, "build.sh")], stdout=subprocess.DEVNULL)
    return HOST_LIB


def load():
    """The host build with the argtypes of the real binding; call bind() to make MadSolver use it."""
    from multigridanisotropicdiffusion_b200 import _lib as B
    build()
    # the device-side Gauss-Jordan inverse of the coarsest operator is ~2 n kernel launches of n x 2n threads: minutes on fibres for
    # the 512-unknown grids most tests end with.  The emulated runs keep the host LU (same inverse to rounding) unless a test asks
    # for the device path explicitly (tests/test_cpu_mad_host.py::test_device_inverse_of_the_coarsest_operator).
    os.environ.setdefault("MADGPU_COARSE_HOST", "1")
    real = B.load()
    L = C.CDLL(HOST_LIB)
    for name in B.EXPORTS:
        if hasattr(L, name):
            f, r = getattr(L, name), getattr(real, name)
            f.argtypes, f.restype = r.argtypes, r.restype
    for n in ("mad_host_launches", "mad_host_switches", "mad_host_live_allocs"):
        getattr(L, n).restype = C.c_longlong
    return L


def bind(lib):
    """Point multigridanisotropicdiffusion_b200._lib at the host build (process-wide; the slab emulation runs in its own process,
    the single-rank tests restore the binding with monkeypatch)."""
    from multigridanisotropicdiffusion_b200 import _lib as B
    B._lib = lib
