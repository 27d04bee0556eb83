"""The C-ABI layer of the VED front-end (multigridanisotropicdiffusion_b200/csrc/ved.cu: context, staging, chunking, call-sequence
checks, statistics, the GenerateData loop of madved_run) driven on the CPU.

tests/ved_cabi_host.cpp compiles ved.cu UNMODIFIED for the host: kernels on host fibres (tests/mad_host/fiber_shim.h), CUDA runtime
calls on host memory (tests/fake_cuda/cuda_runtime.h), and the solver entry points madved_run calls answered by a stand-in backed by
the oracle.  The resulting tests/_build/libmadved_host.so exports the same madved_* symbols as libmadgpu.so, and the product's own
Python binding (multigridanisotropicdiffusion_b200.ved.MadVed) is pointed at it for the duration of a test, so these tests read like
tests/test_gpu_ved.py.  This is a CPU check of the code, not a CPU path of the product: libmadgpu.so still refuses to run without a
GPU (tests/test_cpu_host.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import ved as V
from util import ROOT, load_ved_test, random_image, rel_l2

VED_TEST = dict(alpha=0.5, beta=0.5, gamma=5.0, epsilon=0.01, omega=1.5, sensitivity=10.0)  # test/itkVEDTest_GS.cxx:82-99


@pytest.fixture(scope="module")
def hostlib():
    csrc = os.path.join(ROOT, "multigridanisotropicdiffusion_b200", "csrc")
    src = os.path.join(ROOT, "tests", "ved_cabi_host.cpp")
    deps = [src, os.path.join(ROOT, "tests", "mad_host", "fiber_shim.h"), os.path.join(ROOT, "tests", "fake_cuda", "cuda_runtime.h"), os.path.join(ROOT, "tests", "fake_cuda", "cuda_runtime.h"),
            os.path.join(csrc, "ved.cu"), os.path.join(csrc, "ved_kernels.cuh"), os.path.join(csrc, "ved_math.h"),
            os.path.join(ROOT, "include", "madved.h"), os.path.join(ROOT, "include", "madgpu.h")]
    out = os.path.join(ROOT, "tests", "_build", "libmadved_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    from oracle import oracle as O
    O.build()
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-Wall", "-Wno-unknown-pragmas", "-fno-gnu-unique", "-x", "c++",
                               "-I" + os.path.join(ROOT, "tests", "fake_cuda"), "-I" + os.path.join(ROOT, "tests"), "-o", out, src,
                               "-L" + os.path.join(ROOT, "oracle"), "-lmadoracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    from multigridanisotropicdiffusion_b200 import _lib as B
    real = B.load()
    L = C.CDLL(out)
    for name in B.EXPORTS:
        if name.startswith("madved_"):
            f, r = getattr(L, name), getattr(real, name)
            f.argtypes, f.restype = r.argtypes, r.restype
    L.fake_solver_create.restype = C.c_void_p
    L.fake_solver_create.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_double, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int]
    L.fake_solver_destroy.argtypes = [C.c_void_p]
    L.fake_solver_solves.argtypes = [C.c_void_p]
    L.fake_cuda_live_allocs.restype = C.c_longlong
    return L


@pytest.fixture
def MadVed(hostlib, monkeypatch):
    """The product's MadVed class bound to the host build of ved.cu."""
    from multigridanisotropicdiffusion_b200 import _lib as B
    from multigridanisotropicdiffusion_b200 import ved
    monkeypatch.setattr(B, "_lib", hostlib)
    return ved.MadVed


class FakeSolver:
    """What MadVed.run needs of a MadSolver, on the oracle-backed stand-in."""

    def __init__(self, lib, shape, sp, time_step=0.1, smoother=0, nu=2, cycle=0, tolerance=1e-6, steps=1):
        self._lib, self.shape, self.last_stats = lib, tuple(shape), None
        self._ctx = C.c_void_p(lib.fake_solver_create((C.c_int * 3)(*shape[::-1]), (C.c_double * 3)(*sp), time_step, smoother, nu, cycle,
                                                      tolerance, steps))

    def _stats(self, st):
        self.last_stats = dict(steps=st.steps, cycles_per_step=list(st.cycles_per_step)[: st.steps], kernel_launches=st.kernel_launches)

    def close(self):
        self._lib.fake_solver_destroy(self._ctx)


def _sub_volume():
    img, sp = load_ved_test()
    return np.ascontiguousarray(img[22:38, 26:44, 20:46]), sp  # (16, 18, 26)


@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.float32, np.float64])
def test_hessian_through_the_cabi(MadVed, dtype):
    shape, sp = (7, 9, 37), (0.5, 0.4, 0.8)
    img = np.clip(random_image(shape, seed=2), 0, 250).astype(dtype)
    with MadVed(shape, sp) as v:
        v.set_image(img)
        v.hessian(0.775)
        H = v.get_hessian()
        st = v.stats()
    assert st["kernel_launches"] == 10 + (0 if dtype == np.float32 else 1) and st["hessian_ms"] > 0
    Ho = V.hessian(img.astype(np.float32).astype(np.float64), sp, 0.775)
    for k in range(6):
        assert np.abs(H[..., k] - Ho[..., k]).max() / np.abs(Ho[..., k]).max() < 5e-6


def test_update_on_host_hessians_in_chunks(MadVed, monkeypatch):
    """madved_update_vesselness_host_f64 and the AoS getters with a staging chunk smaller than the volume (ragged last chunk)."""
    monkeypatch.setenv("MADVED_STAGE_VOXELS", "1000")
    img, sp = _sub_volume()  # 7488 voxels
    hs = [V.hessian(img.astype(np.float64), sp, s) for s in V.DEFAULT_SCALES]
    To, st = V.ved_tensor(img, sp, hessians=hs, **VED_TEST)
    with MadVed(img.shape, sp, **VED_TEST) as v:
        for H in hs:
            v.update_vesselness(H)
        resp, T = v.get_response(), v.get_tensor()
        assert v.stats()["scales"] == 5 and v.stats()["kernel_launches"] == 5 * 8
    np.testing.assert_allclose(resp, st.response, rtol=1e-11, atol=1e-300)
    np.testing.assert_allclose(T, To, atol=2e-7)
    assert (st.response > 0).mean() > 0.05


def test_first_scale_rule_and_begin(MadVed):
    rng = np.random.default_rng(3)
    shape = (6, 7, 8)
    hs = [rng.normal(size=shape + (6,)) * s for s in (1.0, 3.0, 0.2)]
    To, st = V.ved_tensor(np.zeros(shape), (1, 1, 1), scales=(1, 2, 3), hessians=hs)
    To2, st2 = V.ved_tensor(np.zeros(shape), (1, 1, 1), scales=(1,), hessians=hs[2:])
    with MadVed(shape, (1, 1, 1)) as v:
        for H in hs:
            v.update_vesselness(H)
        np.testing.assert_allclose(v.get_response(), st.response, rtol=1e-11, atol=1e-300)
        np.testing.assert_allclose(v.get_tensor(), To, atol=1e-6)
        v.begin()
        v.update_vesselness(hs[2])
        np.testing.assert_allclose(v.get_response(), st2.response, rtol=1e-11, atol=1e-300)
        np.testing.assert_allclose(v.get_tensor(), To2, atol=1e-6)
        v.set_params(0.5, 0.5, 5.0, 0.01, 1.5, 10.0)  # omega changes the tensor of the next pass
        v.begin()
        v.update_vesselness(hs[0])
        To3, _ = V.ved_tensor(np.zeros(shape), (1, 1, 1), scales=(1,), hessians=hs[:1], omega=1.5)
        np.testing.assert_allclose(v.get_tensor(), To3, atol=1e-6)


def test_whole_front_end_and_device_tensor_planes(MadVed):
    img, sp = _sub_volume()
    with MadVed(img.shape, sp, **VED_TEST) as v:
        v.set_image(img)
        for s in V.DEFAULT_SCALES:
            v.add_scale(s)
        T, resp = v.get_tensor(), v.get_response()
        planes = v.tensor_planes()
        n = img.size
        raw = np.stack([np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n,)).copy() for p in planes], axis=-1)
        assert v.stats()["scales"] == 5 and v.stats()["kernel_launches"] == 1 + 5 * 11
        assert v.image_device() != 0
    np.testing.assert_array_equal(raw.reshape(img.shape + (6,)), T.astype(np.float32))  # what the solver would ingest
    To, so = V.ved_tensor(img.astype(np.float64), sp, **VED_TEST)
    bad = np.abs(T - To).max(axis=-1) > 1e-3
    assert bad.mean() < 1e-3 and rel_l2(T, To) < 1e-4
    np.testing.assert_allclose(resp[~bad], so.response[~bad], rtol=1e-3, atol=1e-9)


@pytest.mark.parametrize("pixel,iterations,cycle,smoother", [(np.float64, 2, 0, 0), (np.int16, 1, 1, 1), (np.float32, 0, 0, 0)])
def test_run_is_generate_data(MadVed, hostlib, pixel, iterations, cycle, smoother):
    """madved_run = VEDMultigridImageFilter::GenerateData (hxx:63-155): per outer iteration a fresh vesselness state, the tensor to
    the solver in place, the image solved in place, the result cast from the solver's fp64 iterate."""
    img, sp = _sub_volume()
    img = img.astype(pixel)
    kw = dict(iterations=iterations, diffusion_iterations=2, smoother=smoother, cycle=cycle, time_step=0.1, tolerance=1e-8, iterations_per_grid=2,
              **VED_TEST)
    want, info = V.ved_filter(img, sp, V.DEFAULT_SCALES, out_dtype=pixel, **kw)
    s = FakeSolver(hostlib, img.shape, sp, 0.1, smoother, 2, cycle, 1e-8, 2)
    try:
        with MadVed(img.shape, sp, **VED_TEST) as v:
            out = v.run(s, img, V.DEFAULT_SCALES, iterations=iterations)
            st = v.stats()
        assert hostlib.fake_solver_solves(s._ctx) == iterations
        assert out.dtype == pixel and st["scales"] == 5 * iterations
        if iterations:
            assert s.last_stats["cycles_per_step"] == info["cycles"][-1]
            # launches: the input cast (none for float), 11 per scale, 1000 per stand-in solve
            assert st["kernel_launches"] == (0 if pixel == np.float32 else 1) + iterations * (55 + 1000)
        if pixel == np.int16:
            d = np.abs(out.astype(int) - want.astype(int))
            assert d.max() <= 1 and (d != 0).mean() < 1e-3
        else:
            assert rel_l2(out, want) < 2e-6  # fp32 image between the outer iterations, fp32 tensor
    finally:
        s.close()


def test_call_sequence_errors_and_validation(MadVed, hostlib):
    from multigridanisotropicdiffusion_b200 import MadGpuError
    shape, sp = (8, 8, 8), (1, 1, 1)
    with MadVed(shape, sp) as v:
        for call in (lambda: v.hessian(1.0), v.update_vesselness, v.tensor_planes, v.get_tensor, v.get_response, v.get_hessian):
            with pytest.raises(MadGpuError):
                call()  # nothing to work on yet
        v.set_image(random_image(shape))
        with pytest.raises(MadGpuError):
            v.hessian(0.0)
        with pytest.raises(MadGpuError):
            v.hessian(float("nan"))
        with pytest.raises(MadGpuError):
            v.update_vesselness()  # still no Hessian
        v.hessian(1.0)
        v.set_image(random_image(shape, seed=3))  # a new image invalidates the Hessian
        with pytest.raises(MadGpuError):
            v.update_vesselness()
        s = FakeSolver(hostlib, (8, 8, 16), sp)
        try:
            rc = hostlib.madved_run(v._ctx, s._ctx, 2, random_image(shape).ctypes.data_as(C.c_void_p), 2, np.empty(shape, np.float32).ctypes.data_as(C.c_void_p),
                                    (C.c_double * 1)(1.0), 1, 1, None)
            assert rc == -1 and b"different volume size" in hostlib.madved_last_error(v._ctx)
        finally:
            s.close()
    with pytest.raises(MadGpuError):
        MadVed((8, 8, 3), sp)


def test_contexts_release_their_memory_and_need_a_device(MadVed, hostlib):
    from multigridanisotropicdiffusion_b200 import MadGpuError
    before = hostlib.fake_cuda_live_allocs()
    img, sp = _sub_volume()
    with MadVed(img.shape, sp) as v:
        v.set_image(img)
        v.add_scale(1.0)
        v.get_tensor()
        assert hostlib.fake_cuda_live_allocs() == before + 1 + 12 + 6 + 1 + 1  # image, work, tensor, response, staging
    assert hostlib.fake_cuda_live_allocs() == before
    hostlib.fake_cuda_set_devices(0)
    try:
        with pytest.raises(MadGpuError) as e:
            MadVed(img.shape, sp)
        assert "no CUDA device" in str(e.value)
    finally:
        hostlib.fake_cuda_set_devices(1)
