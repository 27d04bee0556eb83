"""CPU tests of the VED tensor front-end.

1. The oracle's restatement of the third-party pieces (oracle/ved_oracle.c: recursive Gaussian, Hessian, eigen-solver) against
   independent implementations: sampled-Gaussian convolution (scipy), polynomial known answers, LAPACK.  This BOUNDS
   restatement errors; it does not pin ITK's filter (parity unpinned, see the header of ved_oracle.c).
2. The source of the CUDA kernels themselves, compiled for the host by tests/ved_host_harness.cpp, against the oracle:
   csrc/ved_math.h (coefficient set-up, line recursion with fp32 intermediate storage, eigen-solver, vesselness, per-voxel update)
   and csrc/ved_kernels.cuh -- the unmodified __global__ kernels with their launch geometry and the pass structure of the
   separable Hessian, run on host fibres through tests/mad_host/fiber_shim.h (threadIdx / __shared__ / __syncwarp / launch), so
   indexing, the shared-memory tile walk of the x pass, ragged edges and buffer reuse are exercised as written.
   The tolerances found here are the ones tests/test_gpu_ved.py uses.
No GPU, no compute call into libmadgpu.so; the context / C-ABI / copy layer of ved.cu is what only the GPU tests reach.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import ved as V
from util import ROOT, load_ved_test, rel_l2

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
VED_TEST = dict(alpha=0.5, beta=0.5, gamma=5.0, epsilon=0.01, omega=1.5, sensitivity=10.0)  # test/itkVEDTest_GS.cxx:82-99


@pytest.fixture(scope="module")
def host():
    """ved_math.h + ved_kernels.cuh compiled for the host."""
    csrc = os.path.join(ROOT, "multigridanisotropicdiffusion_b200", "csrc")
    src = os.path.join(ROOT, "tests", "ved_host_harness.cpp")
    deps = [src, os.path.join(ROOT, "tests", "mad_host", "fiber_shim.h"), os.path.join(ROOT, "tests", "fake_cuda", "cuda_runtime.h"), os.path.join(csrc, "ved_math.h"), os.path.join(csrc, "ved_kernels.cuh")]
    out = os.path.join(ROOT, "tests", "_build", "libvedhost.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-Wall", "-Wno-unknown-pragmas", "-fno-gnu-unique", "-I" + os.path.join(ROOT, "tests", "fake_cuda"), "-I" + os.path.join(ROOT, "tests"), "-o", out, src])
    L = C.CDLL(out)
    L.vh_rg_setup.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int, _dp]
    L.vh_rg_line.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int, _fp, C.c_int, _fp, C.c_double]
    L.vh_hessian.argtypes = [_ip, _dp, C.c_double, _fp, _fp]
    L.vh_hessian.restype = C.c_int
    L.vh_eig3_top.argtypes = [_dp, _dp, _dp]
    L.vh_vesselness.restype = C.c_double
    L.vh_vesselness.argtypes = [_dp, C.c_double, C.c_double, C.c_double]
    L.vh_update.argtypes = [C.c_longlong, C.c_int, _fp, _dp, C.c_int, _dp, _dp, _fp, C.c_longlong]
    L.vh_cast_in_i16.argtypes = [C.c_void_p, _fp, C.c_longlong]
    L.vh_cast_in_u8.argtypes = [C.c_void_p, _fp, C.c_longlong]
    L.vh_cast_in_f64.argtypes = [C.c_void_p, _fp, C.c_longlong]
    L.vh_planes_to_aos.argtypes = [_fp, C.c_longlong, _dp, C.c_longlong, C.c_longlong]
    return L


def _d(a):
    return a.ctypes.data_as(_dp)


def _f(a):
    return a.ctypes.data_as(_fp)


def _sub_volume(dtype=np.float64):
    img, sp = load_ved_test()
    return np.ascontiguousarray(img[20:44, 24:52, 18:48].astype(dtype)), sp


# ---------------------------------------------------------------------------------------------------- oracle, third-party pieces
@pytest.mark.parametrize("sigmad", [0.6, 0.96, 2.0, 6.4])
def test_oracle_recursive_gaussian_known_answers(sigmad):
    """Constant -> 1 / 0 / 0, ramp -> . / 1 / 0, x^2/2 -> . / . / 1 away from the borders (the normalisations alpha0..2)."""
    n = 300
    x = np.arange(n, dtype=np.float64)
    mid = slice(n // 2 - 3, n // 2 + 3)
    for order, poly, expect in ((0, np.ones(n), 1.0), (1, np.ones(n), 0.0), (2, np.ones(n), 0.0), (1, x, 1.0), (2, x, 0.0), (2, 0.5 * x * x, 1.0)):
        y = V.rg_filter_line(V.rg_coefs(sigmad, 1.0, order, False), poly)
        np.testing.assert_allclose(y[mid], expect, atol=2e-9 * max(1.0, np.abs(poly).max()))
    # border handling: the edge sample extends to infinity, so a constant stays a constant up to the border
    y = V.rg_filter_line(V.rg_coefs(sigmad, 1.0, 0, False), np.full(n, 7.0))
    np.testing.assert_allclose(y, 7.0, atol=1e-10)


@pytest.mark.parametrize("sigmad", [0.96, 2.0, 4.0, 6.4])
def test_oracle_recursive_gaussian_vs_sampled_gaussian(sigmad):
    """Deriche's 4th-order recursions approximate the Gaussian and its derivatives to about 1 % (published accuracy)."""
    ndi = pytest.importorskip("scipy.ndimage")
    rng = np.random.default_rng(1)
    sig = rng.normal(size=600)
    for order, tol in ((0, 6e-3), (1, 8e-3), (2, 2.5e-2)):
        y = V.rg_filter_line(V.rg_coefs(sigmad, 1.0, order, False), sig)
        g = ndi.gaussian_filter1d(sig, sigmad, order=order, mode="nearest", truncate=8)
        assert rel_l2(y[60:-60], g[60:-60]) < tol


def test_oracle_scale_normalisation_and_spacing():
    """NormalizeAcrossScale multiplies the k-th derivative by sigma^k (physical); the Hessian divides by the spacings."""
    n = 200
    sp, sigma = 0.5, 1.5
    x = np.arange(n) * sp
    y2 = V.rg_filter_line(V.rg_coefs(sigma, sp, 2, True), 0.5 * x * x)
    assert abs(y2[n // 2] / sp ** 2 - sigma ** 2) < 1e-8  # sigma^2 * d2/dx2 (x^2/2) = sigma^2
    y1 = V.rg_filter_line(V.rg_coefs(sigma, sp, 1, True), x)
    assert abs(y1[n // 2] / sp - sigma) < 1e-9
    # whole Hessian on a quadratic form q = 1/2 x^T A x: sigma^2 * A in the interior
    A = np.array([[2.0, 0.3, -0.4], [0.3, 1.0, 0.25], [-0.4, 0.25, -1.5]])
    h = (0.5, 0.4, 0.8)
    zz, yy, xx = np.meshgrid(np.arange(40) * h[2], np.arange(44) * h[1], np.arange(48) * h[0], indexing="ij")
    P = np.stack([xx, yy, zz], -1)
    q = 0.5 * np.einsum("...i,ij,...j->...", P, A, P)
    H = V.hessian(q, h, 1.0)
    want = [A[0, 0], A[0, 1], A[0, 2], A[1, 1], A[1, 2], A[2, 2]]
    np.testing.assert_allclose(H[20, 22, 24], want, atol=1e-4)  # the tails of the border discontinuity reach this far


def test_oracle_eig3_vs_lapack():
    rng = np.random.default_rng(0)
    for _ in range(500):
        a = rng.normal(size=6) * rng.choice([1e-3, 1.0, 100.0])
        w, Q = V.eig3(a)
        A = np.array([[a[0], a[1], a[2]], [a[1], a[3], a[4]], [a[2], a[4], a[5]]])
        w2 = np.linalg.eigvalsh(A)
        s = np.abs(w2).max()
        assert np.all(np.diff(w) >= 0)
        np.testing.assert_allclose(w, w2, atol=1e-14 * s)
        np.testing.assert_allclose(A @ Q, Q * w, atol=1e-14 * s)
        np.testing.assert_allclose(Q.T @ Q, np.eye(3), atol=1e-14)


def test_oracle_vesselness_closed_forms():
    assert V.vesselness([0.0, 1.0, -2.0]) == 0.0 and V.vesselness([0.0, -1.0, 2.0]) == 0.0  # hxx:183-186
    e = np.array([0.0, -3.0, -3.0])  # ideal tube: RA = 1, RB = 0
    want = np.exp(-2e-10 / 27.0) * (1 - np.exp(-1 / 0.5)) * 1.0 * (1 - np.exp(-18 / 50.0))
    assert abs(V.vesselness(e) - want) < 1e-15
    assert V.vesselness([-3.0, -3.0, -3.0]) < V.vesselness(e)  # a blob scores lower than a tube


def test_oracle_tensor_is_identity_off_vessels_and_anisotropic_on_them():
    img, sp = _sub_volume()
    T, st = V.ved_tensor(img, sp, **VED_TEST)
    off = st.response <= 0
    assert off.any() and (~off).any()
    np.testing.assert_array_equal(T[off], np.array([1.0, 0, 0, 1, 0, 1]) * np.ones((off.sum(), 6)))
    tr = T[..., 0] + T[..., 3] + T[..., 5]
    Vs = st.response[~off] ** 0.1
    np.testing.assert_allclose(tr[~off], 3 + (2 * 0.01 + 1.5 - 3) * Vs, rtol=1e-12)  # trace = 2a + b


# ---------------------------------------------------------------------------------------------------- device arithmetic on the host
@pytest.mark.parametrize("sigma,spacing", [(0.3, 0.5), (0.3, 0.3125), (0.775, 0.3125), (2.0, 0.3125), (2.0, 0.5)])
def test_device_coefficients_match_oracle(host, sigma, spacing):
    names = ["N0", "N1", "N2", "N3", "D1", "D2", "D3", "D4", "M1", "M2", "M3", "M4"]
    for order in (0, 1, 2):
        o = V.rg_coefs(sigma, spacing, order, True)
        d = np.empty(14)
        host.vh_rg_setup(sigma, spacing, order, 1, _d(d))
        want = np.array([getattr(o, k) for k in names])
        np.testing.assert_allclose(d[:12], want, rtol=1e-11, atol=1e-13 * np.abs(want).max())
        SD = 1 + o.D1 + o.D2 + o.D3 + o.D4
        # steady-state responses = ITK's boundary coefficients BN_i / D_i, BM_i / D_i
        assert abs(d[12] - o.BN1 / o.D1) < 1e-11 * max(1.0, abs(d[12])) and abs(d[13] - o.BM1 / o.D1) < 1e-11 * max(1.0, abs(d[13]))
        assert abs(d[12] - (o.N0 + o.N1 + o.N2 + o.N3) / SD) < 1e-11 * max(1.0, abs(d[12]))


@pytest.mark.parametrize("n", [4, 5, 7, 33, 64, 69, 200])
def test_device_line_recursion_matches_oracle(host, n):
    """State-machine form with steady-state start == ITK's explicit border formulas; fp32 storage of the causal half."""
    rng = np.random.default_rng(n)
    x = (100 + 20 * rng.normal(size=n)).astype(np.float32)
    for sigma, sp in ((0.3, 0.3125), (1.245, 0.5), (2.0, 0.3125)):
        for order in (0, 1, 2):
            want = V.rg_filter_line(V.rg_coefs(sigma, sp, order, True), x.astype(np.float64)) * 0.37
            y = np.empty(n, dtype=np.float32)
            host.vh_rg_line(sigma, sp, order, 1, _f(x), n, _f(y), 0.37)
            # one fp32 rounding of the causal half (|c| <~ |x| * a few) and one of the result
            np.testing.assert_allclose(y, want, atol=4e-5, rtol=2e-6)


@pytest.fixture(scope="module")
def hessians(host):
    """(oracle fp64 Hessians, kernel-order fp32 Hessians) of the ved_test sub-volume at the five scales."""
    img, sp = _sub_volume()
    img32 = img.astype(np.float32)
    n = (C.c_int * 3)(*img.shape[::-1])
    h = (C.c_double * 3)(*sp)
    ho, hk = [], []
    for s in V.DEFAULT_SCALES:
        ho.append(V.hessian(img, sp, s))
        H = np.empty((6,) + img.shape, dtype=np.float32)
        host.vh_hessian(n, h, s, _f(img32), _f(H))
        hk.append(H)
    return img, sp, ho, hk


def test_device_hessian_pass_structure_matches_oracle(hessians):
    """The kernels (k_rg_rows, k_rg_lines) with shared x / y passes, fp32 intermediates, per-component scaling:
    <= 2e-6 of the component's range at every scale."""
    _, _, ho, hk = hessians
    for s, Ho, Hk in zip(V.DEFAULT_SCALES, ho, hk):
        for k in range(6):
            err = np.abs(Hk[k] - Ho[..., k]).max() / np.abs(Ho[..., k]).max()
            assert err < 2e-6, (s, k, err)


def test_device_eig3_top_matches_oracle(host):
    rng = np.random.default_rng(5)
    for i in range(3000):
        a = rng.normal(size=6) * rng.choice([1e-6, 1e-2, 1.0, 300.0])
        if i % 7 == 0:
            a[[1, 2, 4]] = 0.0  # already diagonal
        if i % 11 == 0:
            a[[1, 2, 4]] *= 1e-9  # nearly diagonal
        w, t = np.empty(3), np.empty(3)
        host.vh_eig3_top(_d(np.ascontiguousarray(a)), _d(w), _d(t))
        wo, Q = V.eig3(a)
        s = max(np.abs(wo).max(), 1e-300)
        np.testing.assert_allclose(w, wo, atol=4e-15 * s)
        assert abs(np.linalg.norm(t) - 1) < 1e-14
        gap = (wo[2] - wo[1]) / s
        if gap > 1e-6:  # the eigenvector is defined up to its sign and conditioned by the gap
            assert min(np.abs(t - Q[:, 2]).max(), np.abs(t + Q[:, 2]).max()) < 1e-14 / gap
    w, t = np.empty(3), np.empty(3)
    host.vh_eig3_top(_d(np.zeros(6)), _d(w), _d(t))  # zero matrix: no rotation, no NaN
    assert np.all(w == 0) and abs(np.linalg.norm(t) - 1) < 1e-15


def test_device_vesselness_matches_oracle(host):
    rng = np.random.default_rng(6)
    for _ in range(2000):
        e = rng.normal(size=3) * rng.choice([1e-4, 0.1, 1.0, 30.0])
        e = np.ascontiguousarray(e[np.argsort(np.abs(e))])
        if rng.random() < 0.6:
            e[1], e[2] = -abs(e[1]), -abs(e[2])
        assert host.vh_vesselness(_d(e), 0.5, 0.5, 5.0) == V.vesselness(e, 0.5, 0.5, 5.0)


def _update_all(host, shape, hs, soa, params, chunk=0):
    nvox = int(np.prod(shape))
    resp = np.empty(nvox)
    T = np.empty((6, nvox), dtype=np.float32)
    p = np.array(params, dtype=np.float64)
    for i, H in enumerate(hs):
        if soa:
            H = np.ascontiguousarray(H, dtype=np.float32)
            host.vh_update(nvox, 1, _f(H), None, int(i == 0), _d(p), _d(resp), _f(T), nvox)
        else:
            H = np.ascontiguousarray(H, dtype=np.float64)
            host.vh_update(nvox, 0, None, _d(H), int(i == 0), _d(p), _d(resp), _f(T), chunk or nvox)
    return resp.reshape(shape), np.moveaxis(T.reshape((6,) + tuple(shape)), 0, -1)


PARAMS = [VED_TEST[k] for k in ("alpha", "beta", "gamma", "epsilon", "omega", "sensitivity")]


def test_device_update_on_identical_hessians_matches_oracle(host, hessians):
    """Same fp64 Hessians on both sides (the madved_update_vesselness_host_f64 path): the response agrees to rounding, the
    folded tensor a I + (b - a) t t^T equals Q D Q^T to fp32 rounding -- same arg-max scale everywhere."""
    img, sp, ho, _ = hessians
    resp, T = _update_all(host, img.shape, ho, False, PARAMS)
    To, st = V.ved_tensor(img, sp, hessians=ho, **VED_TEST)
    np.testing.assert_allclose(resp, st.response, rtol=1e-12, atol=1e-300)
    np.testing.assert_allclose(T, To, atol=2e-7)


def test_device_update_random_hessians_and_first_scale_rule(host):
    rng = np.random.default_rng(3)
    shape = (6, 7, 8)
    hs = [rng.normal(size=shape + (6,)) * s for s in (1.0, 3.0, 0.2)]
    params = [0.5, 0.5, 5.0, 0.01, 5.0, 10.0]
    resp, T = _update_all(host, shape, hs, False, params, chunk=100)  # 336 voxels in chunks of 100: ragged last chunk
    To, st = V.ved_tensor(np.zeros(shape), (1, 1, 1), scales=(1, 2, 3), hessians=hs)
    np.testing.assert_allclose(resp, st.response, rtol=1e-12, atol=1e-300)
    np.testing.assert_allclose(T, To, atol=1e-6)


def test_device_pipeline_fp32_hessians_vs_oracle(host, hessians):
    """The whole device pipeline (kernel-order fp32 Hessians -> update) against the fp64 oracle.  fp32 rounding of the Hessian
    can flip the arg-max between two scales whose vesselness is nearly tied; those voxels get the other scale's tensor.
    The bound used by the GPU tests: rel-L2 of the tensor field <= 2e-3 and <= 0.5 % of the voxels off by more than 1e-3."""
    img, sp, ho, hk = hessians
    resp, T = _update_all(host, img.shape, [np.ascontiguousarray(H) for H in hk], True, PARAMS)
    To, st = V.ved_tensor(img, sp, hessians=ho, **VED_TEST)
    bad = np.abs(T - To).max(axis=-1) > 1e-3
    assert bad.mean() < 5e-3, bad.mean()
    assert rel_l2(T, To) < 2e-3
    np.testing.assert_allclose(resp[~bad], st.response[~bad], rtol=2e-3, atol=1e-9)


# shapes: nx a multiple of 32, ragged last tile, nx < 32, minimum line length on every axis, rows not a multiple of 32, one row tile
@pytest.mark.parametrize("shape,sp", [((6, 7, 64), (1.0, 1.0, 1.0)), ((9, 7, 69), (0.5, 0.4, 0.8)), ((12, 14, 30), (0.3125, 0.3125, 0.5)),
                                      ((4, 5, 4), (1.0, 1.0, 1.0)), ((5, 13, 97), (0.33, 0.33, 0.33)), ((4, 4, 33), (1.0, 2.0, 0.5))])
def test_device_hessian_kernels_on_awkward_shapes(host, shape, sp):
    """Every output voxel is written (the work volumes are poisoned), 10 launches per scale, Hessian within 5e-6 of its range."""
    img = (100 + 20 * np.random.default_rng(sum(shape)).normal(size=shape)).astype(np.float32)
    for sigma in (0.3, 1.245):
        H = np.empty((6,) + shape, dtype=np.float32)
        launches = host.vh_hessian((C.c_int * 3)(*shape[::-1]), (C.c_double * 3)(*sp), sigma, _f(img), _f(H))
        assert launches == 10
        Ho = V.hessian(img.astype(np.float64), sp, sigma)
        for k in range(6):
            err = np.abs(H[k] - Ho[..., k]).max() / np.abs(Ho[..., k]).max()
            assert err < 5e-6, (shape, sigma, k, err)


def test_device_cast_and_layout_kernels(host):
    rng = np.random.default_rng(9)
    n = 1000  # not a multiple of the block size
    for name, a in (("i16", rng.integers(-3000, 3000, n).astype(np.int16)), ("u8", rng.integers(0, 255, n).astype(np.uint8)),
                    ("f64", rng.normal(size=n) * 1e3)):
        out = np.full(n, np.nan, dtype=np.float32)
        getattr(host, "vh_cast_in_" + name)(a.ctypes.data_as(C.c_void_p), _f(out), n)
        np.testing.assert_array_equal(out, a.astype(np.float32))
    planes = rng.normal(size=(6, n)).astype(np.float32)
    out = np.full((300, 6), np.nan)
    host.vh_planes_to_aos(_f(planes), n, _d(out), 650, 300)  # a chunk in the middle
    np.testing.assert_array_equal(out, planes[:, 650:950].T.astype(np.float64))
