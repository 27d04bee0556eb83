"""VED tensor front-end on the GPU (include/madved.h) against the oracle (oracle/ved_oracle.c, pinned against
the reference's own VED code by tests/test_oracle_vs_ref_ved.py) and the golden vectors recorded from that code.

Tolerances (measured on the CPU with the kernels' own arithmetic, tests/test_cpu_ved.py): the recursive-Gaussian passes store
fp32 intermediates -> Hessian components within 2e-6 of their range (5e-6 asserted); eigen / vesselness / tensor run in fp64 on
identical Hessians -> response to 1e-12, tensor to fp32 rounding; whole pipeline -> tensor rel-L2 <= 1e-4 with at most 0.1 % of
the voxels choosing another scale.  The filter output obeys the solver's Gauss-Seidel bound (<= 1e-4 rel-L2 on the converged image).
"""
import os

import numpy as np
import pytest

from oracle import ved as V
from util import GOLDEN, load_ved_test, random_image, rel_l2

pytestmark = pytest.mark.gpu

VED_TEST = dict(alpha=0.5, beta=0.5, gamma=5.0, epsilon=0.01, omega=1.5, sensitivity=10.0)  # test/itkVEDTest_GS.cxx:82-99


@pytest.fixture(scope="module")
def M():
    import multigridanisotropicdiffusion_b200 as m
    return m


def _ved(M, shape, sp, **kw):
    p = dict(VED_TEST)
    p.update(kw)
    return M.MadVed(shape, sp, **p)


def _sub_volume():
    img, sp = load_ved_test()
    return np.ascontiguousarray(img[20:44, 24:52, 18:48]), sp


# shapes: nx a multiple of 32, nx with a ragged last tile, nx < 32, the minimum line length on every axis, rows not a multiple of 32
@pytest.mark.parametrize("shape,sp", [((24, 28, 30), (0.3125, 0.3125, 0.5)), ((20, 18, 64), (1.0, 1.0, 1.0)), ((9, 7, 69), (0.5, 0.4, 0.8)),
                                      ((4, 5, 4), (1.0, 1.0, 1.0)), ((33, 31, 97), (0.33, 0.33, 0.33))])
@pytest.mark.parametrize("sigma", [0.3, 0.775, 2.0])
def test_hessian_matches_oracle(M, shape, sp, sigma):
    img = random_image(shape, seed=11)
    with _ved(M, shape, sp) as v:
        v.set_image(img)
        v.hessian(sigma)
        H = v.get_hessian()
    Ho = V.hessian(img.astype(np.float64), sp, sigma)
    for k in range(6):
        err = np.abs(H[..., k] - Ho[..., k]).max() / max(np.abs(Ho[..., k]).max(), 1e-30)
        assert err < 5e-6, (k, err)


@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.float32, np.float64])
def test_input_pixel_types(M, dtype):
    """GenerateData's cast of the input (hxx:70-100)."""
    shape, sp = (12, 14, 40), (1.0, 1.0, 1.0)
    img = np.clip(random_image(shape, seed=2), 0, 250).astype(dtype)
    with _ved(M, shape, sp) as v:
        v.set_image(img)
        v.hessian(1.0)
        H = v.get_hessian()
    Ho = V.hessian(img.astype(np.float32).astype(np.float64), sp, 1.0)
    assert rel_l2(H, Ho) < 2e-6


def test_update_on_host_hessians_matches_oracle(M):
    """madved_update_vesselness_host_f64: identical fp64 Hessians on both sides -> same arg-max scale everywhere."""
    img, sp = _sub_volume()
    hs = [V.hessian(img.astype(np.float64), sp, s) for s in V.DEFAULT_SCALES]
    To, st = V.ved_tensor(img, sp, hessians=hs, **VED_TEST)
    with _ved(M, img.shape, sp) as v:
        for H in hs:
            v.update_vesselness(H)
        resp, T = v.get_response(), v.get_tensor()
        assert v.stats()["scales"] == 5
    np.testing.assert_allclose(resp, st.response, rtol=1e-11, atol=1e-300)
    np.testing.assert_allclose(T, To, atol=2e-7)
    assert (st.response > 0).mean() > 0.05


def test_first_scale_rule_and_begin(M):
    """The first Hessian after begin() is stored unconditionally (hxx:272); begin() forgets the state (hxx:121-123)."""
    rng = np.random.default_rng(3)
    shape = (6, 7, 8)
    hs = [rng.normal(size=shape + (6,)) * s for s in (1.0, 3.0, 0.2)]
    To, st = V.ved_tensor(np.zeros(shape), (1, 1, 1), scales=(1, 2, 3), hessians=hs)
    To2, st2 = V.ved_tensor(np.zeros(shape), (1, 1, 1), scales=(1,), hessians=hs[2:])
    with M.MadVed(shape, (1, 1, 1)) as v:
        for H in hs:
            v.update_vesselness(H)
        np.testing.assert_allclose(v.get_response(), st.response, rtol=1e-11, atol=1e-300)
        np.testing.assert_allclose(v.get_tensor(), To, atol=1e-6)
        v.begin()
        v.update_vesselness(hs[2])
        np.testing.assert_allclose(v.get_response(), st2.response, rtol=1e-11, atol=1e-300)
        np.testing.assert_allclose(v.get_tensor(), To2, atol=1e-6)


def test_whole_front_end_on_the_reference_volume(M):
    """Five scales on ved_test.mhd entirely on the device; compared with the oracle and with the reference code's golden tensor."""
    img, sp = load_ved_test()
    with _ved(M, img.shape, sp) as v:
        v.set_image(img)
        for s in V.DEFAULT_SCALES:
            v.add_scale(s)
        T, resp = v.get_tensor(), v.get_response()
        st = v.stats()
    assert st["scales"] == 5 and st["kernel_launches"] >= 5 * 11
    To, so = V.ved_tensor(img.astype(np.float64), sp, **VED_TEST)
    bad = np.abs(T - To).max(axis=-1) > 1e-3
    assert bad.mean() < 1e-3, bad.mean()
    assert rel_l2(T, To) < 1e-4
    np.testing.assert_allclose(resp[~bad], so.response[~bad], rtol=1e-3, atol=1e-9)
    g = np.load(os.path.join(GOLDEN, "ref_vedfilter_gs_v.npz"))
    sl = (slice(None, None, int(g["sub"])),) * 3
    assert rel_l2(T[sl], g["tensor_sample"]) < 1e-4


def test_tensor_stays_on_the_device_for_the_solver(M):
    """madved_tensor_planes -> madgpu_set_tensor_device_f32: no host round trip of the tensor."""
    img, sp = _sub_volume()
    with _ved(M, img.shape, sp) as v, M.MadSolver(img.shape, sp, time_step=0.1) as s:
        v.set_image(img)
        for sc in V.DEFAULT_SCALES:
            v.add_scale(sc)
        s.set_tensor_device(v.tensor_planes())
        got = s.op_get_tensor(0)  # (6, nz, ny, nx)
        T = v.get_tensor()
    np.testing.assert_array_equal(np.moveaxis(got, 0, -1), T.astype(np.float32))


def test_whole_ved_filter_matches_reference_test(M):
    """test/itkVEDTest_GS.cxx ("v"): short pixels, GS, nu 3, 4 diffusion steps to 1e-10, through the filter class."""
    img, sp = load_ved_test()
    g = np.load(os.path.join(GOLDEN, "ref_vedfilter_gs_v.npz"))

    def run(out_double):
        f = M.VEDMultigridImageFilter("gs")
        f.SetCycle(f.VCYCLE)
        f.SetDiffusionIterationsPerGrid(3)
        f.SetInput(img.astype(np.float64) if out_double else img, sp)
        f.SetScales([0.300, 0.482, 0.775, 1.245, 2.000])
        f.SetAlpha(0.5); f.SetBeta(0.5); f.SetGamma(5.0); f.SetEpsilon(0.01); f.SetSensitivity(10.0)
        f.SetIterations(1)
        f.SetTolerance(1e-10)
        f.SetTimeStep(0.1)
        f.SetDiffusionIterations(4)
        f.SetOmega(1.5)
        f.Update()
        return f.GetOutput(), f.stats, f.ved_stats

    out, st, vst = run(True)
    assert out.dtype == np.float64 and st["steps"] == 4 and max(st["final_relres"]) <= 1e-10 and max(st["cycles_per_step"]) < 100
    assert vst["scales"] == 5
    sl = (slice(None, None, int(g["sub"])),) * 3
    assert rel_l2(out[sl], g["sample"]) < 1e-4  # Gauss-Seidel: converged image (BASELINE.json north_star)
    assert abs(np.linalg.norm(out) / g["stats"][0] - 1) < 1e-5
    short, _, _ = run(False)
    assert short.dtype == np.int16
    d = np.abs(short.astype(int) - g["out_short"].astype(int))
    assert d.max() <= 1 and (d != 0).mean() < 1e-3  # truncation of values that agree to ~1e-6


@pytest.mark.parametrize("smoother,cycle", [("wj", 0), ("gs", 1)])
def test_two_outer_iterations_match_oracle(M, smoother, cycle):
    """Iterations = 2: the vesselness state is rebuilt from the diffused image of the first pass (hxx:105-129)."""
    img, sp = _sub_volume()
    kw = dict(iterations=2, diffusion_iterations=2, smoother=0 if smoother == "gs" else 1, cycle=cycle, time_step=0.1, tolerance=1e-9,
              iterations_per_grid=2, **VED_TEST)
    want, _ = V.ved_filter(img, sp, V.DEFAULT_SCALES, **kw)
    with _ved(M, img.shape, sp) as v, M.MadSolver(img.shape, sp, time_step=0.1, smoother=kw["smoother"], iterations_per_grid=2, cycle=cycle,
                                                  tolerance=1e-9, max_cycles=100, number_of_steps=2) as s:
        out = v.run(s, img, V.DEFAULT_SCALES, iterations=2, out_dtype=np.float64)
        assert v.stats()["scales"] == 10
    assert rel_l2(out, want) < (1e-5 if smoother == "wj" else 1e-4)


def test_call_sequence_errors(M):
    shape, sp = (8, 8, 8), (1, 1, 1)
    with M.MadVed(shape, sp) as v:
        with pytest.raises(M.MadGpuError):
            v.hessian(1.0)  # no image
        with pytest.raises(M.MadGpuError):
            v.update_vesselness()  # no Hessian
        with pytest.raises(M.MadGpuError):
            v.tensor_planes()  # nothing consumed yet
        v.set_image(random_image(shape))
        with pytest.raises(M.MadGpuError):
            v.hessian(0.0)
        with M.MadSolver((8, 8, 16), sp) as s, pytest.raises(M.MadGpuError):
            v.run(s, random_image(shape))  # solver of another size
    with pytest.raises(M.MadGpuError):
        M.MadVed((8, 8, 3), sp)  # a line shorter than four samples
