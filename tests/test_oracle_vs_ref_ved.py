"""Pins the VED part of the oracle (oracle/ved_oracle.c + oracle/ved.py: vesselness function, magnitude sort, arg-max over
scales, eigenvector bookkeeping, Q D Q^T, the GenerateData loop with DiffusionStep and the output cast) against the
reference's OWN code: /root/reference/include/itkVEDMultigridImageFilter.{h,hxx} compiled unmodified into
oracle/_ref/libmadref.so.  The Hessian filter and the eigen-solver underneath are third-party (ITK / VXL, absent) and are
the same stand-in on both sides, so they are NOT pinned by this file (see the header of oracle/ved_oracle.c).

CPU only; skipped where libmadref.so is absent (tests/test_cpu_golden.py holds vectors recorded from it).
"""
import numpy as np
import pytest

from oracle import ref as R
from oracle import ved as V
from util import load_ved_test, random_image

pytestmark = pytest.mark.skipif(not (R.available() and R.ved_available()), reason="oracle/_ref/libmadref.so (with the VED filter) not built")

VED_TEST = dict(alpha=0.5, beta=0.5, gamma=5.0, epsilon=0.01, omega=1.5, sensitivity=10.0)  # test/itkVEDTest_GS.cxx:82-99


def test_vesselness_function():
    rng = np.random.default_rng(0)
    for _ in range(500):
        e = rng.normal(size=3) * rng.choice([1e-4, 0.1, 1.0, 30.0])
        e = e[np.argsort(np.abs(e))]
        if rng.random() < 0.6:
            e[1], e[2] = -abs(e[1]), -abs(e[2])  # the branch that is not identically zero
        for a, b, g in ((0.5, 0.5, 5.0), (0.3, 0.7, 25.0)):
            assert V.vesselness(e, a, b, g) == R.ved_vesselness(e, a, b, g)
    assert R.ved_vesselness([0.0, 0.0, 0.0]) == 0.0
    assert R.ved_vesselness([0.1, -1.0, 2.0]) == 0.0


def _sub_volume():
    img, sp = load_ved_test()
    return img[20:44, 24:52, 18:48].astype(np.float64), sp


def test_update_vesselness_and_tensor_on_real_hessians():
    img, sp = _sub_volume()
    hs = [V.hessian(img, sp, s) for s in V.DEFAULT_SCALES]
    ref = R.ved_tensor_from_hessians(hs, sp, **VED_TEST)
    T, st = V.ved_tensor(img, sp, hessians=hs, **VED_TEST)
    np.testing.assert_array_equal(st.response, ref["response"])
    np.testing.assert_array_equal(st.eigenvalues, ref["eigenvalues"])
    np.testing.assert_array_equal(st.eigenvectors, ref["eigenvectors"])
    np.testing.assert_allclose(T, ref["tensor"], rtol=0, atol=1e-14)
    assert (st.response > 0).mean() > 0.05  # the arg-max branch is exercised, not only the identity fallback
    # eigenvalues of the tensor lie in [epsilon, omega]
    M = np.zeros(T.shape[:-1] + (3, 3))
    for k, (a, b) in enumerate(((0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2))):
        M[..., a, b] = M[..., b, a] = T[..., k]
    w = np.linalg.eigvalsh(M)
    assert w.min() >= VED_TEST["epsilon"] - 1e-12 and w.max() <= VED_TEST["omega"] + 1e-12


def test_update_vesselness_random_hessians_and_first_scale_rule():
    """Random symmetric matrices (all sign patterns); the first Hessian is stored unconditionally (hxx:272)."""
    rng = np.random.default_rng(3)
    shape = (5, 6, 7)
    hs = [rng.normal(size=shape + (6,)) * s for s in (1.0, 3.0, 0.2)]
    ref = R.ved_tensor_from_hessians(hs, (1.0, 1.0, 1.0))
    T, st = V.ved_tensor(np.zeros(shape), (1.0, 1.0, 1.0), scales=(1, 2, 3), hessians=hs)
    np.testing.assert_array_equal(st.response, ref["response"])
    np.testing.assert_array_equal(st.eigenvectors, ref["eigenvectors"])
    np.testing.assert_allclose(T, ref["tensor"], rtol=0, atol=1e-14)


@pytest.mark.parametrize("pixel,smoother,cycle", [("double", 0, 0), ("short", 0, 0), ("double", 1, 1)])
def test_whole_ved_filter(pixel, smoother, cycle):
    """GenerateData (hxx:63-155): two outer iterations so that the vesselness state is reset between them (:121-123)."""
    img, sp = _sub_volume()
    if pixel == "short":
        img = np.round(img)
    kw = dict(iterations=2, diffusion_iterations=2, smoother=smoother, cycle=cycle, time_step=0.1, tolerance=1e-8, iterations_per_grid=2,
              **VED_TEST)
    out, T = R.run_ved_filter(img, sp, V.DEFAULT_SCALES, pixel=pixel, **kw)
    o_out, info = V.ved_filter(img, sp, V.DEFAULT_SCALES, out_dtype=np.int16 if pixel == "short" else None, **kw)
    np.testing.assert_allclose(info["tensors"][-1], T, rtol=0, atol=1e-9)
    if pixel == "short":
        assert np.abs(o_out.astype(np.float64) - out).max() <= 1  # truncation of values that agree to 1e-9
        assert (o_out.astype(np.float64) != out).mean() < 1e-3
    else:
        np.testing.assert_allclose(o_out, out, rtol=0, atol=1e-8)


def test_line_shorter_than_four_samples_is_refused():
    with pytest.raises(RuntimeError):
        V.hessian(random_image((3, 8, 8)).astype(np.float64), (1, 1, 1), 1.0)
