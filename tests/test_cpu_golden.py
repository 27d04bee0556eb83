"""The oracle against golden vectors recorded from the reference's own code (tests/golden/make_golden.py, which
runs the unmodified reference headers through oracle/_ref).  CPU only; needs nothing but the committed fixtures,
so the oracle stays pinned on machines where oracle/_ref is not available."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from util import GOLDEN, load_lena, load_ved_test, random_image, random_spd_tensor

SM = {"gs": 0, "wj": 1}
CY = {"v": 0, "fmg": 1, "s": 2}


def _golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def _stats(a):
    return np.array([np.linalg.norm(a), a.mean(), a.min(), a.max(), np.abs(np.diff(a, axis=-1)).sum()])


def _check(out, cyc, hist, g, float_pixels=False):
    assert list(cyc) == list(g["cycles"])
    sub = int(g["sub"])
    sl = tuple(slice(None, None, sub) for _ in out.shape)
    if float_pixels:  # the reference wrote float pixels: compare after the same cast
        out = out.astype(np.float32).astype(np.float64)
    np.testing.assert_allclose(out[sl], g["sample"], rtol=0, atol=1e-9 * np.abs(g["sample"]).max())
    np.testing.assert_allclose(_stats(out), g["stats"], rtol=1e-10)
    for step in range(len(cyc)):
        np.testing.assert_allclose(hist[step][:cyc[step]], g["relres"][step][:cyc[step]], rtol=1e-6, atol=1e-15)


@pytest.mark.parametrize("cycle", ["v", "fmg", "s"])
@pytest.mark.parametrize("smoother", ["wj", "gs"])
def test_reference_2d_tests(smoother, cycle):
    """test/itk2DDiffusionTest_{WJ,GS}.cxx with argv[1] in {v, fmg, s}."""
    g = _golden(f"ref_lena_{smoother}_{cycle}")
    img = load_lena().astype(np.float64)
    T = np.zeros(img.shape + (3,))
    T[..., 0] = 50.0
    T[..., 2] = 30.0
    o = O.Oracle(img.shape, (1.0, 1.0), T, 0.1, smoother=SM[smoother], nu=2)
    out, cyc, hist = o.solve(img, cycle=CY[cycle], tolerance=1e-10, max_cycles=100)
    _check(out, cyc, hist, g, float_pixels=True)


@pytest.mark.parametrize("cycle", ["v", "fmg"])
@pytest.mark.parametrize("smoother", ["wj", "gs"])
@pytest.mark.parametrize("tag,shape,sp", [("small2d", (49, 33), (0.7, 1.3)), ("small3d", (23, 25, 27), (0.3125, 0.3125, 0.5))])
def test_cross_terms_mixed_centring(tag, shape, sp, smoother, cycle):
    g = _golden(f"ref_{tag}_{smoother}_{cycle}")
    T = random_spd_tensor(shape, seed=2).astype(np.float64)
    img = random_image(shape, seed=5).astype(np.float64)
    o = O.Oracle(shape, sp, T, 0.1, smoother=SM[smoother], nu=2)
    out, cyc, hist = o.solve(img, cycle=CY[cycle], tolerance=1e-10, max_cycles=100, number_of_steps=2)
    _check(out, cyc, hist, g)


def test_reference_ved_diffusion_step():
    """DiffusionStep of test/itkVEDTest_GS.cxx on the reference's own volume."""
    from multigridanisotropicdiffusion_b200 import phantom
    g = _golden("ref_ved_gs_v")
    vol, sp = load_ved_test()
    _, D = phantom.vessel_phantom(vol.shape, spacing=sp)
    T = phantom.planes_to_aos(D).numpy().astype(np.float64)
    o = O.Oracle(vol.shape, sp, T, 0.1, smoother=0, nu=3)
    out, cyc, hist = o.solve(vol.astype(np.float64), tolerance=1e-10, max_cycles=100, number_of_steps=4)
    _check(out, cyc, hist, g)


def test_reference_ved_whole_filter():
    """test/itkVEDTest_GS.cxx ("v") end to end -- Hessians, vesselness, tensor, DiffusionStep, cast -- against the vectors the
    reference's own VED filter produced (tests/golden/make_golden_ved.py)."""
    from oracle import ved as V
    g = _golden("ref_vedfilter_gs_v")
    vol, sp = load_ved_test()
    out, info = V.ved_filter(vol, sp, V.DEFAULT_SCALES, alpha=0.5, beta=0.5, gamma=5.0, epsilon=0.01, omega=1.5, sensitivity=10.0,
                             iterations=1, diffusion_iterations=4, smoother=0, cycle=0, time_step=0.1, tolerance=1e-10,
                             iterations_per_grid=3)
    sub = int(g["sub"])
    sl = (slice(None, None, sub),) * 3
    T = info["tensors"][-1]
    np.testing.assert_allclose(T[sl], g["tensor_sample"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(np.stack([_stats(T[..., k]) for k in range(6)]), g["tensor_stats"], rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(out[sl], g["sample"], rtol=0, atol=1e-9 * np.abs(g["sample"]).max())
    np.testing.assert_allclose(_stats(out), g["stats"], rtol=1e-10)
    short = np.trunc(out).astype(np.int16)  # static_cast< short >
    assert (short != g["out_short"]).mean() < 1e-5 and np.abs(short.astype(int) - g["out_short"].astype(int)).max() <= 1
