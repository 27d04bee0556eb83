"""Pins the oracle (oracle/mad_oracle.c, the C restatement every parity test compares the CUDA path with)
against the reference's OWN code: oracle/_ref/libmadref.so is the unmodified /root/reference/include headers
compiled against the stand-in ITK of oracle/shim.  Every routine of the hot path is compared; agreement is to
rounding (in practice bit for bit, the restatement follows the reference's operation order).

CPU only.  libmadref.so is built in the authoring container (`make -C oracle ref`, needs /root/reference) and
travels as a built artefact; where it is absent these tests skip and tests/test_cpu_golden.py (vectors recorded
from the same library) still pins the oracle.
"""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import ref as R
from util import random_image, random_spd_tensor

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libmadref.so not built (needs /root/reference)")

CASES = [
    ((20, 22), (0.7, 1.3), 0.1),
    ((33, 48), (1.0, 1.0), 0.3),
    ((12, 14, 13), (0.3125, 0.3125, 0.5), 0.1),
    ((13, 12, 15), (1.0, 0.5, 2.0), 0.05),
    ((14, 25, 12), (0.33, 0.33, 0.33), 0.1),
]


@pytest.mark.parametrize("case", CASES)
def test_hierarchy_operator_rows_smoothers_residual_direct_solver(case):
    shape, sp, dt = case
    T = random_spd_tensor(shape, seed=2).astype(np.float64)
    o = O.Oracle(shape, sp, T, dt)
    r = R.Reference(shape, sp, T, dt)
    assert o.nlevels == r.nlevels
    for lo, lr in zip(o.levels, r.levels):
        assert lo["n"] == lr["n"] and lo["centering"] == lr["centering"]
        np.testing.assert_allclose(lo["h"], lr["h"], rtol=1e-15)
    dim = len(shape)
    for l in range(o.nlevels):
        S, active = r.stencil(l)
        np.testing.assert_array_equal(S, o.stencil(l))  # GenerateDCA, incl. Neumann folding and one-sided differences
        assert (active > 0).sum() == (9 if dim == 2 else 19)
        # active offsets are kept in Neighborhood raster order (StencilImage.hxx:57-65)
        assert list(active[active > 0]) == sorted(active[active > 0])
        shp = o.levels[l]["shape"]
        u, f = random_image(shp, seed=l + 1).astype(np.float64), random_image(shp, seed=l + 9).astype(np.float64)
        for sm in (O.Oracle.GS, O.Oracle.WJ):
            o.set_smoother(sm)
            np.testing.assert_allclose(o.smooth(l, u, f), r.smooth(l, u, f, sm), rtol=0, atol=1e-12)
            np.testing.assert_allclose(o.residual(l, u, f), r.residual(l, u, f, sm), rtol=0, atol=1e-12)
    fl = random_image(o.levels[-1]["shape"], seed=3).astype(np.float64)
    np.testing.assert_allclose(o.direct_solve(fl), r.direct_solve(fl), rtol=1e-11, atol=1e-11)


@pytest.mark.parametrize("shape", [(16, 18), (17, 19), (16, 19), (7, 6), (12, 14, 16), (13, 15, 17), (12, 15, 16), (13, 14, 7)])
def test_transfers(shape):
    cent = tuple(1 if n % 2 == 0 else 0 for n in shape[::-1])
    a = random_image(shape, seed=4).astype(np.float64)
    rc, oc = R.restrict(a, cent), O.restrict(a, cent)
    assert rc.shape == oc.shape
    np.testing.assert_allclose(oc, rc, rtol=0, atol=1e-12)
    rp, op = R.interpolate(rc, cent), O.interpolate(oc, cent)
    assert rp.shape == op.shape == shape
    np.testing.assert_allclose(op, rp, rtol=0, atol=1e-12)


@pytest.mark.parametrize("shape,sp", [((49, 33), (0.7, 1.3)), ((18, 20, 17), (0.3125, 0.3125, 0.5))])
@pytest.mark.parametrize("smoother", [0, 1])
@pytest.mark.parametrize("cycle", [0, 1, 2])
def test_filter_generate_data(shape, sp, smoother, cycle):
    """The whole GenerateData() of the reference (time-step loop, V-cycle / FMG / smoother-only, stop test) against
    oracle.solve: same cycle counts, same per-cycle relative residuals, same image."""
    T = random_spd_tensor(shape, seed=6).astype(np.float64)
    img = random_image(shape, seed=7).astype(np.float64)
    kw = dict(nu=2, time_step=0.1, tolerance=1e-9, max_cycles=12, number_of_steps=2)
    out, cycles, log = R.run_filter(img, sp, T, smoother=smoother, cycle=cycle, pixel="double", **kw)
    o = O.Oracle(shape, sp, T, 0.1, smoother=smoother, nu=2)
    oo, oc, hist = o.solve(img, cycle=cycle, tolerance=1e-9, max_cycles=12, number_of_steps=2, faithful=True)
    assert cycles == oc
    np.testing.assert_allclose(oo, out, rtol=0, atol=1e-10)
    rr = R.relres_per_cycle(log, cycle == 2)
    for step in range(2):
        np.testing.assert_allclose(hist[step][:oc[step]], rr[step], rtol=1e-6, atol=1e-15)  # residuals near 1e-10 carry rounding noise
    # the lean V-cycle (no logging-only residual passes) is the same iteration
    ol, ocl, _ = o.solve(img, cycle=cycle, tolerance=1e-9, max_cycles=12, number_of_steps=2, faithful=False)
    assert ocl == oc
    np.testing.assert_allclose(ol, oo, rtol=0, atol=1e-10)


def test_output_pixel_cast_is_truncation():
    """static_cast<OutputPixelType>(double), …Filter.hxx:277: short output truncates toward zero."""
    shape, sp = (14, 16, 13), (1.0, 1.0, 1.0)
    # the tensor image has the INPUT pixel type (…Filter.h:111-112), so a short image comes with a short tensor:
    # use integer-valued entries to give both runs the same operator
    T = np.round(8.0 * random_spd_tensor(shape, seed=8).astype(np.float64))
    T[..., [0, 3, 5]] += 4.0
    img = np.round(random_image(shape, seed=9).astype(np.float64) - 100.0)  # negative values too
    kw = dict(smoother=0, cycle=0, nu=2, time_step=0.1, tolerance=1e-10, max_cycles=50, number_of_steps=1)
    d, _, _ = R.run_filter(img, sp, T, pixel="double", **kw)
    s, _, _ = R.run_filter(img, sp, T, pixel="short", **kw)
    np.testing.assert_array_equal(s, np.trunc(d))
