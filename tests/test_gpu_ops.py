"""Per-operator parity of the CUDA path (through the C-ABI) against the CPU oracle.

Tolerances: the GPU works in fp32, the oracle in fp64; both get bit-identical fp32-rounded inputs.
Transfers use dyadic weights, so they agree to fp32 rounding; smoothers/residuals to a few ulp of
the operands' magnitude.
"""
import numpy as np
import pytest

from util import random_image, random_spd_tensor, rel_l2

pytestmark = pytest.mark.gpu

# (shape zyx, spacing xyz, dt): odd/even/mixed centring, anisotropic spacing
CASES_3D = [
    ((24, 26, 28), (1.0, 1.0, 1.0), 0.1),
    ((23, 25, 27), (0.3125, 0.3125, 0.5), 0.1),
    ((30, 35, 34), (0.330017, 0.330017, 0.330017), 0.05),   # mixed centring chain like ved_test_2
    ((13, 40, 21), (0.5, 0.25, 1.0), 0.02),
]
CASES_2D = [
    ((48, 64), (1.0, 1.0), 0.1),
    ((49, 33), (1.0, 0.5), 0.1),
    ((50, 97), (0.7, 1.3), 0.3),
]
CASES = CASES_3D + CASES_2D


def _mk(case, smoother=0, nu=2, seed=0):
    from multigridanisotropicdiffusion_b200 import MadSolver
    from oracle import oracle as O
    shape, sp, dt = case
    T = random_spd_tensor(shape, seed=seed)
    s = MadSolver(shape, sp, time_step=dt, smoother=smoother, iterations_per_grid=nu)
    s.set_tensor(T)
    o = O.Oracle(shape, sp, T.astype(np.float64), dt, smoother=smoother, nu=nu)
    return s, o


@pytest.mark.parametrize("case", CASES)
def test_level_schedule_matches_oracle(case):
    s, o = _mk(case)
    assert s.nlevels == o.nlevels
    for a, b in zip(s.levels, o.levels):
        assert a["n"] == b["n"]
        assert a["centering"] == b["centering"]
        np.testing.assert_allclose(a["h"], b["h"], rtol=1e-15)
    s.close()


@pytest.mark.parametrize("case", CASES)
def test_tensor_restriction(case):
    s, o = _mk(case)
    for l in range(s.nlevels):
        g = s.op_get_tensor(l)
        r = o.tensor(l)
        assert rel_l2(g, r) < 2e-7 * (l + 1), (l, rel_l2(g, r))
    s.close()


@pytest.mark.parametrize("case", CASES)
def test_operator_rows(case):
    """GenerateDCA: explicit 3^dim rows incl. Neumann folding and one-sided tensor derivatives."""
    s, o = _mk(case)
    for l in range(s.nlevels):
        g = s.op_assemble(l).astype(np.float64)
        r = o.stencil(l)
        scale = np.abs(r).max()
        assert np.abs(g - r).max() < 2e-6 * scale * (l + 1), (l, np.abs(g - r).max(), scale)
        # rows sum to one
        np.testing.assert_allclose(g.sum(-1), 1.0, atol=5e-5 * scale)
    s.close()


@pytest.mark.parametrize("case", CASES)
def test_weighted_jacobi_sweep(case):
    s, o = _mk(case, smoother=1)
    for l in range(s.nlevels):
        shp = s.levels[l]["shape"]
        u, f = random_image(shp, seed=l), random_image(shp, seed=l + 50)
        g = s.op_smooth(l, u, f, smoother=1, n_iter=1)
        r = o.smooth(l, u.astype(np.float64), f.astype(np.float64))
        assert rel_l2(g, r) < 2e-6, (l, rel_l2(g, r))
        g3 = s.op_smooth(l, u, f, smoother=1, n_iter=3)
        r3 = r
        for _ in range(2):
            r3 = o.smooth(l, r3, f.astype(np.float64))
        assert rel_l2(g3, r3) < 4e-6, (l, rel_l2(g3, r3))
    s.close()


@pytest.mark.parametrize("case", CASES)
def test_residual_and_norm(case):
    s, o = _mk(case)
    for l in range(s.nlevels):
        shp = s.levels[l]["shape"]
        u, f = random_image(shp, seed=l + 7), random_image(shp, seed=l + 57)
        g, nrm = s.op_residual(l, u, f)
        r = o.residual(l, u.astype(np.float64), f.astype(np.float64))
        scale = np.abs(u).max() * np.abs(o.stencil(l)).sum(-1).max()
        assert np.abs(g - r).max() < 4e-6 * scale, (l, np.abs(g - r).max(), scale)
        assert abs(nrm - np.linalg.norm(g.astype(np.float64))) < 1e-6 * nrm
    s.close()


@pytest.mark.parametrize("case", CASES)
def test_residual_f64(case):
    """Level-0 stop-test residual: fp64 arithmetic on fp32-stored tensor planes."""
    s, o = _mk(case)
    shp = s.levels[0]["shape"]
    u, f = random_image(shp, seed=3).astype(np.float64), random_image(shp, seed=4).astype(np.float64)
    g, nrm = s.op_residual_f64(u, f)
    r = o.residual(0, u, f)
    assert np.abs(g - r).max() < 1e-11 * np.abs(u).max() * np.abs(o.stencil(0)).sum(-1).max()
    assert abs(nrm - np.linalg.norm(r)) < 1e-12 * np.linalg.norm(r)
    s.close()


@pytest.mark.parametrize("case", CASES)
def test_restriction_and_prolongation(case):
    from oracle import oracle as O
    s, o = _mk(case)
    for l in range(s.nlevels - 1):
        cent = s.levels[l + 1]["centering"]
        fine = random_image(s.levels[l]["shape"], seed=l + 11)
        g = s.op_restrict(l, fine)
        r = O.restrict(fine.astype(np.float64), cent)
        assert g.shape == r.shape
        assert rel_l2(g, r) < 3e-7, (l, rel_l2(g, r))
        coarse = random_image(s.levels[l + 1]["shape"], seed=l + 21)
        gp = s.op_prolong(l, coarse)
        rp = O.interpolate(coarse.astype(np.float64), cent)
        assert gp.shape == rp.shape
        assert rel_l2(gp, rp) < 3e-7, (l, rel_l2(gp, rp))
    s.close()


@pytest.mark.parametrize("case", CASES)
def test_coarse_solve(case):
    s, o = _mk(case)
    shp = s.levels[-1]["shape"]
    f = random_image(shp, seed=99)
    g = s.op_coarse_solve(f)
    r = o.direct_solve(f.astype(np.float64))
    assert rel_l2(g, r) < 2e-6, rel_l2(g, r)
    s.close()


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("nu", [1, 2, 3])
def test_vcycle_weighted_jacobi(case, nu):
    """north_star: weighted Jacobi within 1e-5 relative L2 per V-cycle."""
    s, o = _mk(case, smoother=1, nu=nu)
    for l in range(s.nlevels):
        shp = s.levels[l]["shape"]
        f = random_image(shp, seed=l + 31)
        g = s.op_vcycle(l, f, f)
        r = o.vcycle(f.astype(np.float64), f.astype(np.float64), level=l)
        assert rel_l2(g, r) < 1e-5, (l, rel_l2(g, r))
    s.close()


@pytest.mark.parametrize("case", [CASES_3D[1], CASES_2D[0]])
@pytest.mark.parametrize("ncolors", [4, 8])
def test_multicolour_gs_is_a_gauss_seidel_ordering(case, ncolors, monkeypatch):
    """One multicolour sweep == sequential Gauss-Seidel in colour-major order (checked on the CPU with the
    oracle's explicit operator rows): the colouring has no intra-colour coupling."""
    monkeypatch.setenv("MADGPU_FAST2D", "0")  # the 64-pixel-wide 2-D case would otherwise get the strip sweep of mad_fast2d.cuh (tests/test_gpu_fast2d.py)
    from multigridanisotropicdiffusion_b200 import MadSolver
    from oracle import oracle as O
    shape, sp, dt = case
    dim = len(shape)
    if dim == 2 and ncolors == 8:
        pytest.skip("2-D always uses 4 colours")
    T = random_spd_tensor(shape, seed=5)
    s = MadSolver(shape, sp, time_step=dt, smoother=0, gs_colors=ncolors)
    s.set_tensor(T)
    o = O.Oracle(shape, sp, T.astype(np.float64), dt)
    u, f = random_image(shape, seed=1), random_image(shape, seed=2)
    g = s.op_smooth(0, u, f, smoother=0, n_iter=1)
    # CPU colour-major GS with the oracle's rows
    S = o.stencil(0)
    un = u.astype(np.float64).copy()
    idx = np.indices(shape)
    if dim == 2:
        col = (idx[1] & 1) | ((idx[0] & 1) << 1)
        offs = [(oy, ox) for oy in (-1, 0, 1) for ox in (-1, 0, 1)]
    else:
        p = (idx[2] & 1) | ((idx[1] & 1) << 1) | ((idx[0] & 1) << 2)
        col = p if ncolors == 8 else np.minimum(p, 7 - p)
        offs = [(oz, oy, ox) for oz in (-1, 0, 1) for oy in (-1, 0, 1) for ox in (-1, 0, 1)]
    centre = len(offs) // 2
    for c in range(4 if dim == 2 else ncolors):
        acc = f.astype(np.float64).copy()
        for k, off in enumerate(offs):
            if k == centre:
                continue
            coef = S[..., k]
            if not np.any(coef):
                continue
            sh = np.zeros_like(un)
            src = [slice(max(o_, 0), un.shape[a] + min(o_, 0)) for a, o_ in enumerate(off)]
            dst = [slice(max(-o_, 0), un.shape[a] + min(-o_, 0)) for a, o_ in enumerate(off)]
            sh[tuple(dst)] = un[tuple(src)]
            acc -= coef * sh
        new = acc / S[..., centre]
        un = np.where(col == c, new, un)
    assert rel_l2(g, un) < 2e-6, rel_l2(g, un)
    s.close()
