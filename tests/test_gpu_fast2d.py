"""The 2-D streaming kernels (csrc/mad_fast2d.cuh: k2_sweep, k2_restrict, k2_prolong) -- the path of the reference's
itk2DDiffusionTest_{WJ,GS} images -- operator by operator against the oracle, on ragged shapes: several 128-column strips, partial
strips, nx % 4 in {0,1,2,3}, vertex- / cell-centred / mixed transfers, several y chunks.  MADGPU_FAST_MIN_NX=8 puts the coarser
levels of these small images on the streaming kernels too."""
import numpy as np
import pytest

from util import random_image, random_spd_tensor, rel_l2

pytestmark = pytest.mark.gpu

CASES = [
    ((40, 200), (1.0, 1.0), 0.1),      # cell-centred along both axes, two strips, the second one partial
    ((33, 131), (1.0, 0.5), 0.1),      # vertex-centred, nx % 4 == 3
    ((50, 97), (0.7, 1.3), 0.3),       # mixed centring, nx % 4 == 1
    ((64, 128), (0.3125, 0.3125), 0.1),  # exactly one full strip
    ((37, 66), (1.0, 1.0), 0.05),      # nx % 4 == 2
    ((19, 300), (0.5, 0.25), 0.1),     # three strips, few rows
]


@pytest.fixture(autouse=True)
def _streaming_everywhere(monkeypatch):
    monkeypatch.setenv("MADGPU_FAST_MIN_NX", "8")
    monkeypatch.setenv("MADGPU_FAST2D_MIN_PIXELS", "0")  # default 2^20: small levels keep the one-pixel-per-thread kernels except for Gauss-Seidel


def _mk(case, smoother=0, nu=2, seed=0):
    from multigridanisotropicdiffusion_b200 import MadSolver
    from oracle import oracle as O
    shape, sp, dt = case
    T = random_spd_tensor(shape, seed=seed)
    s = MadSolver(shape, sp, time_step=dt, smoother=smoother, iterations_per_grid=nu)
    s.set_tensor(T)
    o = O.Oracle(shape, sp, T.astype(np.float64), dt, smoother=smoother, nu=nu)
    return s, o


def _streaming_levels(s):
    """levels the 2-D streaming kernels run on under MADGPU_FAST_MIN_NX=8 (nx >= 8, ny >= 4)"""
    return [l for l in range(s.nlevels) if s.levels[l]["shape"][1] >= 8 and s.levels[l]["shape"][0] >= 4]


@pytest.mark.parametrize("case", CASES)
def test_the_streaming_kernels_are_the_ones_that_run(case):
    s, _ = _mk(case, smoother=0)
    t = s.gs_tile(0)
    assert t is not None and t[0] == 128 and t[2] == 1, t
    s.close()


@pytest.mark.parametrize("case", CASES)
def test_weighted_jacobi_sweep(case):
    s, o = _mk(case, smoother=1)
    for l in _streaming_levels(s):
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
    for l in _streaming_levels(s):
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
    # the residual image comes back as fp32 (it is the fp32 defect of the inner cycle); the norm is accumulated in fp64
    assert np.abs(g - r).max() < 1e-6 * np.abs(u).max() * np.abs(o.stencil(0)).sum(-1).max()
    assert abs(nrm - np.linalg.norm(r)) < 1e-12 * np.linalg.norm(r)
    s.close()


@pytest.mark.parametrize("case", CASES)
def test_restriction_and_prolongation(case):
    from oracle import oracle as O
    s, o = _mk(case)
    for l in _streaming_levels(s):
        if l + 1 >= s.nlevels:
            continue
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
@pytest.mark.parametrize("nu", [1, 3])
def test_vcycle_weighted_jacobi(case, nu):
    """north_star: weighted Jacobi within 1e-5 relative L2 per V-cycle."""
    s, o = _mk(case, smoother=1, nu=nu)
    shp = s.levels[0]["shape"]
    f = random_image(shp, seed=31)
    g = s.op_vcycle(0, f, f)
    r = o.vcycle(f.astype(np.float64), f.astype(np.float64), level=0)
    assert rel_l2(g, r) < 1e-5, rel_l2(g, r)
    s.close()


def gs_strip_sweep(S, u, f, tile):
    """CPU model (numpy, explicit 9-point rows `S` of the oracle) of k2_sweep<M2_GS>: tiles of (tx, ty) pixels; inside a tile rows in
    y order, inside a row the even columns, then the odd columns; pixels outside the tile keep the previous sweep's values."""
    TX, TY = tile[0], tile[1]
    ny, nx = u.shape
    old = np.zeros((ny + 2, nx + 2))
    old[1:-1, 1:-1] = u
    new = old.copy()
    tx = np.arange(-1, nx + 1) // TX
    ty = np.arange(-1, ny + 1) // TY
    offs = [(oy, ox) for oy in (-1, 0, 1) for ox in (-1, 0, 1)]
    xs_all = np.arange(nx)
    for y in range(ny):
        for par in (0, 1):
            xs = xs_all[xs_all % 2 == par]
            acc = f[y, xs].astype(np.float64).copy()
            for k, (oy, ox) in enumerate(offs):
                if (oy, ox) == (0, 0):
                    continue
                coef = S[y, xs, k]
                if not coef.any():
                    continue
                same = (ty[y + oy + 1] == ty[y + 1]) & (tx[xs + ox + 1] == tx[xs + 1])
                v = np.where(same, new[y + oy + 1, xs + ox + 1], old[y + oy + 1, xs + ox + 1])
                acc -= coef * v
            new[y + 1, xs + 1] = acc / S[y, xs, 4]
    return new[1:-1, 1:-1].copy()


@pytest.mark.parametrize("case", CASES)
def test_gs_sweep_is_the_documented_ordering(case):
    """One 2-D sweep == sequential Gauss-Seidel in the documented order, tile-local, evaluated on the CPU with the oracle's rows."""
    s, o = _mk(case, smoother=0)
    tile = s.gs_tile(0)
    shape = case[0]
    u, f = random_image(shape, seed=1), random_image(shape, seed=2)
    S = o.stencil(0)
    g1 = s.op_smooth(0, u, f, smoother=0, n_iter=1)
    r1 = gs_strip_sweep(S, u.astype(np.float64), f.astype(np.float64), tile)
    assert rel_l2(g1, r1) < 2e-6, rel_l2(g1, r1)
    assert np.abs(g1 - r1).max() < 3e-5 * np.abs(r1).max()
    g2 = s.op_smooth(0, u, f, smoother=0, n_iter=2)
    r2 = gs_strip_sweep(S, r1, f.astype(np.float64), tile)
    assert rel_l2(g2, r2) < 4e-6, rel_l2(g2, r2)
    s.close()


@pytest.mark.parametrize("case", CASES[:4])
@pytest.mark.parametrize("smoother", ["wj", "gs"])
def test_whole_solve_matches_oracle(case, smoother):
    from multigridanisotropicdiffusion_b200 import MadSolver
    from oracle import oracle as O
    shape, sp, dt = case
    sm = 1 if smoother == "wj" else 0
    T, img = random_spd_tensor(shape, seed=2), random_image(shape, seed=5)
    with MadSolver(shape, sp, time_step=dt, smoother=sm, iterations_per_grid=2, tolerance=1e-9, max_cycles=60, number_of_steps=2) as s:
        s.set_tensor(T)
        out = s.solve(img, out_dtype=np.float64)
        st = s.last_stats
    o = O.Oracle(shape, sp, T.astype(np.float64), dt, smoother=sm, nu=2)
    ref, cyc, _ = o.solve(img.astype(np.float64), tolerance=1e-9, max_cycles=60, number_of_steps=2)
    assert max(st["final_relres"]) <= 1e-9
    assert all(abs(a - b) <= (0 if smoother == "wj" else 2) for a, b in zip(st["cycles_per_step"], cyc)), (st["cycles_per_step"], cyc)
    assert rel_l2(out, ref) < (1e-5 if smoother == "wj" else 1e-4)
    assert rel_l2(out, ref) < 1e-7


def test_default_thresholds_on_a_megapixel_image(monkeypatch):
    """Default settings (streaming kernels from nx >= 64 and 2^20 pixels; Gauss-Seidel strips at every size): level 0 of a
    1030 x 1100 image runs k2_sweep / k2_restrict / k2_prolong, the coarser levels the one-pixel-per-thread kernels."""
    from multigridanisotropicdiffusion_b200 import MadSolver
    from oracle import oracle as O
    monkeypatch.delenv("MADGPU_FAST_MIN_NX")
    monkeypatch.delenv("MADGPU_FAST2D_MIN_PIXELS")
    shape, sp = (1030, 1100), (1.0, 1.0)
    T, img = random_spd_tensor(shape, seed=2), random_image(shape, seed=5)
    for smoother, sm in (("wj", 1), ("gs", 0)):
        with MadSolver(shape, sp, time_step=0.1, smoother=sm, iterations_per_grid=2, tolerance=1e-9, max_cycles=60) as s:
            s.set_tensor(T)
            out = s.solve(img, out_dtype=np.float64)
            st = s.last_stats
            if smoother == "gs":
                assert s.gs_tile(0) is not None and s.gs_tile(0)[2] == 1
        o = O.Oracle(shape, sp, T.astype(np.float64), 0.1, smoother=sm, nu=2)
        ref, cyc, _ = o.solve(img.astype(np.float64), tolerance=1e-9, max_cycles=60)
        assert st["final_relres"][0] <= 1e-9
        assert abs(st["cycles_per_step"][0] - cyc[0]) <= (0 if smoother == "wj" else 2), (st["cycles_per_step"], cyc)
        assert rel_l2(out, ref) < 1e-7, rel_l2(out, ref)
