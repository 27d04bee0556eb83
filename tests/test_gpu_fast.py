"""Parity of the streaming 3-D kernels (csrc/mad_fast.cuh: 4 voxels per thread, z-marching in
registers) against the CPU oracle.  The library only switches to them for nx >= 64; the test hook
MADGPU_FAST_MIN_NX forces them on every level so that ragged sizes are covered: nx not a multiple
of 4 (last voxel in slot 0..3 of a thread), nx just above a multiple of 128 (second warp column
with one active lane), 3-voxel axes (mirror and one-sided tensor differences overlap).
"""
import numpy as np
import pytest

from util import random_image, random_spd_tensor, rel_l2

pytestmark = pytest.mark.gpu

# (shape zyx, spacing xyz, dt)
CASES = [
    ((12, 14, 64), (1.0, 1.0, 1.0), 0.1),
    ((9, 11, 129), (0.3125, 0.3125, 0.5), 0.1),
    ((10, 9, 130), (0.5, 0.25, 1.0), 0.05),
    ((7, 13, 131), (1.0, 0.7, 1.3), 0.1),
    ((11, 10, 261), (1.0, 1.0, 1.0), 0.2),
    ((23, 25, 27), (0.3125, 0.3125, 0.5), 0.1),
    ((40, 6, 33), (0.33, 0.33, 0.33), 0.1),
    ((6, 37, 30), (1.0, 2.0, 0.5), 0.1),
    ((3, 3, 9), (1.0, 1.0, 1.0), 0.1),
]


@pytest.fixture(params=[0, 1, 4, 5], ids=lambda c: f"cfg{c}")
def fast_env(request, monkeypatch):
    monkeypatch.setenv("MADGPU_FAST_MIN_NX", "0")
    monkeypatch.setenv("MADGPU_FAST_CFG", str(request.param))
    return request.param


# cases with a proper hierarchy and a small coarsest grid (the oracle factorises it densely)
CASES_MG = [
    ((12, 14, 64), (1.0, 1.0, 1.0), 0.1),
    ((24, 26, 140), (0.3125, 0.3125, 0.5), 0.1),
    ((25, 27, 133), (0.5, 0.25, 1.0), 0.05),
    ((23, 25, 27), (0.3125, 0.3125, 0.5), 0.1),
]


def _mk(case, smoother=1, nu=2, seed=0, max_coarse=2000):
    """max_coarse: the oracle skips its dense coarsest-grid factorisation above this size (operator-level
    tests do not need it; thin ragged volumes stop coarsening early and would take minutes)."""
    from multigridanisotropicdiffusion_b200 import MadSolver
    from oracle import oracle as O
    shape, sp, dt = case
    T = random_spd_tensor(shape, seed=seed)
    s = MadSolver(shape, sp, time_step=dt, smoother=smoother, iterations_per_grid=nu)
    s.set_tensor(T)
    o = O.Oracle(shape, sp, T.astype(np.float64), dt, smoother=smoother, nu=nu, max_coarse=max_coarse)
    return s, o


@pytest.mark.parametrize("case", CASES)
def test_fast_weighted_jacobi_sweep(case, fast_env):
    s, o = _mk(case)
    for l in range(s.nlevels):
        shp = s.levels[l]["shape"]
        u, f = random_image(shp, seed=l), random_image(shp, seed=l + 50)
        g = s.op_smooth(l, u, f, smoother=1, n_iter=1)
        r = o.smooth(l, u.astype(np.float64), f.astype(np.float64))
        assert rel_l2(g, r) < 2e-6, (l, rel_l2(g, r))
        assert np.abs(g - r).max() < 2e-5 * np.abs(r).max(), (l, np.abs(g - r).max())
    s.close()


@pytest.mark.parametrize("case", CASES)
def test_fast_residual_and_norm(case, fast_env):
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


@pytest.mark.parametrize("rows", ["fp32", "fp64"])
@pytest.mark.parametrize("case", CASES)
def test_fast_residual_f64(case, fast_env, rows, monkeypatch):
    """The level-0 stop-test / defect residual of the solve loop (streaming kernel, fp32 r + fp64 norm).  Default: the operator row
    is evaluated in fp32 -- bit for bit the row the fp32 sweeps relax -- and APPLIED in fp64 to the fp64 iterate, so the norm agrees
    with the all-fp64 oracle to the rounding of the row (~1e-7 relative; what the fp32 storage of the tensor costs as well).
    MADGPU_RES64_COEF32=0 evaluates the row in fp64 too: 1e-12."""
    monkeypatch.setenv("MADGPU_RES64_COEF32", "1" if rows == "fp32" else "0")
    s, o = _mk(case)
    shp = s.levels[0]["shape"]
    u, f = random_image(shp, seed=3).astype(np.float64), random_image(shp, seed=4).astype(np.float64)
    # norm_only runs the kernel of the solve loop
    _, nrm = s.op_residual_f64(u, f, norm_only=True)
    r = o.residual(0, u, f)
    assert abs(nrm - np.linalg.norm(r)) < (2e-6 if rows == "fp32" else 1e-12) * np.linalg.norm(r)
    s.close()


@pytest.mark.parametrize("case", CASES)
def test_fast_stop_test_residual(case, fast_env):
    """cycles_begin forms r = f - A f in fp64 arithmetic with the streaming kernel; after 0 cycles the
    relative residual reported by one SMOOTHER-mode iteration must match the oracle's."""
    from multigridanisotropicdiffusion_b200 import MadSolver
    s, o = _mk(case, smoother=1, nu=2)
    shape = case[0]
    img = random_image(shape, seed=9)
    s.set_solver(cycle=MadSolver.SMOOTHER)
    s.cycles_begin(img)
    rr, _, _ = s.cycles_run(2)
    u = img.astype(np.float64)
    f = u.copy()
    ref = []
    for _ in range(2):
        u = o.smooth(0, u, f)
        ref.append(np.linalg.norm(o.residual(0, u, f)) / np.linalg.norm(f))
    np.testing.assert_allclose(rr, ref, rtol=2e-5)
    out = s.cycles_end()
    assert rel_l2(out, u) < 1e-6
    s.close()


@pytest.mark.parametrize("case", CASES_MG)
def test_fast_vcycle_weighted_jacobi(case, fast_env):
    """north_star: weighted Jacobi within 1e-5 relative L2 per V-cycle."""
    s, o = _mk(case, smoother=1, nu=2)
    f = random_image(case[0], seed=31)
    g = s.op_vcycle(0, f, f)
    r = o.vcycle(f.astype(np.float64), f.astype(np.float64), level=0)
    assert rel_l2(g, r) < 1e-5, rel_l2(g, r)
    s.close()


# ---------------------------------------------------------------------------------- fused Gauss-Seidel
GS_CASES = [
    ((12, 14, 64), (1.0, 1.0, 1.0), 0.1),
    ((10, 9, 130), (0.5, 0.25, 1.0), 0.05),      # two tiles along x, odd row count
    ((21, 11, 140), (0.3125, 0.3125, 0.5), 0.1),  # several z chunks? (tile from gs_tile)
    ((23, 25, 27), (0.3125, 0.3125, 0.5), 0.1),
    ((9, 6, 131), (1.0, 0.7, 1.3), 0.1),
    ((14, 18, 200), (1.0, 1.0, 1.0), 0.1),        # even ny > 8: several row-pair tiles
    ((40, 36, 136), (0.3125, 0.3125, 0.5), 0.1),  # temporal blocking: several 16-row tiles, two z chunks on the shifted grid, two warp columns
    ((19, 16, 64), (1.0, 1.0, 1.0), 0.1),         # odd plane count, exactly one tile of rows
]


@pytest.fixture(params=[(0, 0, 0, 1), (1, 0, 0, 1), (7, 0, 0, 1), (8, 0, 0, 1), (0, 1, 0, 1), (0, 1, 1, 1), (0, 1, 1, 3), (0, 1, 1, 0), (0, 1, 1, -2), (0, 1, 1, -3), (0, 1, 2, 1), (0, 1, 3, 1)],
                ids=lambda c: f"cfg{c[0]}-rows{'16' if c[1] else '32'}{'-pairs' if c[2] == 1 else '-private' if c[2] == 2 else '-privalt' if c[2] == 3 else ''}{'-tb3' if c[3] == 3 else '-tbsingle' if c[3] == 0 else '-staged' + str(-c[3]) if c[3] < 0 else ''}")
def gs_env(request, monkeypatch):
    """(kernel variant, packed fp16 operator rows, one warp per row pair, sweeps fused per pass): the default is (0, 1, 1, 1); the
    exact-row variants pin the ordering to fp32 rounding; the last one is the opt-in temporal blocking (MADGPU_GS_TB=3)."""
    monkeypatch.setenv("MADGPU_GS_TB", str(max(request.param[3], 1)))
    # every sweep through k_coef_gs_tb<1>: 1 = u ring only, 2 / 3 = packed rows and f staged by cp.async as well (tiles of 8 / 16 rows)
    monkeypatch.setenv("MADGPU_GS_TB_SINGLE", "1" if request.param[3] == 0 else str(-request.param[3]) if request.param[3] < 0 else "0")
    monkeypatch.setenv("MADGPU_FAST_MIN_NX", "0")
    monkeypatch.setenv("MADGPU_FAST_CFG", str(request.param[0]))
    monkeypatch.setenv("MADGPU_GS_COEF16", str(request.param[1]))
    monkeypatch.setenv("MADGPU_GS_PAIRS", str(min(request.param[2], 1)))
    monkeypatch.setenv("MADGPU_GS_PRIVATE", str(max(request.param[2] - 1, 0)))  # warp-private tiles of the row-pair kernel (2: alternating grid)
    return request.param


@pytest.mark.parametrize("case", GS_CASES)
def test_fused_gs_sweep_is_the_documented_ordering(case, gs_env):
    """A Gauss-Seidel leg == the documented ordering, evaluated on the CPU with the oracle's explicit operator rows: passes as
    planned by the library (gs_leg_plan); inside a tile sequential Gauss-Seidel (z planes; even rows: even x, odd x; odd rows);
    a pass that fuses several sweeps (temporal blocking, packed rows only) keeps the values outside the tile frozen at those it
    started from and alternates the tile grid between passes."""
    from util import gs_leg_model
    s, o = _mk(case, smoother=0)
    tile = s.gs_tile(0)
    assert tile is not None and tile[0] == 128
    shape = case[0]
    u, f = random_image(shape, seed=1), random_image(shape, seed=2)
    S = o.stencil(0)
    # exact rows: fp32 rounding only; packed fp16 rows: the operator itself is rounded to 11 bits (the smoother inside
    # an exact defect-correction loop, see k_coef_gs)
    tol = 2e-6 if gs_env[1] == 0 else 2e-3
    fused_seen = 0
    for n_iter in (1, 2, 3, 5):
        plan = s.gs_leg_plan(0, n_iter)  # before the call: the plan depends on the alternation state the call advances
        assert sum(p["fused"] for p in plan) == n_iter
        fused_seen = max(fused_seen, max(p["fused"] for p in plan))
        g = s.op_smooth(0, u, f, smoother=0, n_iter=n_iter)
        r = gs_leg_model(S, u.astype(np.float64), f.astype(np.float64), plan)
        assert rel_l2(g, r) < n_iter * tol, (n_iter, plan, rel_l2(g, r))
        assert np.abs(g - r).max() < 15 * n_iter * tol * np.abs(r).max()
    if gs_env[3] == 3 and shape[1] % 2 == 0 and shape[1] >= 8 and shape[0] >= 8:
        assert fused_seen == 3  # MADGPU_GS_TB=3 really fuses
    if gs_env[3] <= 1:
        assert fused_seen == 1
    s.close()


@pytest.mark.parametrize("case", CASES_MG)
def test_fused_gs_converged_image(case, gs_env):
    """north_star: Gauss-Seidel within 1e-4 relative L2 on the converged diffused image (the reference sweeps
    lexicographically; only the fixed point is comparable)."""
    import multigridanisotropicdiffusion_b200 as M
    from oracle import oracle as O
    shape, sp, dt = case
    T = random_spd_tensor(shape, seed=3)
    img = random_image(shape, seed=4)
    f = M.MultigridAnisotropicDiffusionImageFilter("gs")
    f.SetInput(img, sp)
    f.SetDiffusionTensor(T)
    f.SetTimeStep(dt)
    f.SetIterationsPerGrid(2)
    f.SetTolerance(1e-9)
    f.SetNumberOfSteps(2)
    f.Update()
    out = f.GetOutput().astype(np.float64)
    st = f.stats
    f.close()
    o = O.Oracle(shape, sp, T.astype(np.float64), dt, smoother=0, nu=2)
    ref, cyc, _ = o.solve(img.astype(np.float64), tolerance=1e-9, number_of_steps=2)
    assert rel_l2(out, ref) < 1e-4, rel_l2(out, ref)
    assert max(st["final_relres"][:2]) <= 1e-9
    assert max(st["cycles_per_step"][:2]) <= max(cyc) + 2, (st["cycles_per_step"], cyc)


@pytest.mark.parametrize("case", CASES + [((20, 22, 260), (1.0, 1.0, 1.0), 0.1), ((17, 19, 257), (1.0, 1.0, 1.0), 0.1),
                                  # all axes cell-centred (k_fast_prolong_cell): odd coarse row length / a last thread with 2 of 4 voxels,
                                  # a row that ends exactly with a thread, two warp columns
                                  ((12, 14, 134), (1.0, 1.0, 1.0), 0.1), ((8, 12, 132), (1.0, 1.0, 1.0), 0.1), ((16, 24, 264), (1.0, 1.0, 1.0), 0.1)])
def test_fast_restriction_and_prolongation(case, fast_env):
    """Streaming transfer kernels (4 fine voxels per thread) against the oracle, every centring combination."""
    from oracle import oracle as O
    s, o = _mk(case)
    for l in range(s.nlevels - 1):
        cent = s.levels[l + 1]["centering"]
        fine = random_image(s.levels[l]["shape"], seed=l + 11)
        g = s.op_restrict(l, fine)
        r = O.restrict(fine.astype(np.float64), cent)
        assert g.shape == r.shape
        assert rel_l2(g, r) < 3e-7, (l, rel_l2(g, r))
        assert np.abs(g - r).max() < 1e-6 * np.abs(r).max()
        coarse = random_image(s.levels[l + 1]["shape"], seed=l + 21)
        gp = s.op_prolong(l, coarse)
        rp = O.interpolate(coarse.astype(np.float64), cent)
        assert gp.shape == rp.shape
        assert rel_l2(gp, rp) < 3e-7, (l, rel_l2(gp, rp))
        assert np.abs(gp - rp).max() < 1e-6 * np.abs(rp).max()
    s.close()
