"""CPU-only tests of the oracle (oracle/mad_oracle.c, the restatement of the reference's algorithm): the
closed-form properties SURVEY.md appendix A derives from the reference code, the level tables of
SURVEY.md section 8(a), and the behaviour of the drivers on the reference's own test configurations."""
import numpy as np
import pytest

from oracle import oracle as O
from util import load_lena, load_ved_test, random_image, random_spd_tensor, rel_l2


# ---------------------------------------------------------------------------- level schedule (GridsHierarchy.hxx:36-106)
@pytest.mark.parametrize("size,expect", [
    ((512, 512), [(512, 512), (256, 256), (128, 128), (64, 64), (32, 32), (16, 16), (8, 8)]),
    ((69, 77, 69), [(69, 77, 69), (35, 39, 35), (18, 20, 18), (9, 10, 9)]),
    ((134, 140, 119), [(134, 140, 119), (67, 70, 60), (34, 35, 30), (17, 18, 15), (9, 9, 8)]),
    ((256, 256, 256), [(256,) * 3, (128,) * 3, (64,) * 3, (32,) * 3, (16,) * 3, (8,) * 3]),
    ((10, 40, 40), [(10, 40, 40)]),
    ((11, 40, 40), [(11, 40, 40), (6, 20, 20)]),
])
def test_level_schedule_tables(size, expect):
    got = O.level_schedule(size)
    assert [g[0] for g in got] == expect
    for (nf, _), (nc, cent) in zip(got[:-1], got[1:]):
        for d in range(len(size)):
            assert cent[d] == (1 if nf[d] % 2 == 0 else 0)
            assert nc[d] == (nf[d] // 2 if nf[d] % 2 == 0 else (nf[d] - 1) // 2 + 1)


def test_mixed_centring_chain_of_ved_test_2():
    got = O.level_schedule((134, 140, 119))
    assert [g[1] for g in got[1:]] == [(1, 1, 0), (0, 1, 1), (1, 0, 1), (0, 1, 0)]


# ---------------------------------------------------------------------------- transfers (InterGridOperators .h:101-127)
@pytest.mark.parametrize("shape", [(16, 18), (17, 19), (16, 19), (12, 14, 16), (13, 15, 17), (12, 15, 16)])
def test_transfers_preserve_constants_and_are_linear(shape):
    cent = tuple(1 if n % 2 == 0 else 0 for n in shape[::-1])
    one = np.ones(shape)
    c = O.restrict(one, cent)
    np.testing.assert_allclose(c, 1.0, rtol=0, atol=1e-15)  # full-weighting rows sum to one
    f = O.interpolate(np.ones_like(c), cent)
    assert f.shape == shape
    np.testing.assert_allclose(f, 1.0, rtol=0, atol=1e-15)
    a, b = random_image(shape, 1).astype(np.float64), random_image(shape, 2).astype(np.float64)
    np.testing.assert_allclose(O.restrict(2 * a - 3 * b, cent), 2 * O.restrict(a, cent) - 3 * O.restrict(b, cent), atol=1e-10)


def test_transfer_stencils_1d_tables():
    """Vertex axis: [1/4 1/2 1/4] with injection at both ends; cell axis: [1/8 3/8 3/8 1/8], ends [1/2 3/8 1/8]
    (InterGridOperators.h:115-127); interpolation: vertex linear, cell 3/4-1/4 with end copies (.h:101-113)."""
    e = np.zeros((9, 9)); e[4, 4] = 1.0  # vertex/vertex
    r = O.restrict(e, (0, 0))
    assert r.shape == (5, 5) and r[2, 2] == 0.25
    x = np.arange(10, dtype=np.float64)[None, :].repeat(10, 0)  # cell/cell
    r = O.restrict(x, (1, 1))
    np.testing.assert_allclose(r[0], [0.5 * 0 + 0.375 * 1 + 0.125 * 2] + [0.125 * (2 * i - 1) + 0.375 * (2 * i) + 0.375 * (2 * i + 1) + 0.125 * (2 * i + 2) for i in range(1, 4)] + [0.125 * 7 + 0.375 * 8 + 0.5 * 9])
    c = np.arange(5, dtype=np.float64)[None, :].repeat(5, 0)
    p = O.interpolate(c, (1, 1))
    np.testing.assert_allclose(p[0], [0, 0.75 * 0 + 0.25 * 1, 0.75 * 1 + 0.25 * 0, 0.75 * 1 + 0.25 * 2, 0.75 * 2 + 0.25 * 1, 0.75 * 2 + 0.25 * 3,
                                      0.75 * 3 + 0.25 * 2, 0.75 * 3 + 0.25 * 4, 0.75 * 4 + 0.25 * 3, 4])
    pv = O.interpolate(c, (0, 0))
    np.testing.assert_allclose(pv[0], [0, .5, 1, 1.5, 2, 2.5, 3, 3.5, 4])


# ---------------------------------------------------------------------------- operator rows (GenerateDCA, GridsHierarchy.hxx:298-516)
@pytest.mark.parametrize("shape,sp", [((20, 22), (1.0, 1.0)), ((21, 18), (0.7, 1.3)), ((12, 14, 13), (0.3125, 0.3125, 0.5))])
def test_operator_rows_closed_form(shape, sp):
    dim = len(shape)
    dt = 0.1
    T = random_spd_tensor(shape, seed=2).astype(np.float64)
    o = O.Oracle(shape, sp, T, dt)
    S = o.stencil(0)
    np.testing.assert_allclose(S.sum(-1), 1.0, atol=1e-12)  # rows sum to one, Neumann folding included
    c = 4 if dim == 2 else 13
    # interior voxel against SURVEY appendix A
    idx = tuple(n // 2 for n in shape)
    row = S[idx]
    Dxx = T[idx][0]; Dyy = T[idx][2 if dim == 2 else 3]
    w = [dt / (h * h) for h in sp]
    if dim == 2:
        assert abs(row[c] - (1 + 2 * (w[0] * Dxx + w[1] * Dyy))) < 1e-12
        Dxy = T[idx][1]
        cxy = dt * Dxy / (2 * sp[0] * sp[1])
        assert abs(row[8] + cxy) < 1e-12 and abs(row[0] + cxy) < 1e-12 and abs(row[2] - cxy) < 1e-12 and abs(row[6] - cxy) < 1e-12
        assert abs((row[5] + row[3]) + 2 * w[0] * Dxx) < 1e-12
    else:
        Dzz = T[idx][5]
        assert abs(row[c] - (1 + 2 * (w[0] * Dxx + w[1] * Dyy + w[2] * Dzz))) < 1e-12
        for corner in (0, 2, 6, 8, 18, 20, 24, 26):
            assert np.all(S[..., corner] == 0)  # 19-point: corners inactive (:493-513)
    # boundary voxel: every entry that would leave the grid is exactly zero
    assert np.all(S[(0,) * dim][[k for k in range(3 ** dim) if k % 3 == 0]] == 0)  # x-1 at x = 0


def test_constant_tensor_of_the_2d_tests_has_no_cross_terms():
    img = load_lena()
    T = np.zeros(img.shape + (3,)); T[..., 0] = 50.0; T[..., 2] = 30.0  # test/itk2DDiffusionTest_WJ.cxx:66-73
    o = O.Oracle(img.shape, (1.0, 1.0), T, 0.1)
    S = o.stencil(0)
    assert np.all(S[..., [0, 2, 6, 8]] == 0)
    np.testing.assert_allclose(S[100, 100], [0, -3.0, 0, -5.0, 17.0, -5.0, 0, -3.0, 0])
    assert o.nlevels == 7


# ---------------------------------------------------------------------------- smoothers, direct solver, cycles
@pytest.mark.parametrize("shape,sp", [((24, 26), (1.0, 1.0)), ((14, 13, 15), (0.3125, 0.3125, 0.5))])
def test_smoothers_and_direct_solver(shape, sp):
    T = random_spd_tensor(shape, seed=4).astype(np.float64)
    f = random_image(shape, seed=5).astype(np.float64)
    for sm in (0, 1):
        o = O.Oracle(shape, sp, T, 0.1, smoother=sm)
        S = o.stencil(0)
        u1 = o.smooth(0, f, f)
        if sm == 1:  # weighted Jacobi closed form, WeightedJacobiSmoother.hxx:88-89
            c = 3 ** len(shape) // 2
            r = o.residual(0, f, f)
            np.testing.assert_allclose(u1, f + (2.0 / 3.0) * r / S[..., c], rtol=1e-12, atol=1e-10)
        r0, r1 = O.l2norm(o.residual(0, f, f)), O.l2norm(o.residual(0, u1, f))
        assert r1 < r0
        L = o.nlevels - 1
        fl = random_image(o.levels[L]["shape"], seed=6).astype(np.float64)
        e = o.direct_solve(fl)
        assert O.l2norm(o.residual(L, e, fl)) < 1e-10 * O.l2norm(fl)


def test_lena_wj_and_gs_converge_like_the_survey_says():
    """SURVEY.md section 6 [scratch]: V(2,2) on lena to 1e-10 takes 16 (WJ) / 9 (lexicographic GS) cycles."""
    img = load_lena().astype(np.float64)
    T = np.zeros(img.shape + (3,)); T[..., 0] = 50.0; T[..., 2] = 30.0
    res = {}
    for sm in (1, 0):
        o = O.Oracle(img.shape, (1.0, 1.0), T, 0.1, smoother=sm, nu=2)
        out, cyc, hist = o.solve(img, tolerance=1e-10)
        res[sm] = out
        assert cyc[0] == (16 if sm == 1 else 9), cyc
        assert hist[0, cyc[0] - 1] <= 1e-10 < hist[0, cyc[0] - 2]
    assert rel_l2(res[0], res[1]) < 1e-9  # both smoothers reach the same fixed point
    # FMG needs fewer cycles afterwards; SMOOTHER mode runs into MaxCycles like the reference's `s` tests
    o = O.Oracle(img.shape, (1.0, 1.0), T, 0.1, smoother=1, nu=2)
    _, cyc_fmg, _ = o.solve(img, cycle=O.Oracle.FMG, tolerance=1e-10)
    assert cyc_fmg[0] < 16
    _, cyc_s, hist_s = o.solve(img, cycle=O.Oracle.SMOOTHER, tolerance=1e-10, max_cycles=20)
    assert cyc_s[0] == 20 and hist_s[0, 19] < hist_s[0, 0]


def test_faithful_and_lean_vcycles_agree():
    """The reference recomputes residual + norm after every sweep for logging (…Filter.hxx:384-411); dropping
    those passes must not change the iterate."""
    shape, sp = (18, 20, 19), (0.3125, 0.3125, 0.5)
    T = random_spd_tensor(shape, seed=7).astype(np.float64)
    f = random_image(shape, seed=8).astype(np.float64)
    for sm in (0, 1):
        o = O.Oracle(shape, sp, T, 0.1, smoother=sm, nu=3)
        a, b = o.vcycle(f, f, faithful=True), o.vcycle(f, f, faithful=False)
        assert np.array_equal(a, b)


def test_ved_configuration_runs_and_time_steps_chain():
    """itkVEDTest_GS.cxx settings on the reference's own volume: 4 implicit steps, each restarting from the previous image."""
    from multigridanisotropicdiffusion_b200 import phantom
    img, sp = load_ved_test()
    _, D = phantom.vessel_phantom(img.shape, spacing=sp)
    T = phantom.planes_to_aos(D).numpy().astype(np.float64)
    o = O.Oracle(img.shape, sp, T, 0.1, smoother=0, nu=3)
    assert [L["n"] for L in o.levels] == [(69, 77, 69), (35, 39, 35), (18, 20, 18), (9, 10, 9)]
    out4, cyc4, _ = o.solve(img.astype(np.float64), tolerance=1e-10, number_of_steps=4)
    out1, cyc1, _ = o.solve(img.astype(np.float64), tolerance=1e-10, number_of_steps=1)
    again, _, _ = o.solve(out1, tolerance=1e-10, number_of_steps=3)
    assert rel_l2(again, out4) < 1e-9
    assert all(c <= 8 for c in cyc4)
    # diffusion conserves nothing exactly here (non-symmetric operator) but must smooth: total variation drops
    tv = lambda a: np.abs(np.diff(a, axis=0)).sum() + np.abs(np.diff(a, axis=1)).sum() + np.abs(np.diff(a, axis=2)).sum()
    assert tv(out4) < tv(img.astype(np.float64))
