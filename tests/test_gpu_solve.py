"""End-to-end parity of the filter path (GenerateData replacement) against the CPU oracle on the
reference's own test configurations (BASELINE.json configs[0..2]) and on the cases its tests miss.

Tolerances (BASELINE.json north_star): weighted Jacobi <= 1e-5 relative L2 per V-cycle;
Gauss-Seidel <= 1e-4 relative L2 on the converged diffused image (multicolour ordering differs
from the reference's lexicographic sweep, so only the fixed point is comparable).
"""
import numpy as np
import pytest

from util import load_lena, load_ved_test, random_image, random_spd_tensor, rel_l2

pytestmark = pytest.mark.gpu

WJ_PER_CYCLE_TOL = 1e-5
GS_CONVERGED_TOL = 1e-4


def _lena_case():
    img = load_lena().astype(np.float32)  # the reference casts uchar -> float before the filter
    T = np.zeros(img.shape + (3,), dtype=np.float32)
    T[..., 0] = 50.0  # test/itk2DDiffusionTest_WJ.cxx:66-73
    T[..., 2] = 30.0
    return img, T, (1.0, 1.0)


def _run_filter(img, T, spacing, smoother, cycle, nu, dt=0.1, tol=1e-10, steps=1, max_cycles=100):
    import multigridanisotropicdiffusion_b200 as M
    f = M.MultigridAnisotropicDiffusionImageFilter(smoother)
    f.SetInput(img, spacing)
    f.SetDiffusionTensor(T)
    f.SetIterationsPerGrid(nu)
    f.SetTimeStep(dt)
    f.SetNumberOfSteps(steps)
    f.SetMaxCycles(max_cycles)
    f.SetTolerance(tol)
    f.SetCycle(cycle)
    f.Update()
    out, st = f.GetOutput(), f.stats
    f.close()
    return out, st


def _run_oracle(img, T, spacing, smoother, cycle, nu, dt=0.1, tol=1e-10, steps=1, max_cycles=100):
    from oracle import oracle as O
    o = O.Oracle(img.shape, spacing, T.astype(np.float64), dt, smoother=smoother, nu=nu)
    out, cyc, hist = o.solve(img.astype(np.float64), cycle=cycle, tolerance=tol, max_cycles=max_cycles, number_of_steps=steps)
    return out, cyc, hist


# ---------------------------------------------------------------------------------- config 0/1: lena 2-D
@pytest.mark.parametrize("cycle", [0, 1, 2], ids=["v", "fmg", "s"])
def test_itk2DDiffusionTest_WJ(cycle):
    """test/itk2DDiffusionTest_WJ.cxx with argv[1] in {v, fmg, s}."""
    img, T, sp = _lena_case()
    max_cycles = 100
    out, st = _run_filter(img, T, sp, "wj", cycle, nu=2, max_cycles=max_cycles)
    ref, cyc, hist = _run_oracle(img, T, sp, 1, cycle, nu=2, max_cycles=max_cycles)
    assert abs(st["cycles_per_step"][0] - cyc[0]) <= 1, (st["cycles_per_step"], cyc)
    if cycle == 2:
        # smoother-only does not converge in 100 iterations; compare the iterate itself
        assert st["cycles_per_step"][0] == 100 and cyc[0] == 100
        assert rel_l2(out, ref) < WJ_PER_CYCLE_TOL
    else:
        assert st["final_relres"][0] <= 1e-10
        assert rel_l2(out, ref) < 1e-6, rel_l2(out, ref)


def test_WJ_per_vcycle_lena():
    """One V-cycle at a time (tolerance 0, max_cycles k): the iterate after k cycles tracks the oracle's."""
    img, T, sp = _lena_case()
    from multigridanisotropicdiffusion_b200 import MadSolver
    from oracle import oracle as O
    o = O.Oracle(img.shape, sp, T.astype(np.float64), 0.1, smoother=1, nu=2)
    s = MadSolver(img.shape, sp, time_step=0.1, smoother=1, iterations_per_grid=2, tolerance=0.0, max_cycles=1)
    s.set_tensor(T)
    u = img.astype(np.float64)
    for k in range(1, 5):
        s.set_solver(max_cycles=k)
        g = s.solve(img, out_dtype=np.float64)
        u = o.vcycle(u, img.astype(np.float64))
        assert rel_l2(g, u) < WJ_PER_CYCLE_TOL, (k, rel_l2(g, u))
        # relres reported by the stop test agrees with the oracle's residual of its own iterate
        rr = np.linalg.norm(o.residual(0, u, img.astype(np.float64))) / np.linalg.norm(img.astype(np.float64))
        assert abs(s.last_stats["final_relres"][0] - rr) < 1e-3 * rr + 1e-12
    s.close()


@pytest.mark.parametrize("cycle", [0, 1], ids=["v", "fmg"])
def test_itk2DDiffusionTest_GS(cycle):
    """test/itk2DDiffusionTest_GS.cxx: multicolour GS converges to the lexicographic GS fixed point."""
    img, T, sp = _lena_case()
    out, st = _run_filter(img, T, sp, "gs", cycle, nu=2)
    ref, cyc, _ = _run_oracle(img, T, sp, 0, cycle, nu=2)
    assert st["final_relres"][0] <= 1e-10
    assert abs(st["cycles_per_step"][0] - cyc[0]) <= 3, (st["cycles_per_step"], cyc)
    assert rel_l2(out, ref) < GS_CONVERGED_TOL, rel_l2(out, ref)
    assert rel_l2(out, ref) < 1e-6  # in fact both are converged to 1e-10


# ---------------------------------------------------------------------------------- config 2: VED 3-D
def _ved_case():
    import torch
    from multigridanisotropicdiffusion_b200 import phantom
    img, sp = load_ved_test()
    _, D = phantom.vessel_phantom(img.shape, spacing=sp)
    T = phantom.planes_to_aos(D).numpy().astype(np.float64)  # VED hands a double tensor (VED.h:64)
    return img, T, sp


@pytest.mark.parametrize("smoother,cycle", [("gs", 0), ("gs", 1), ("wj", 0)])
def test_itkVEDTest_diffusion_step(smoother, cycle):
    """DiffusionStep of test/itkVEDTest_GS.cxx: int16 volume, nu=3, dt=.1, 4 time steps, tol 1e-10."""
    import multigridanisotropicdiffusion_b200 as M
    img, T, sp = _ved_case()
    f = M.VEDMultigridImageFilter(smoother)
    f.SetInput(img, sp)
    f.SetDiffusionTensor(T)
    f.SetDiffusionIterationsPerGrid(3)
    f.SetTolerance(1e-10)
    f.SetTimeStep(0.1)
    f.SetDiffusionIterations(4)
    f.SetCycle(cycle)
    out = f.Update().GetOutput()
    st = f.stats
    assert out.dtype == np.int16 and st["steps"] == 4
    ref, cyc, _ = _run_oracle(img, T, sp, 0 if smoother == "gs" else 1, cycle, nu=3, steps=4)
    assert all(r <= 1e-10 for r in st["final_relres"])
    # static_cast<short> truncates: a voxel may differ by one where the double value sits on an integer
    ref_i = np.trunc(ref).astype(np.int16)
    diff = np.abs(out.astype(np.int32) - ref_i.astype(np.int32))
    assert diff.max() <= 1
    assert (diff != 0).mean() < 1e-4
    for a, b in zip(st["cycles_per_step"], cyc):
        assert abs(a - b) <= 2, (st["cycles_per_step"], cyc)


def test_ved_double_output_matches_oracle():
    from multigridanisotropicdiffusion_b200 import MadSolver
    img, T, sp = _ved_case()
    for smoother, tol in ((1, 1e-6), (0, GS_CONVERGED_TOL)):
        s = MadSolver(img.shape, sp, time_step=0.1, smoother=smoother, iterations_per_grid=3, tolerance=1e-10,
                      number_of_steps=4)
        s.set_tensor(T)
        g = s.solve(img, out_dtype=np.float64)
        ref, _, _ = _run_oracle(img, T, sp, smoother, 0, nu=3, steps=4)
        assert rel_l2(g, ref) < tol, (smoother, rel_l2(g, ref))
        s.close()


# ---------------------------------------------------------------------------------- what the reference tests miss
@pytest.mark.parametrize("shape,sp", [((45, 47, 40), (0.330017,) * 3), ((33, 64, 21), (0.5, 0.25, 1.0)), ((97, 129), (0.7, 1.3))])
@pytest.mark.parametrize("smoother", [0, 1])
def test_mixed_centring_cross_terms(shape, sp, smoother):
    from multigridanisotropicdiffusion_b200 import MadSolver
    T = random_spd_tensor(shape, seed=11)
    img = random_image(shape, seed=12)
    s = MadSolver(shape, sp, time_step=0.1, smoother=smoother, iterations_per_grid=2, tolerance=1e-10, number_of_steps=2)
    s.set_tensor(T)
    g = s.solve(img, out_dtype=np.float64)
    ref, cyc, _ = _run_oracle(img, T, sp, smoother, 0, nu=2, steps=2)
    assert all(r <= 1e-10 for r in s.last_stats["final_relres"]), s.last_stats
    assert rel_l2(g, ref) < (1e-6 if smoother == 1 else GS_CONVERGED_TOL), rel_l2(g, ref)
    s.close()


def test_output_pixel_casts():
    """static_cast<OutputPixelType> semantics (…Filter.hxx:267-284): truncation toward zero for integers."""
    from multigridanisotropicdiffusion_b200 import MadSolver
    shape = (40, 48)
    T = random_spd_tensor(shape, seed=3)
    img8 = np.clip(random_image(shape, seed=4), 0, 255).astype(np.uint8)
    s = MadSolver(shape, (1, 1), time_step=0.1, smoother=1, tolerance=1e-10)
    s.set_tensor(T)
    d = s.solve(img8, out_dtype=np.float64)
    o8 = s.solve(img8)
    o16 = s.solve(img8.astype(np.int16))
    o32 = s.solve(img8.astype(np.float32))
    assert o8.dtype == np.uint8 and o16.dtype == np.int16 and o32.dtype == np.float32
    np.testing.assert_array_equal(o8, np.trunc(d).astype(np.uint8))
    np.testing.assert_array_equal(o16, np.trunc(d).astype(np.int16))
    np.testing.assert_allclose(o32, d.astype(np.float32), rtol=0, atol=1e-4)
    s.close()


def test_error_paths():
    from multigridanisotropicdiffusion_b200 import MadGpuError, MadSolver
    with pytest.raises(MadGpuError):
        MadSolver((2, 8, 8))  # fewer than 3 voxels on an axis
    s = MadSolver((16, 16))
    with pytest.raises(MadGpuError):
        s.solve(np.zeros((16, 16), np.float32))  # tensor not set
    with pytest.raises(MadGpuError):
        s.set_tensor(np.zeros((16, 16, 6), np.float32))  # wrong component count
    s.close()


def test_single_level_and_zero_rhs():
    """Volumes that cannot be coarsened (any axis < 12) are solved directly on level 0; a zero image gives
    relres NaN in the reference (0/0, …Filter.hxx:204,239) and one cycle."""
    from multigridanisotropicdiffusion_b200 import MadSolver
    shape = (10, 9, 11)
    T = random_spd_tensor(shape, seed=8)
    img = random_image(shape, seed=9)
    s = MadSolver(shape, (1, 1, 1), time_step=0.1, smoother=0, tolerance=1e-10)
    s.set_tensor(T)
    assert s.nlevels == 1
    g = s.solve(img, out_dtype=np.float64)
    ref, cyc, _ = _run_oracle(img, T, (1, 1, 1), 0, 0, nu=2)
    assert rel_l2(g, ref) < 1e-6
    z = s.solve(np.zeros(shape, np.float32))
    assert s.last_stats["cycles_per_step"] == [1]
    assert np.all(z == 0)
    s.close()


@pytest.mark.parametrize("smoother", ["gs", "wj"])
def test_a_context_survives_new_tensors(smoother):
    """SetDiffusionTensor on a live context (VED.hxx:381-402 calls it once per outer iteration): the packed rows are rebuilt, the captured
    coarse-level cycle and the captured Gauss-Jordan elimination are replayed on the new operator.  Tensor A, B, A again on one context:
    the third solve repeats the first bit for bit, the second equals a fresh context's."""
    from multigridanisotropicdiffusion_b200 import MadSolver
    shape, sp = (40, 48, 136), (0.3125, 0.3125, 0.5)
    sm = MadSolver.GS if smoother == "gs" else MadSolver.WJ
    TA, TB = random_spd_tensor(shape, seed=11), random_spd_tensor(shape, seed=12)
    img = random_image(shape, seed=13)
    kw = dict(time_step=0.1, smoother=sm, iterations_per_grid=3, tolerance=1e-9, max_cycles=60, number_of_steps=2)
    with MadSolver(shape, sp, **kw) as s:
        outs = []
        for T in (TA, TB, TA):
            s.set_tensor(T)
            outs.append(s.solve(img, out_dtype=np.float64))
            assert max(s.last_stats["final_relres"]) <= 1e-9
    with MadSolver(shape, sp, **kw) as s:
        s.set_tensor(TB)
        fresh = s.solve(img, out_dtype=np.float64)
    assert np.array_equal(outs[0], outs[2])
    assert np.array_equal(outs[1], fresh)
    assert rel_l2(outs[0], outs[1]) > 1e-4
