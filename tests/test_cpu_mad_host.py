"""The solver source itself -- multigridanisotropicdiffusion_b200/csrc/madgpu.cu with mad_kernels.cuh and mad_fast.cuh, UNMODIFIED --
executed on the CPU and compared with the oracle.  tests/mad_host/madgpu_host.cpp compiles it for the host: every CUDA thread of a
block is a fibre, switched at __syncthreads / warp shuffles / named barriers (tests/mad_host/fiber_shim.h), the CUDA runtime is
host memory (tests/fake_cuda/).  The product's own Python binding (MadSolver) is pointed at that build for the duration of a test,
so these read like the GPU parity tests: hierarchy, operator evaluation, both smoothers (generic multicolour kernels and the
streaming kernels with packed fp16 rows), residuals, transfers, coarse solve, V-cycle / FMG drivers, the output casts.

A check of the code, not a CPU path of the product: libmadgpu.so still refuses to run without a GPU (tests/test_cpu_host.py).
Volumes are tiny on purpose (a fibre switch costs ~1 us); their coarsest grids stay below the 2048 unknowns of the dense inverse."""
import os
import sys

import numpy as np
import pytest

from oracle import oracle as O
from util import ROOT, random_image, random_spd_tensor, rel_l2

sys.path.insert(0, os.path.join(ROOT, "tests", "mad_host"))
import hostlib  # noqa: E402


@pytest.fixture(scope="module")
def host():
    return hostlib.load()


@pytest.fixture
def MadSolver(host, monkeypatch):
    from multigridanisotropicdiffusion_b200 import _lib as B
    from multigridanisotropicdiffusion_b200 import solver
    monkeypatch.setattr(B, "_lib", host)
    return solver.MadSolver


@pytest.mark.parametrize("shape,sp", [((33, 48), (0.7, 1.3)), ((13, 12, 15), (1.0, 0.5, 2.0))])
@pytest.mark.parametrize("smoother", ["wj", "gs"])
def test_generic_kernels_solve_matches_oracle(MadSolver, shape, sp, smoother):
    """2-D and 3-D, vertex and cell centring, cross terms: the kernels of mad_kernels.cuh through whole solves."""
    sm = 1 if smoother == "wj" else 0
    T, img = random_spd_tensor(shape, seed=2), random_image(shape, seed=5)
    with MadSolver(shape, sp, time_step=0.1, smoother=sm, iterations_per_grid=2, tolerance=1e-9, max_cycles=40, number_of_steps=2) as s:
        s.set_tensor(T)
        out = s.solve(img, out_dtype=np.float64)
        st = s.last_stats
    o = O.Oracle(shape, sp, T.astype(np.float64), 0.1, smoother=sm, nu=2)
    ref, cyc, _ = o.solve(img.astype(np.float64), tolerance=1e-9, max_cycles=40, number_of_steps=2)
    assert max(st["final_relres"]) <= 1e-9
    assert all(abs(a - b) <= (0 if smoother == "wj" else 2) for a, b in zip(st["cycles_per_step"], cyc))
    assert rel_l2(out, ref) < (1e-5 if smoother == "wj" else 1e-4)  # BASELINE.json tolerances; in fact ~1e-9
    assert rel_l2(out, ref) < 1e-7


@pytest.mark.parametrize("smoother", ["wj", "gs"])
def test_streaming_kernels_solve_matches_oracle(MadSolver, host, monkeypatch, smoother):
    """The tuned kernels of mad_fast.cuh (4 voxels per thread, z-marching in registers, warp shuffles, packed fp16 operator rows for
    Gauss-Seidel, row-pair sweep) forced onto a 64-voxel-wide volume."""
    monkeypatch.setenv("MADGPU_FAST_MIN_NX", "8")
    sm = 1 if smoother == "wj" else 0
    shape, sp = (12, 16, 64), (0.3125, 0.3125, 0.5)
    T, img = random_spd_tensor(shape, seed=2), random_image(shape, seed=5)
    before = host.mad_host_switches()
    with MadSolver(shape, sp, time_step=0.1, smoother=sm, iterations_per_grid=3, tolerance=1e-7, max_cycles=40) as s:
        assert (s.gs_tile(0) is not None) == (smoother == "gs")  # the fused sweep, not one pass per colour
        s.set_tensor(T)
        out = s.solve(img, out_dtype=np.float64)
        st = s.last_stats
    assert host.mad_host_switches() - before > 10000  # the shuffles really went through the fibre scheduler
    o = O.Oracle(shape, sp, T.astype(np.float64), 0.1, smoother=sm, nu=3)
    ref, cyc, _ = o.solve(img.astype(np.float64), tolerance=1e-7, max_cycles=40)
    assert st["final_relres"][0] <= 1e-7 and abs(st["cycles_per_step"][0] - cyc[0]) <= (0 if smoother == "wj" else 2)
    assert rel_l2(out, ref) < (1e-5 if smoother == "wj" else 1e-4)


def test_packed_rows_with_a_large_diagonal(MadSolver, monkeypatch):
    """Small spacings (SI units) make diag = 1 + 2 dt sum D_dd / h^2 large.  The packed Gauss-Seidel rows keep 1/diag scaled by
    2^14 (a plain fp16 1/diag is subnormal beyond 1.6e4 and zero beyond 3.4e7 -> a NaN image with rc 0).  h = 2e-3: diag ~ 1e4..4e5."""
    monkeypatch.setenv("MADGPU_FAST_MIN_NX", "8")
    h = 2e-3
    shape, sp = (12, 16, 64), (h, h, 1.6 * h)
    T, img = random_spd_tensor(shape, seed=3), random_image(shape, seed=6)
    with MadSolver(shape, sp, time_step=0.1, smoother=0, iterations_per_grid=3, tolerance=1e-8, max_cycles=60) as s:
        s.set_tensor(T)
        out = s.solve(img, out_dtype=np.float64)
        st = s.last_stats
        assert s.gs_tile(0)[1] in (2, 8)  # the row-pair packed sweep (tiles of 128 x 8, or 128 x 2 with warp-private tiles)
    o = O.Oracle(shape, sp, T.astype(np.float64), 0.1, smoother=0, nu=3)
    ref, cyc, _ = o.solve(img.astype(np.float64), tolerance=1e-8, max_cycles=60)
    assert np.isfinite(out).all() and st["final_relres"][0] <= 1e-8
    assert st["cycles_per_step"][0] <= cyc[0] + 2  # a stiff system (dt / h^2 = 2.5e4): ~20 cycles either way
    assert rel_l2(out, ref) < 1e-4


def test_diagonal_beyond_the_packed_range_falls_back_or_fails_loudly(MadSolver, monkeypatch):
    """dt / h^2 = 2.5e8: diag > 1e8 leaves the range of the packed rows -> the level is relaxed with exact rows (128 x 4 tiles).
    fp32 cycles cannot solve a system this stiff (the fp64 reference can); what is checked is that the call never hands back
    NaNs with rc 0: either finite numbers or MADGPU_ENUMERIC."""
    from multigridanisotropicdiffusion_b200 import MadGpuError
    monkeypatch.setenv("MADGPU_FAST_MIN_NX", "8")
    h = 2e-5
    shape, sp = (12, 16, 64), (h, h, 1.6 * h)
    T, img = random_spd_tensor(shape, seed=3), random_image(shape, seed=6)
    with MadSolver(shape, sp, time_step=0.1, smoother=0, iterations_per_grid=3, tolerance=1e-8, max_cycles=5) as s:
        s.set_tensor(T)
        try:
            out = s.solve(img, out_dtype=np.float64)
            assert np.isfinite(out).all()
        except MadGpuError as e:
            assert "not finite" in str(e)
        assert s.gs_tile(0)[1] == 4


@pytest.mark.parametrize("smoother,cycle", [("gs", 0), ("wj", 0), ("gs", 1)])
def test_captured_coarse_cycle_is_the_same_cycle(MadSolver, monkeypatch, smoother, cycle):
    """The launch-bound part of a V-cycle (levels of <= 128^3 voxels) is captured into a CUDA graph on its second use and replayed.
    Same kernels, same arguments: the solve must be bit-identical to the one with graphs switched off (MADGPU_GRAPH_VOXELS=0)."""
    monkeypatch.setenv("MADGPU_FAST_MIN_NX", "8")
    sm = 1 if smoother == "wj" else 0
    shape, sp = (12, 16, 64), (0.3125, 0.3125, 0.5)
    T, img = random_spd_tensor(shape, seed=2), random_image(shape, seed=5)
    outs, stats = [], []
    for voxels in ("0", None):
        if voxels is None:
            monkeypatch.delenv("MADGPU_GRAPH_VOXELS")
        else:
            monkeypatch.setenv("MADGPU_GRAPH_VOXELS", voxels)
        with MadSolver(shape, sp, time_step=0.1, smoother=sm, iterations_per_grid=2, cycle=cycle, tolerance=1e-8, max_cycles=40, number_of_steps=2) as s:
            s.set_tensor(T)
            outs.append(s.solve(img, out_dtype=np.float64))
            stats.append(s.last_stats)
            s.set_tensor(T * np.float32(1.5))  # a new tensor drops the graphs (they read the packed rows of the old one)
            outs.append(s.solve(img, out_dtype=np.float64))
            stats.append(s.last_stats)
    assert stats[0]["graph_launches"] == 0 and stats[1]["graph_launches"] == 0
    assert stats[2]["graph_launches"] >= stats[2]["total_cycles"] - 2 and stats[3]["graph_launches"] >= stats[3]["total_cycles"] - 2
    for a, b in ((0, 2), (1, 3)):
        assert stats[a]["cycles_per_step"] == stats[b]["cycles_per_step"] and stats[a]["kernel_launches"] == stats[b]["kernel_launches"]
        assert np.array_equal(outs[a], outs[b])
    assert not np.array_equal(outs[0], outs[1])


@pytest.mark.parametrize("shape,sp", [((12, 14, 12), (0.3125, 0.3125, 0.5)), ((13, 12), (0.7, 1.3))])
def test_device_inverse_of_the_coarsest_operator(MadSolver, monkeypatch, shape, sp):
    """The coarsest-grid direct solver (mad/itkDirectSolver.hxx:32-147): operator assembled and inverted on the device (k_coarse_matrix,
    Gauss-Jordan with partial pivoting) against the oracle's dense LU, and against the host LU path it replaced."""
    T = random_spd_tensor(shape, seed=4)
    o = O.Oracle(shape, sp, T.astype(np.float64), 0.1, smoother=1, nu=2)
    res = {}
    for host in ("0", "1"):
        monkeypatch.setenv("MADGPU_COARSE_HOST", host)
        with MadSolver(shape, sp, time_step=0.1, smoother=1, iterations_per_grid=2) as s:
            s.set_tensor(T)
            fl = random_image(s.levels[-1]["shape"], seed=3)
            res[host] = s.op_coarse_solve(fl)
    ref = o.direct_solve(fl.astype(np.float64))
    assert rel_l2(res["0"], ref) < 1e-6 and rel_l2(res["1"], ref) < 1e-6
    assert rel_l2(res["0"], res["1"]) < 1e-6
    # a second tensor on the same context replays the captured elimination on the new operator
    monkeypatch.setenv("MADGPU_COARSE_HOST", "0")
    T2 = random_spd_tensor(shape, seed=9)
    with MadSolver(shape, sp, time_step=0.1, smoother=1, iterations_per_grid=2) as s:
        s.set_tensor(T)
        first = s.op_coarse_solve(fl)
        s.set_tensor(T2)
        second = s.op_coarse_solve(fl)
    assert np.array_equal(first, res["0"])
    o2 = O.Oracle(shape, sp, T2.astype(np.float64), 0.1, smoother=1, nu=2)
    assert rel_l2(second, o2.direct_solve(fl.astype(np.float64))) < 1e-6
    assert rel_l2(second, first) > 1e-3


def test_operators_and_casts(MadSolver):
    """Per-operator entry points on a mixed-centring hierarchy, FMG, smoother-only mode, integer pixels."""
    shape, sp = (14, 25, 12), (0.33, 0.33, 0.33)
    T = random_spd_tensor(shape, seed=2)
    o = O.Oracle(shape, sp, T.astype(np.float64), 0.1, smoother=1, nu=2)
    with MadSolver(shape, sp, time_step=0.1, smoother=1, iterations_per_grid=2, tolerance=1e-9, max_cycles=30) as s:
        s.set_tensor(T)
        assert [l["shape"] for l in s.levels] == [l["shape"] for l in o.levels]
        for l in range(s.nlevels):
            np.testing.assert_allclose(np.moveaxis(s.op_get_tensor(l), 0, -1).reshape(-1), np.moveaxis(o.tensor(l), 0, -1).reshape(-1), rtol=2e-6, atol=1e-6)
            A = s.op_assemble(l).astype(np.float64)
            np.testing.assert_allclose(A, o.stencil(l), rtol=0, atol=2e-5 * np.abs(o.stencil(l)).max())
            u, f = random_image(s.levels[l]["shape"], seed=l + 1), random_image(s.levels[l]["shape"], seed=l + 9)
            assert rel_l2(s.op_smooth(l, u, f, smoother=1), o.smooth(l, u.astype(np.float64), f.astype(np.float64))) < 2e-6
            r, nrm = s.op_residual(l, u, f)
            ro = o.residual(l, u.astype(np.float64), f.astype(np.float64))
            assert rel_l2(r, ro) < 2e-5 and abs(nrm - np.linalg.norm(ro)) < 1e-4 * np.linalg.norm(ro)
        fine = random_image(shape, seed=4)
        cent = o.levels[1]["centering"]
        c = s.op_restrict(0, fine)
        assert rel_l2(c, O.restrict(fine.astype(np.float64), cent)) < 1e-6
        assert rel_l2(s.op_prolong(0, c), O.interpolate(c.astype(np.float64), cent, shape)) < 1e-6
        fl = random_image(s.levels[-1]["shape"], seed=3)
        assert rel_l2(s.op_coarse_solve(fl), o.direct_solve(fl.astype(np.float64))) < 1e-5
        img = np.round(random_image(shape, seed=5) - 100.0)
        for cycle in (s.FMG, s.SMOOTHER):
            s.set_solver(cycle=cycle, max_cycles=12 if cycle == s.SMOOTHER else 30)
            out = s.solve(img.astype(np.float32), out_dtype=np.float64)
            ref, cyc, _ = o.solve(img, cycle=cycle, tolerance=1e-9, max_cycles=12 if cycle == s.SMOOTHER else 30)
            assert abs(s.last_stats["cycles_per_step"][0] - cyc[0]) <= 1 and rel_l2(out, ref) < 1e-5
        s.set_solver(cycle=s.VCYCLE, max_cycles=30)
        o16 = s.solve(img.astype(np.int16))  # static_cast< short >: truncation toward zero, negative values included
        ref, _, _ = o.solve(img, tolerance=1e-9, max_cycles=30)
        d = np.abs(o16.astype(int) - np.trunc(ref).astype(int))
        assert o16.dtype == np.int16 and d.max() <= 1 and (d != 0).mean() < 5e-3


def test_contexts_release_device_memory(MadSolver, host):
    before = host.mad_host_live_allocs()
    shape = (12, 14, 16)
    with MadSolver(shape, (1, 1, 1), time_step=0.1) as s:
        s.set_tensor(random_spd_tensor(shape, seed=1))
        s.solve(random_image(shape))
        assert host.mad_host_live_allocs() > before
    assert host.mad_host_live_allocs() == before


@pytest.mark.parametrize("cycle", ["v", "fmg"])
@pytest.mark.parametrize("smoother", ["wj", "gs"])
@pytest.mark.parametrize("tag,shape,sp", [("small2d", (49, 33), (0.7, 1.3)), ("small3d", (23, 25, 27), (0.3125, 0.3125, 0.5))])
def test_golden_vectors_of_the_reference_code(MadSolver, tag, shape, sp, smoother, cycle):
    """The CUDA source against the vectors recorded from the reference's own code (tests/golden/make_golden.py): cross terms,
    mixed vertex / cell centring, two time steps -- the same assertions as tests/test_gpu_golden.py makes on the GPU."""
    from util import GOLDEN
    g = np.load(os.path.join(GOLDEN, f"ref_{tag}_{smoother}_{cycle}.npz"))
    T, img = random_spd_tensor(shape, seed=2), random_image(shape, seed=5)
    with MadSolver(shape, sp, time_step=0.1, smoother=1 if smoother == "wj" else 0, iterations_per_grid=2, cycle=1 if cycle == "fmg" else 0,
                   tolerance=1e-10, max_cycles=100, number_of_steps=2) as s:
        s.set_tensor(T)
        out = s.solve(img, out_dtype=np.float64)
        st = s.last_stats
    assert all(r <= 1e-10 for r in st["final_relres"][:2])
    assert rel_l2(out, g["sample"]) < 1e-6
    for a, b in zip(st["cycles_per_step"], g["cycles"]):
        assert abs(a - int(b)) <= (1 if smoother == "wj" else 3)


# ---- the VED filter end to end: front-end kernels + the real solver, both from the product's CUDA source ------------------------
def _ved_sub_volume():
    from util import load_ved_test
    img, sp = load_ved_test()
    return np.ascontiguousarray(img[20:44, 24:52, 18:48]), sp  # (24, 28, 30) around vessels


@pytest.mark.parametrize("pixel,smoother,cycle", [(np.int16, "gs", 0), (np.float64, "wj", 1)])
def test_whole_ved_filter_on_the_emulated_device(host, monkeypatch, pixel, smoother, cycle):
    """VEDMultigridImageFilter.Update() -- madved_run: Hessians, vesselness, tensor handed to the solver in "device" memory
    (madgpu_set_tensor_device_f32), DiffusionStep in place (madgpu_solve_device_f32), output cast from the fp64 iterate
    (madgpu_fetch_output) -- against the oracle's GenerateData, with the parameters of test/itkVEDTest_GS.cxx."""
    from multigridanisotropicdiffusion_b200 import _lib as B
    import multigridanisotropicdiffusion_b200 as M
    from oracle import ved as V
    monkeypatch.setattr(B, "_lib", host)
    img, sp = _ved_sub_volume()
    img = img.astype(pixel)
    f = M.VEDMultigridImageFilter(smoother)
    f.SetCycle(cycle)
    f.SetDiffusionIterationsPerGrid(3)
    f.SetInput(img, sp)
    f.SetScales([0.300, 0.482, 0.775, 1.245, 2.000])
    f.SetAlpha(0.5); f.SetBeta(0.5); f.SetGamma(5.0); f.SetEpsilon(0.01); f.SetSensitivity(10.0)
    f.SetIterations(2)
    f.SetTolerance(1e-9)
    f.SetTimeStep(0.1)
    f.SetDiffusionIterations(2)
    f.SetOmega(1.5)
    f.Update()
    out = f.GetOutput()
    assert out.dtype == pixel and f.ved_stats["scales"] == 10 and f.stats["steps"] == 2 and max(f.stats["final_relres"]) <= 1e-9
    want, info = V.ved_filter(img, sp, V.DEFAULT_SCALES, alpha=0.5, beta=0.5, gamma=5.0, epsilon=0.01, omega=1.5, sensitivity=10.0, iterations=2,
                              diffusion_iterations=2, smoother=0 if smoother == "gs" else 1, cycle=cycle, time_step=0.1, tolerance=1e-9,
                              iterations_per_grid=3, out_dtype=pixel)
    if pixel == np.int16:
        d = np.abs(out.astype(int) - want.astype(int))
        assert d.max() <= 1 and (d != 0).mean() < 1e-3
    else:
        assert rel_l2(out, want) < 1e-6
    assert host.mad_host_live_allocs() == 0


def test_fetch_output_casts(MadSolver):
    """madgpu_fetch_output: the current fp64 iterate cast to every pixel type (…Filter.hxx:267-284), after a device-buffer solve."""
    import ctypes as C
    from multigridanisotropicdiffusion_b200 import _lib as B
    shape = (12, 14, 16)
    img = np.round(random_image(shape, seed=5) - 60.0).astype(np.float32)
    with MadSolver(shape, (1, 1, 1), time_step=0.1, tolerance=1e-9) as s:
        s.set_tensor(random_spd_tensor(shape, seed=1))
        ref = s.solve(img, out_dtype=np.float64)
        work = img.copy()  # "device" memory of the host build is host memory
        s.solve_device(work.ctypes.data, work.ctypes.data)
        np.testing.assert_array_equal(work, ref.astype(np.float32))
        for dt, code in ((np.float64, B.PIX_F64), (np.float32, B.PIX_F32), (np.int16, B.PIX_I16), (np.uint8, B.PIX_U8)):
            out = np.empty(shape, dtype=dt)
            assert s._lib.madgpu_fetch_output(s._ctx, code, C.c_void_p(out.ctypes.data)) == 0
            if dt == np.uint8:
                pos = ref >= 0
                np.testing.assert_array_equal(out[pos], np.trunc(ref[pos]).astype(np.uint8))
            else:
                np.testing.assert_array_equal(out, np.trunc(ref).astype(dt) if dt == np.int16 else ref.astype(dt))


def _run_gpu_tests_on_the_emulated_device(*pytest_args):
    import subprocess
    env = dict(os.environ, MADGPU_EMULATED_DEVICE="1")
    r = subprocess.run([sys.executable, "-m", "pytest", *pytest_args, "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider"], capture_output=True,
                       text=True, env=env, cwd=ROOT, timeout=1800)
    assert r.returncode == 0 and " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


def test_gpu_operator_tests_on_the_emulated_device():
    """tests/test_gpu_ops.py -- the per-operator GPU parity tests, green on a B200 -- run unchanged against the host build of the
    CUDA source (MADGPU_EMULATED_DEVICE=1, tests/conftest.py): the emulation and the GPU agree on what passes."""
    _run_gpu_tests_on_the_emulated_device(os.path.join(ROOT, "tests", "test_gpu_ops.py"))


def test_gpu_edge_case_tests_on_the_emulated_device():
    """The cheap edge cases of tests/test_gpu_solve.py: error paths, a single-level volume, a zero right-hand side, every output
    pixel cast.  (The whole file passes on the emulation too, in about a quarter of an hour.)"""
    _run_gpu_tests_on_the_emulated_device(os.path.join(ROOT, "tests", "test_gpu_solve.py"), "-k", "error_paths or single_level or pixel_casts")


@pytest.mark.parametrize("smoother", ["wj", "gs"])
def test_fp64_residual_with_fp32_operator_rows(MadSolver, monkeypatch, smoother):
    """MADGPU_RES64_COEF32=1 (opt-in, k_fast_sweep<MODE_RES_C32>): the level-0 stop-test residual applies, in fp64, the operator
    row evaluated in fp32 -- the row the fp32 sweeps use -- instead of re-evaluating it in fp64.  The residual of an arbitrary
    iterate then differs from the fp64-row residual by the rounding of the row (~1e-7 relative to |A||u|), the converged image by
    about as much, cycle counts not at all."""
    monkeypatch.setenv("MADGPU_FAST_MIN_NX", "8")
    sm = 1 if smoother == "wj" else 0
    shape, sp = (12, 16, 64), (0.3125, 0.3125, 0.5)
    T, img = random_spd_tensor(shape, seed=2), random_image(shape, seed=5)
    o = O.Oracle(shape, sp, T.astype(np.float64), 0.1, smoother=sm, nu=3)
    ref, cyc, _ = o.solve(img.astype(np.float64), tolerance=1e-9, max_cycles=40)
    u, f = random_image(shape, seed=3).astype(np.float64), random_image(shape, seed=4).astype(np.float64)
    r_o = o.residual(0, u, f)
    scale = np.abs(u).max() * np.abs(o.stencil(0)).sum(-1).max()
    res, outs, cycles = {}, {}, {}
    for flag in ("0", "1"):
        monkeypatch.setenv("MADGPU_RES64_COEF32", flag)
        with MadSolver(shape, sp, time_step=0.1, smoother=sm, iterations_per_grid=3, tolerance=1e-9, max_cycles=40) as s:
            s.set_tensor(T)
            res[flag], nrm = s.op_residual_f64(u, f, norm_only=False)
            _, nrm_fast = s.op_residual_f64(u, f, norm_only=True)  # the streaming kernel of the solve loop
            assert abs(nrm_fast - np.linalg.norm(r_o)) < (1e-12 if flag == "0" else 3e-7) * np.linalg.norm(r_o) * (1 if flag == "0" else scale / np.linalg.norm(r_o) * np.sqrt(u.size))
            outs[flag] = s.solve(img, out_dtype=np.float64)
            cycles[flag] = s.last_stats["cycles_per_step"]
            assert s.last_stats["final_relres"][0] <= 1e-9
    assert cycles["0"] == cycles["1"] and abs(cycles["1"][0] - cyc[0]) <= (0 if smoother == "wj" else 2)
    assert rel_l2(outs["1"], outs["0"]) < 5e-7 and rel_l2(outs["1"], ref) < 1e-6


def test_continuing_from_the_fp64_result(MadSolver):
    """madgpu_solve_device_f32 with d_in == NULL: two time steps in one call == one time step + one continued time step, bit for bit
    (the fp64 iterate is the carrier, not its fp32 copy); without a previous solve it is a state error."""
    from multigridanisotropicdiffusion_b200 import MadGpuError
    shape, sp = (12, 14, 12), (0.3125, 0.3125, 0.5)
    T, img = random_spd_tensor(shape, seed=4), np.ascontiguousarray(random_image(shape, seed=6), dtype=np.float32)
    kw = dict(time_step=0.1, smoother=0, iterations_per_grid=2, tolerance=1e-9, max_cycles=40)
    out2 = np.empty(shape, np.float32)
    with MadSolver(shape, sp, number_of_steps=2, **kw) as s:
        s.set_tensor(T)
        s.solve_device(img.ctypes.data, out2.ctypes.data)  # the emulated device's memory is host memory
    a, b = np.empty(shape, np.float32), np.empty(shape, np.float32)
    with MadSolver(shape, sp, number_of_steps=1, **kw) as s:
        s.set_tensor(T)
        with pytest.raises(MadGpuError):
            s.solve_device(None, a.ctypes.data)
        s.solve_device(img.ctypes.data, a.ctypes.data)
        s.solve_device(None, b.ctypes.data)
        # restarting from the fp32 copy instead is a different (rounded) right-hand side
        c = np.empty(shape, np.float32)
        s.solve_device(a.ctypes.data, c.ctypes.data)
    assert np.array_equal(b, out2)
    assert rel_l2(c, out2) < 1e-6
