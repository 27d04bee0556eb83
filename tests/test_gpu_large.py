"""Parity at BASELINE.json's sizes against the oracle (itkMultigridAnisotropicDiffusionImageFilter.hxx:207-246 restated in
oracle/mad_oracle.c): configs[2] = 256^3 and a 512 x 512 x 64 slab of configs[3] (as many tiles of the fused Gauss-Seidel sweep
per plane as 512^3 has, z-chunks and 32-bit offsets far beyond the small cases of test_gpu_fast.py / test_gpu_solve.py).

The oracle is a scalar fp64 port (~1 Mvoxel/s per pass): its solves run in worker PROCESSES started when the module is first used
-- one per case, side by side on the host cores -- while the GPU does its part; ~2 minutes of wall time in all.

  weighted Jacobi : ONE V(3,3) cycle from the same iterate, rel-L2 <= 1e-5 (north_star's per-V-cycle bound)
  Gauss-Seidel    : one time step solved to relres 1e-8 on both sides, rel-L2 of the converged image <= 1e-4
Size-independent properties at the full 512^3 are in test_gpu_props.py."""
import concurrent.futures as cf
import multiprocessing as mp

import numpy as np
import pytest

from util import rel_l2

pytestmark = pytest.mark.gpu

import os

CASES = {"256": (256, 256, 256), "slab512": (64, 512, 512)}  # (nz, ny, nx)
if os.environ.get("MADGPU_LARGE_TEST_DRYRUN") == "1":  # logic check of this file on the CPU dry-run build (conftest.py): tiny stand-ins
    CASES = {"256": (24, 24, 64), "slab512": (24, 24, 128)}
NU, DT, GS_TOL = 3, 0.1, 1e-8


def _inputs(shape):
    from multigridanisotropicdiffusion_b200 import phantom
    img_t, D = phantom.vessel_phantom(shape)  # CPU, seeded: identical in the workers and in the test process
    return img_t.numpy(), phantom.planes_to_aos(D).numpy()


def _oracle_job(name):
    """Worker: the oracle's weighted-Jacobi V-cycle and Gauss-Seidel time step for one case."""
    from multigridanisotropicdiffusion_b200 import phantom
    from oracle import oracle as O
    shape = CASES[name]
    img, T = _inputs(shape)
    img64 = img.astype(np.float64)
    o = O.Oracle(shape, phantom.VED_SPACING, T.astype(np.float64), DT, smoother=1, nu=NU)
    wj = o.vcycle(img64.copy(), img64).astype(np.float64)  # u0 = f: the reference's initial guess (…Filter.hxx:182-199)
    o.set_smoother(0, nu=NU)
    gs, cyc, _ = o.solve(img64, tolerance=GS_TOL, max_cycles=30)
    return name, wj, gs, cyc


@pytest.fixture(scope="module")
def oracle_results():
    ctx = mp.get_context("spawn")
    pool = cf.ProcessPoolExecutor(max_workers=len(CASES), mp_context=ctx)
    futs = {n: pool.submit(_oracle_job, n) for n in CASES}
    yield futs
    pool.shutdown(wait=False, cancel_futures=True)


@pytest.mark.parametrize("name", list(CASES))
def test_baseline_size_parity(name, oracle_results):
    from multigridanisotropicdiffusion_b200 import MadSolver, phantom
    shape = CASES[name]
    img, T = _inputs(shape)
    # --- GPU: one weighted-Jacobi V(3,3) cycle of the outer loop, then the Gauss-Seidel time step ---
    with MadSolver(shape, phantom.VED_SPACING, time_step=DT, smoother=MadSolver.WJ, iterations_per_grid=NU, tolerance=0.0, max_cycles=1) as s:
        s.set_tensor(T)
        assert s.levels[0]["shape"] == shape
        s.cycles_begin(img)
        s.cycles_run(1)
        wj_gpu = s.cycles_end()
        s.set_solver(smoother=MadSolver.GS, tolerance=GS_TOL, max_cycles=30)
        gs_gpu = s.solve(img, out_dtype=np.float64)
        st = s.last_stats
        tile = s.gs_tile(0)
    assert tile is not None and tile[0] == 128  # the fused streaming sweep, not the generic colour passes
    _, wj_ref, gs_ref, cyc = oracle_results[name].result(timeout=1500)
    e_wj, e_gs = rel_l2(wj_gpu, wj_ref), rel_l2(gs_gpu, gs_ref)
    print(f"[{name}] WJ one V-cycle rel-L2 {e_wj:.3e}; GS converged rel-L2 {e_gs:.3e}, cycles gpu {st['cycles_per_step']} oracle {cyc}")
    assert e_wj <= 1e-5
    assert st["final_relres"][0] <= GS_TOL and abs(st["cycles_per_step"][0] - cyc[0]) <= 2
    assert e_gs <= 1e-4
