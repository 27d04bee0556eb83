#!/usr/bin/env python
"""Benchmark of the multigrid anisotropic-diffusion hot path (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU

A "step" is ONE V(nu,nu) cycle of the 3-D VED diffusion step (GS smoother, nu=3, dt=0.1, the
settings of the reference's test/itkVEDTest_GS.cxx) on a synthetic vessel volume, INCLUDING the
level-0 residual-norm stop test and its read-back, exactly one pass of the reference's do-while body
(itkMultigridAnisotropicDiffusionImageFilter.hxx:207-246).

  value  : fine voxels / device time of one cycle, inputs resident in HBM (Mvoxel/s)
  e2e    : the same metric through the filter call with HOST (pinned) buffers: one DiffusionStep
           (SetDiffusionTensor + 4 time steps to tolerance 1e-10), tensor/image H2D and result D2H
           inside the timed region; value = voxels x cycles executed / wall time
  roofline: level-0 smoother sweep, 36 B/voxel algorithmic (u, f, u', six tensor planes) over the
           CUDA-event time of those launches, against the measured HBM copy peak.  (The Gauss-Seidel sweep reads
           pre-evaluated fp16 operator rows, 20 B/voxel, instead of the 24 B of tensor planes: ncu traffic 32 B/voxel.)
  cpu_baseline / --impl reference: the reference's own code on the host cores -- oracle/_ref/libmadref.so, the unmodified
           reference headers compiled against the stand-in ITK (kind "reference"), or, where that library is absent, the oracle
           port (kind "port") -- lexicographic GS, double, one independent single-threaded solve per core (the reference has no
           threading), on a bounded sample of the same workload
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ALG_BYTES_SWEEP_3D = 36.0  # u read + f read + u write + six tensor planes, fp32


def alg_bytes_per_cycle(nu: int, dim: int = 3) -> float:
    op = 24.0 if dim == 3 else 12.0
    geo = 8.0 / 7.0 if dim == 3 else 4.0 / 3.0
    return (2 * nu * (12 + op) + (8 + op + 0.5) + 8.5) * geo + (8 + op)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_run(size: int, nu: int, smoother: int, steps: int, warmup: int):
    """The reference algorithm on the host: oracle V-cycles in `faithful` mode (the reference's redundant
    residual + norm after every sweep, …Filter.hxx:384-411,437-439,460-487) plus the outer residual/norm."""
    import numpy as np

    from multigridanisotropicdiffusion_b200 import phantom
    from oracle import oracle as O
    shape = (size, size, size)
    img_t, D = phantom.vessel_phantom(shape)
    img = img_t.numpy().astype(np.float64)
    T = phantom.planes_to_aos(D).numpy().astype(np.float64)
    t0 = time.perf_counter()
    o = O.Oracle(shape, phantom.VED_SPACING, T, 0.1, smoother=smoother, nu=nu)
    setup_s = time.perf_counter() - t0
    u = img.copy()
    rhs_norm = O.l2norm(img)
    relres = None
    for _ in range(warmup):
        u = o.vcycle(u, img, faithful=True)
    t0 = time.perf_counter()
    for _ in range(steps):
        u = o.vcycle(u, img, faithful=True)
        relres = O.l2norm(o.residual(0, u, img)) / rhs_norm
    dt = time.perf_counter() - t0
    n = size ** 3
    return dict(mvox_s=n * steps / dt / 1e6, s_per_cycle=dt / steps, setup_s=setup_s, relres=relres, voxels=n)


def cpu_reference_run_ref(size: int, nu: int, smoother: int, steps: int, warmup: int):
    """The reference's OWN code on the host: oracle/_ref/libmadref.so = the unmodified /root/reference/include headers compiled
    against the stand-in ITK of oracle/shim.  GenerateData() is run with tolerance 0 for exactly warmup + steps V-cycles (each with
    the stop-test residual + norm, …Filter.hxx:207-246); the set-up (GridsHierarchy + DirectSolver, timed separately through the
    same library) is subtracted, so the figure is cycles only, like the GPU arm."""
    import numpy as np

    from multigridanisotropicdiffusion_b200 import phantom
    from oracle import ref as R
    shape = (size, size, size)
    img_t, D = phantom.vessel_phantom(shape)
    img = img_t.numpy().astype(np.float64)
    T = phantom.planes_to_aos(D).numpy().astype(np.float64)
    t0 = time.perf_counter()
    h = R.Reference(shape, phantom.VED_SPACING, T, 0.1)
    h.direct_solve(np.zeros(h.levels[-1]["shape"]))  # forces the coarsest-grid factorisation
    setup_s = time.perf_counter() - t0
    del h
    ncyc = max(steps + warmup, 1)
    t0 = time.perf_counter()
    R.run_filter(img, phantom.VED_SPACING, T, smoother=smoother, cycle=0, nu=nu, time_step=0.1, tolerance=0.0, max_cycles=ncyc,
                 number_of_steps=1, pixel="double", verbose=False)
    total = time.perf_counter() - t0
    per = max(total - setup_s, 1e-9) / ncyc
    n = size ** 3
    return dict(mvox_s=n / per / 1e6, s_per_cycle=per, setup_s=setup_s, relres=None, voxels=n, kind="reference")


def cpu_impl_kind(requested="auto"):
    """'reference' (oracle/_ref: the compiled reference headers) when that library is present, else 'port' (the C restatement)."""
    if requested in ("reference", "port"):
        return requested
    try:
        from oracle import ref as R
        if R.available():
            R.lib()  # loadable here? (built in the authoring container, travels as an artefact)
            return "reference"
    except Exception:  # noqa: BLE001
        pass
    return "port"


def _cpu_worker(a):
    size, nu, smoother, steps, warmup, kind = a
    if kind == "reference":
        return cpu_reference_run_ref(size, nu, smoother, steps, warmup)
    r = cpu_reference_run(size, nu, smoother, steps, warmup)
    r["kind"] = "port"
    return r


def cpu_reference_all_cores(size: int, nu: int, smoother: int, steps: int, warmup: int, procs: int, kind: str = "port"):
    """The reference solver is single-threaded by construction (GenerateData, not ThreadedGenerateData; its Gauss-Seidel
    sweep is lexicographic), so "all the host cores" means one independent solve per core: `procs` processes each run the
    same bounded sample concurrently and the aggregate voxel rate is reported."""
    import multiprocessing as mp
    if procs <= 1:
        r = _cpu_worker((size, nu, smoother, steps, warmup, kind))
        r["procs"] = 1
        return r
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        rs = pool.map(_cpu_worker, [(size, nu, smoother, steps, warmup, kind)] * procs)
    wall = time.perf_counter() - t0
    slowest = max(r["s_per_cycle"] for r in rs)
    n = size ** 3
    return dict(mvox_s=procs * n / slowest / 1e6, s_per_cycle=slowest, setup_s=max(r["setup_s"] for r in rs), relres=rs[0]["relres"],
                voxels=n, procs=procs, wall_s=wall, single_core_mvox_s=rs[0]["mvox_s"], kind=kind)


def cpu_sample_text(r, size, nu, cycles):
    what = ("the reference's own code (unmodified /root/reference/include headers compiled against the stand-in ITK of oracle/shim, "
            "oracle/_ref/libmadref.so): GenerateData() with tolerance 0, set-up subtracted" if r["kind"] == "reference" else
            "the oracle port (C restatement, bit-identical to the reference headers) in faithful mode")
    return (f"{r['procs']} concurrent single-threaded solves (the reference has no threading), each a {size}^3 volume of the same "
            f"phantom/tensor, {cycles} V({nu},{nu}) cycles incl. the stop-test residual, lexicographic GS, double: {what}; "
            f"setup {r['setup_s']:.1f}s excluded; one core alone: {r.get('single_core_mvox_s', r['mvox_s']):.3f} Mvoxel/s")


def host_procs(limit=16):
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    return max(1, min(n, limit))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nu, sm = args.nu, (0 if args.smoother == "gs" else 1)
    procs = host_procs()
    kind = cpu_impl_kind(args.cpu_impl)
    size = args.cpu_size or (96 if kind == "reference" else 128)
    r = cpu_reference_all_cores(size, nu, sm, args.steps, args.warmup, procs, kind)
    line = {
        "impl": "reference",
        "metric": "3D VED V-cycle Mvoxels/s at 512^3", "value": r["mvox_s"], "unit": "Mvoxel/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["s_per_cycle"] * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(args, f"host CPU: {r['procs']} independent single-threaded solves"), size=[size] * 3, full_size=[args.size] * 3,
                       sample=f"bounded sample of the {args.size}^3 workload: {r['procs']} concurrent {size}^3 volumes of the same phantom and tensor "
                              f"(the reference's ~1 KB/voxel StencilImage cannot hold {args.size}^3)"),
        "cpu_baseline": {"value": r["mvox_s"], "unit": "Mvoxel/s", "cores": r["procs"], "kind": kind,
                         "sample": cpu_sample_text(r, size, nu, args.steps + (args.warmup if kind == "reference" else 0))},
        "e2e": {"value": r["mvox_s"], "unit": "Mvoxel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, parallelism):
    return {"workload": f"3-D VED diffusion step, synthetic {args.size}^3 float32 vessel phantom, spacing (.3125,.3125,.5), "
                        f"analytic VED-form tensor (eps .01, omega 1.5), {args.smoother.upper()} smoother, V({args.nu},{args.nu}), dt 0.1",
            "size": [args.size] * 3, "smoother": args.smoother, "iterations_per_grid": args.nu, "time_step": 0.1,
            "parallelism": parallelism, "l2": "working set per sweep (>= 4.8 GB at 512^3) exceeds the 126 MB L2; no flush needed"}


def run_ved_filter(args, img_np, device):
    """itkVEDTest_GS.cxx's filter call on the bench volume: five-scale Hessian + vesselness + tensor on the device (include/madved.h),
    then DiffusionStep (4 time steps to 1e-10); host image in, host image out, tensor never leaves HBM."""
    import numpy as np

    import multigridanisotropicdiffusion_b200 as M
    from multigridanisotropicdiffusion_b200 import phantom
    times, st, vst = [], None, None
    for rep in range(4):  # first repetition is warm-up; the median of the other three is reported (the call creates and frees ~15 GB of device memory)
        f = M.VEDMultigridImageFilter(args.smoother, device)
        f.SetInput(img_np, phantom.VED_SPACING)
        f.SetOmega(1.5)
        f.SetDiffusionIterationsPerGrid(args.nu)
        f.SetDiffusionIterations(4)
        f.SetTolerance(1e-10)
        t0 = time.perf_counter()
        f.Update()
        times.append(time.perf_counter() - t0)
        st, vst = f.stats, f.ved_stats
    n = int(np.prod(img_np.shape))
    timed = sorted(times[1:])
    return {"s_per_call": timed[len(timed) // 2], "s_per_call_all_reps": [round(t, 4) for t in times[1:]], "voxels": n, "front_end_ms": {"hessian": vst["hessian_ms"], "vesselness": vst["vesselness_ms"]},
            "front_end_Mvoxel_scale_per_s": n * vst["scales"] / ((vst["hessian_ms"] + vst["vesselness_ms"]) * 1e-3) / 1e6,
            "diffusion_ms": vst["diffusion_ms"], "h2d_ms": vst["h2d_ms"], "d2h_ms": vst["d2h_ms"], "cycles_per_step": st["cycles_per_step"],
            "call": "VEDMultigridImageFilter.Update(): 5 scales, Iterations 1, DiffusionIterations 4, tol 1e-10 (context creation included)",
            "_launches": int(vst["kernel_launches"])}


def timed_cycles(s, steps, warmup, barrier, gpu_index):
    """W untimed cycles, then exactly K cycles bracketed by barrier + synchronize; clocks sampled through both (same kernels)."""
    sampler = ClockSampler(gpu_index)
    sampler.start()
    time.sleep(0.3)  # nvidia-smi needs ~0.2 s to start
    if warmup > 0:
        s.cycles_run(warmup)
    s.set_profiling(True)
    barrier()
    t_wall = time.perf_counter()
    relres, dev_ms, st = s.cycles_run(steps)
    barrier()
    wall_ms = (time.perf_counter() - t_wall) * 1e3
    clocks = sampler.stop()
    s.set_profiling(False)
    return relres, dev_ms, st, clocks, wall_ms


def make_slab_solver(shape_global, smoother, nu, local_rank, rank, world, want_peer, **kw):
    """One z-slab context per rank (madgpu_create_slab) + the peer-memory halo when asked for and granted on every rank."""
    from multigridanisotropicdiffusion_b200 import MadSolver, phantom, slabs
    s = MadSolver(shape_global, phantom.VED_SPACING, time_step=0.1, smoother=smoother, iterations_per_grid=nu, device=local_rank,
                  rank=rank, world_size=world, nccl_id=slabs.create_unique_id(), **kw)
    peer = bool(want_peer) and slabs.enable_peer_halo(s)
    return s, peer


def slab_parity(rank, world, local_rank, dev, want_peer):
    """Hardware evidence for the z-slab path inside the driver's own N > 1 run (the multi-GPU pytest files skip on a 1-GPU box): a
    128 x 128 x 32N volume is solved as N slabs, gathered on rank 0 and compared with the single-GPU solve of the whole volume
    there -- weighted Jacobi and Gauss-Seidel, V-cycles and FMG (itkMultigridAnisotropicDiffusionImageFilter.hxx:207-246, 300-338).
    Bounds: WJ is the same iteration (same cycle counts, rel-L2 <= 1e-6), GS relaxes slab faces like tile faces (<= 1e-4)."""
    import numpy as np
    import torch.distributed as dist

    from multigridanisotropicdiffusion_b200 import MadSolver, phantom, slabs
    shape = (32 * world, 128, 128)
    img_t, D = phantom.vessel_phantom(shape, device=dev)
    img = img_t.cpu().numpy()
    T = phantom.planes_to_aos(D).cpu().numpy()
    del img_t, D
    out = {"volume": [128, 128, 32 * world], "tolerance": 1e-9, "nu": 3}
    kw = dict(tolerance=1e-9, max_cycles=40, number_of_steps=1)
    s, peer = make_slab_solver(shape, MadSolver.GS, 3, local_rank, rank, world, want_peer, **kw)
    out["halo"] = "peer stores" if peer else "nccl"
    s.set_tensor(slabs.cut(T, rank, world))
    r = None
    if rank == 0:
        r = MadSolver(shape, phantom.VED_SPACING, time_step=0.1, smoother=MadSolver.GS, iterations_per_grid=3, device=local_rank, **kw)
        r.set_tensor(T)
    ok = True
    for name, sm, cyc, bound in (("wj_v", MadSolver.WJ, MadSolver.VCYCLE, 1e-6), ("gs_v", MadSolver.GS, MadSolver.VCYCLE, 1e-4),
                                 ("wj_fmg", MadSolver.WJ, MadSolver.FMG, 1e-6), ("gs_fmg", MadSolver.GS, MadSolver.FMG, 1e-4)):
        s.set_solver(smoother=sm, cycle=cyc)
        local = s.solve(slabs.cut(img, rank, world), out_dtype=np.float64)
        cs = s.last_stats["cycles_per_step"][0]
        full = slabs.gather_volume(local)
        if rank == 0:
            r.set_solver(smoother=sm, cycle=cyc)
            ref = r.solve(img, out_dtype=np.float64)
            cr = r.last_stats["cycles_per_step"][0]
            err = float(np.linalg.norm(full - ref) / np.linalg.norm(ref))
            good = err <= bound and (cs == cr if sm == MadSolver.WJ else abs(cs - cr) <= 2) and s.last_stats["final_relres"][0] <= 1e-9
            out[name] = {"rel_l2": err, "cycles_slab": cs, "cycles_single": cr, "cycles_equal": cs == cr, "bound": bound, "ok": bool(good)}
            ok = ok and good
        dist.barrier()
    s.close()
    if r is not None:
        r.close()
    out["ok"] = bool(ok)
    return out


def extra_volume_run(shape_global, args, smoother, rank, world, local_rank, dev, want_peer, barrier, active=True):
    """K timed cycles on another volume with the same settings (own clock record): the 1024^3 z-slab run of BASELINE.json configs[4]
    and the one-GPU 512^3 run it is compared with.  world == 1 -> plain single-GPU context on this rank."""
    import numpy as np
    import torch

    from multigridanisotropicdiffusion_b200 import MadSolver, phantom, slabs
    if not active:
        return None
    nz = shape_global[0]
    if world > 1:
        z0, z1 = slabs.slab_range(nz, rank, world)
        img, D = phantom.vessel_phantom(shape_global, device=dev, z_range=(z0, z1))
        s, peer = make_slab_solver(shape_global, smoother, args.nu, local_rank, rank, world, want_peer, tolerance=0.0, max_cycles=1 << 20)
    else:
        img, D = phantom.vessel_phantom(shape_global, device=dev)
        s, peer = MadSolver(shape_global, phantom.VED_SPACING, time_step=0.1, smoother=smoother, iterations_per_grid=args.nu, tolerance=0.0,
                            max_cycles=1 << 20, device=local_rank), False
    torch.cuda.synchronize()
    s.set_tensor_device([D[c].data_ptr() for c in range(6)])
    s.cycles_begin(d_in=img.data_ptr())
    del D
    relres, dev_ms, st, clocks, _ = timed_cycles(s, args.steps, max(args.warmup, 3), barrier, local_rank)
    ms = dev_ms / args.steps
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    nvox = int(np.prod(shape_global))
    res = {"size": [shape_global[2], shape_global[1], shape_global[0]], "n_gpus": world, "value": nvox / (ms * 1e-3) / 1e6, "unit": "Mvoxel/s",
           "ms_per_step": ms, "steps": args.steps, "halo": ("peer stores" if peer else "nccl") if world > 1 else None, "clocks": clocks,
           "class_ms_per_cycle": {k: v / args.steps for k, v in st["prof_ms"].items() if v > 0},
           "cycle_frac": alg_bytes_per_cycle(args.nu) * nvox / world / (ms * 1e-3) / 1e9 / measured_peaks()[0],
           "relres_after_timed_cycles": float(relres[-1]) if len(relres) else None}
    s.close()
    del img
    torch.cuda.empty_cache()
    return res



def run_ours(args):
    import numpy as np
    import torch

    from multigridanisotropicdiffusion_b200 import MadGpuError, MadSolver, phantom

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = args.size
    shape = (n, n, n)
    nu = args.nu
    smoother = MadSolver.GS if args.smoother == "gs" else MadSolver.WJ
    slab_mode = world > 1 and not args.replicas
    peer_halo = False
    want_peer = False
    if slab_mode:
        # ONE volume cut into z-slabs, one per GPU (strong scaling); halo = peer stores from the producing kernels (CUDA IPC over
        # NVLink) with bounded in-stream waits, NCCL send/recv where the library or a rank declines
        from multigridanisotropicdiffusion_b200 import slabs
        z0, z1 = slabs.slab_range(n, rank, world)
        img, D = phantom.vessel_phantom(shape, device=dev, z_range=(z0, z1))
        torch.cuda.synchronize()
        want_peer = not args.nccl_halo
        # arrival counters are awaited by the bounded k_halo_wait (the library's default): a signal that never arrives costs a
        # time-out and an error on every rank -- caught by the trial cycles below -- instead of a hung stream
        os.environ.setdefault("MADGPU_P2P_TIMEOUT_MS", "3000")
        s, peer_halo = make_slab_solver(shape, smoother, nu, local_rank, rank, world, want_peer, tolerance=0.0, max_cycles=1 << 20)
        if peer_halo:
            # trial cycles: a time-out is reported by every rank in the same cycle (the flag travels with the norm all-reduce)
            try:
                s.set_tensor_device([D[c].data_ptr() for c in range(6)])
                s.cycles_begin(d_in=img.data_ptr())
                s.cycles_run(2)
            except MadGpuError as e:
                if rank == 0:
                    print(f"bench.py: peer-memory halo failed on {world} ranks ({e}); falling back to the NCCL halo", file=sys.stderr, flush=True)
                s.close()
                want_peer = False
                s, peer_halo = make_slab_solver(shape, smoother, nu, local_rank, rank, world, False, tolerance=0.0, max_cycles=1 << 20)
    else:
        img, D = phantom.vessel_phantom(shape, device=dev)
        torch.cuda.synchronize()
        s = MadSolver(shape, phantom.VED_SPACING, time_step=0.1, smoother=smoother, iterations_per_grid=nu, tolerance=0.0,
                      max_cycles=1 << 20, device=local_rank)
    shape = s.shape            # local slab (== the volume on one GPU / in replica mode)
    nvox = int(np.prod(shape))  # voxels this rank processes
    s.set_tensor_device([D[c].data_ptr() for c in range(6)])
    s.cycles_begin(d_in=img.data_ptr())
    # ---- device-resident timing: W warm-up cycles, then exactly K cycles ----
    relres, dev_ms, st, clocks, wall_ms = timed_cycles(s, args.steps, args.warmup, barrier, local_rank)
    ms_per_step = dev_ms / args.steps
    if world > 1:
        t = torch.tensor([ms_per_step], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step = float(t.item())
    # slabs: one volume over all ranks (strong scaling); replicas: every rank runs its own volume (weak)
    total_vox = nvox * world
    value = total_vox / (ms_per_step * 1e-3) / 1e6

    # ---- roofline of the dominant kernel: level-0 smoother sweeps ----
    peak, peak_kind = measured_peaks()
    sm_ms = st["prof_ms"]["smooth0"]
    sm_launches = st["prof_launches"]["smooth0"]
    sweeps = 2 * nu * args.steps
    achieved = ALG_BYTES_SWEEP_3D * nvox * sweeps / (sm_ms * 1e-3) / 1e9 if sm_ms > 0 else None
    tile = s.gs_tile(0) if args.smoother == "gs" else None
    tb_env = int(os.environ.get("MADGPU_GS_TB", "1") or 1) > 1 or int(os.environ.get("MADGPU_GS_TB_SINGLE", "0") or 0) > 0
    launched_kernel = ("k_fast_sweep<MODE_WJ>" if args.smoother == "wj" else
                       "k_coef_gs_tb" if tile and tb_env else  # opt-in shared-memory sweeps
                       "k_coef_gs2<4,3,PRIVATE>" if tile and tile[1] == 2 else "k_coef_gs2<4,3>" if tile and tile[1] == 8 else "k_coef_gs" if tile else "k_gs_color")
    traffic, traffic_src = None, None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))
        for key in ("smooth0", "smooth0_shared_tiles", "smooth0_" + args.smoother):
            ent = prof.get(key, {})
            # only a capture of the kernel that was actually launched, at this size, counts
            if ent.get("size") == n and ent.get("smoother") == args.smoother and world == 1 and ent.get("kernel", "") == launched_kernel:
                traffic, traffic_src = ent.get("dram_bytes_per_launch"), ent.get("source")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": f"level-0 smoother sweep ({launched_kernel})", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "peak_kind": peak_kind, "traffic": traffic, "traffic_source": traffic_src,
                "alg_bytes_per_launch": ALG_BYTES_SWEEP_3D * nvox * sweeps / max(sm_launches, 1),
                "launches": sm_launches, "ms_per_launch": sm_ms / max(sm_launches, 1),
                "dram_bytes_per_voxel": (traffic / nvox) if traffic else None,
                "frac_on_dram_bytes": (traffic / (sm_ms / max(sm_launches, 1) * 1e-3) / 1e9 / peak) if traffic and sm_ms > 0 else None,
                "note": "achieved counts the canonical 36 B per voxel and sweep; the Gauss-Seidel sweep reads pre-evaluated fp16 operator "
                        "rows (20 B) instead of the six fp32 tensor planes (24 B), so its DRAM traffic is 32 B per voxel and frac can pass 1 "
                        "(frac_on_dram_bytes = the ncu-measured bytes over the same event time)",
                "share_of_step": sm_ms / dev_ms if dev_ms > 0 else None,
                "class_ms_per_cycle": {k: v / args.steps for k, v in st["prof_ms"].items() if v > 0},
                "class_launches_per_cycle": {k: v / args.steps for k, v in st["prof_launches"].items() if v > 0},
                "cycle_alg_bytes_per_voxel": alg_bytes_per_cycle(nu),
                "cycle_frac": alg_bytes_per_cycle(nu) * nvox / (ms_per_step * 1e-3) / 1e9 / peak}
    launches_timed = st["kernel_launches"]

    # ---- e2e: the filter call with host (pinned) buffers ----
    e2e = None
    img_np = None
    if args.e2e_reps > 0:
        T_h = torch.empty(shape + (6,), dtype=torch.float32, pin_memory=True)
        T_h.copy_(phantom.planes_to_aos(D))
        img_h = torch.empty(shape, dtype=torch.float32, pin_memory=True)
        img_h.copy_(img)
        out_h = torch.empty(shape, dtype=torch.float32, pin_memory=True)
        del D
        torch.cuda.synchronize()
        s.set_solver(tolerance=1e-10, max_cycles=100, number_of_steps=4)  # DiffusionStep of itkVEDTest_GS.cxx
        T_np, img_np, out_np = T_h.numpy(), img_h.numpy(), out_h.numpy()
        cycles = 0
        times = []
        for rep in range(args.e2e_reps + 1):  # first repetition is warm-up
            barrier()
            t0 = time.perf_counter()
            s.set_tensor(T_np)
            s.solve(img_np, out=out_np)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if rep > 0:
                times.append(dt)
                cycles = s.last_stats["total_cycles"]
                launches_timed += s.last_stats["kernel_launches"]
        times.sort()
        dt = times[len(times) // 2]  # median repetition (host-side copy times vary on shared hosts); all are reported
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": total_vox * cycles / dt / 1e6, "unit": "Mvoxel/s",
               "h2d_bytes_per_step": int(T_h.numel() * 4 + img_h.numel() * 4) * world, "d2h_bytes_per_step": int(out_h.numel() * 4) * world,
               "call": "SetDiffusionTensor(host fp32 AoS) + solve(host fp32 image): 4 time steps to relres 1e-10",
               "cycles_per_call": cycles, "s_per_call": dt, "s_per_call_all_reps": [round(t, 4) for t in times], "cycles_per_step": s.last_stats["cycles_per_step"],
               "final_relres": max(s.last_stats["final_relres"]), "setup_ms": s.last_stats["setup_ms"],
               "h2d_ms": s.last_stats["h2d_ms"], "d2h_ms": s.last_stats["d2h_ms"], "graph_launches": s.last_stats.get("graph_launches")}
        del T_h, out_h, T_np, out_np
    elif args.ved and world == 1:
        img_np = img.cpu().numpy()
    s.close()
    del img
    D = None
    torch.cuda.empty_cache()

    # ---- the whole VED filter (tensor front-end on the device + DiffusionStep) through the filter call, host buffers ----
    ved = None
    if args.ved and world == 1:
        try:
            ved = run_ved_filter(args, img_np, local_rank)
            launches_timed += ved.pop("_launches", 0)
        except Exception as e:  # noqa: BLE001 -- an extra, never the headline
            ved = {"error": f"{type(e).__name__}: {e}"}

    # ---- N > 1: parity of the z-slab path with the single-GPU solve, on this run's GPUs ----
    parity = None
    if slab_mode and not args.no_slab_parity:
        try:
            parity = slab_parity(rank, world, local_rank, dev, want_peer and peer_halo)
        except Exception as e:  # noqa: BLE001
            parity = {"ok": False, "error": f"{type(e).__name__}: {e}"}

    # ---- N = 8 (or --weak): BASELINE.json configs[4], 1024^3 on 8 GPUs, next to 512^3 on ONE GPU of the same box in the same run ----
    extra = None
    if slab_mode and (world == 8 or args.weak) and args.size == 512:
        try:
            wshape = (1024, 1024, 1024) if world == 8 else (512 * world, 512, 512)
            w = extra_volume_run(wshape, args, smoother, rank, world, local_rank, dev, want_peer and peer_halo, barrier)
            barrier()
            one = extra_volume_run((512, 512, 512), args, smoother, 0, 1, local_rank, dev, False, lambda: torch.cuda.synchronize(), active=rank == 0)
            barrier()
            big = None
            if world == 8 and not args.no_n1_1024:
                # the same 1024^3 volume on ONE GPU of this box (it fits: ~80 GB): the strong-scaling denominator of configs[4]
                try:
                    big = extra_volume_run(wshape, args, smoother, 0, 1, local_rank, dev, False, lambda: torch.cuda.synchronize(), active=rank == 0)
                except Exception as e:  # noqa: BLE001 -- e.g. out of memory on a box whose GPU 0 is shared
                    big = {"error": f"{type(e).__name__}: {e}"}
                torch.cuda.empty_cache()
                barrier()
            if rank == 0:
                w["vs_n1"] = w["value"] / one["value"]
                w["n1_same_run"] = one
                if big is not None:
                    w["n1_same_volume"] = big
                    if big.get("value"):
                        w["vs_n1_same_volume"] = w["value"] / big["value"]
                w["target"] = "BASELINE.json north_star: >= 6x one GPU at 512^3 for 1024^3 on 8 GPUs (same work per GPU)"
            extra = {"weak_1024" if world == 8 else f"weak_512x512x{512 * world}": w}
        except Exception as e:  # noqa: BLE001
            extra = {"weak_1024": {"error": f"{type(e).__name__}: {e}"}}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            kind = cpu_impl_kind(args.cpu_impl)
            size = args.cpu_size or (96 if kind == "reference" else 128)
            r = cpu_reference_all_cores(size, nu, 0 if args.smoother == "gs" else 1, args.cpu_steps, 0, host_procs(), kind)
            cpu = {"value": r["mvox_s"], "unit": "Mvoxel/s", "cores": r["procs"], "kind": kind, "sample": cpu_sample_text(r, size, nu, args.cpu_steps)}
        except Exception as e:  # noqa: BLE001 -- the GPU numbers above must still be printed
            cpu = {"value": None, "unit": "Mvoxel/s", "cores": 0, "kind": "port", "sample": f"CPU baseline failed: {type(e).__name__}: {e}"}

    if rank == 0:
        line = {
            "metric": "3D VED V-cycle Mvoxels/s at 512^3", "value": value, "unit": "Mvoxel/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if (slab_mode or world == 1) else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, f"z-slabs: {world} x {shape[0]} planes, halo = " + ("NVLink peer stores from the producing kernels + stream memory ops"
                                      if peer_halo else "NCCL send/recv per sweep") + ", coarse levels agglomerated on rank 0"
                                      if slab_mode else "independent replicas (one volume per GPU)" if world > 1 else "single GPU"),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "ved_filter": ved, "slab_parity": parity, "extra": extra,
            "gpu_launches": int(launches_timed), "clocks": clocks,
            "relres_after_timed_cycles": float(relres[-1]) if len(relres) else None, "wall_ms_timed_region": wall_ms,
            "timed_iterate": "the timed cycles continue from the warm-up cycles' iterate (tolerance 0: the loop never stops early); every kernel "
                             "of a cycle does the same work whatever the residual, so the time per cycle does not depend on it",
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--smoother", default="gs", choices=["gs", "wj"])
    ap.add_argument("--nu", type=int, default=3)
    ap.add_argument("--cpu-size", type=int, default=0, help="edge of the CPU sample volume (default: 96 for the compiled reference, 128 for the port)")
    ap.add_argument("--cpu-impl", default="auto", choices=["auto", "reference", "port"],
                    help="CPU legs: oracle/_ref (the compiled reference headers) when present, else the oracle port")
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--e2e-reps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--nccl-halo", action="store_true", help="N > 1: keep the NCCL send/recv halo exchange instead of peer stores")
    ap.add_argument("--peer-halo", action="store_true", help="(default since round 2: the peer-memory halo is used at every N unless --nccl-halo)")
    ap.add_argument("--ved", dest="ved", action="store_true", default=True, help="N = 1: also time the whole VED filter (tensor front-end + diffusion) through the filter call (default)")
    ap.add_argument("--no-ved", dest="ved", action="store_false")
    ap.add_argument("--no-slab-parity", action="store_true", help="N > 1: skip the slab-vs-single-GPU parity solves after the timed region")
    ap.add_argument("--weak", action="store_true", help="N > 1: also run the weak-scaling volume (512 x 512 x 512N; 1024^3 at N = 8 is always run)")
    ap.add_argument("--no-n1-1024", action="store_true", help="N = 8: skip the one-GPU run of the 1024^3 volume (the strong-scaling denominator)")
    ap.add_argument("--replicas", action="store_true", help="N > 1: independent volumes per GPU instead of z-slabs of one volume")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
