#!/usr/bin/env python
"""z-slab decomposition of libmadgpu on the CPU: every rank is a THREAD of this process driving its own context of the host build
(tests/mad_host/madgpu_host.cpp), NCCL is tests/mad_host/fake_nccl.cpp, CUDA IPC handles are pointers, stream memory operations
are release stores / blocking waits (tests/fake_cuda/cuda_runtime.h).  The distributed solve is compared with the single-context
solve of the whole volume by the same build.  A halo protocol error shows up as a wrong result, a dead-lock as a time-out abort.

    python tests/mad_host/slab_emulation.py --world 8 --peer 1 --agglomerate-voxels 1000 --shape 128,32,32 --smoother gs --cycle v

TEST INFRASTRUCTURE; run in its own process by tests/test_cpu_mad_host_slabs.py."""
import argparse
import ctypes as C
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "mad_host"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, required=True)
    ap.add_argument("--peer", type=int, default=0, help="1: peer-store halo (IPC import + handshake), 0: NCCL send/recv")
    ap.add_argument("--agglomerate-voxels", type=int, default=64 ** 3)
    ap.add_argument("--shape", default="")
    ap.add_argument("--smoother", default="gs", choices=["gs", "wj"])
    ap.add_argument("--cycle", default="v", choices=["v", "fmg"])
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--nu", type=int, default=2)
    ap.add_argument("--wait", default="memop", choices=["memop", "kernel"], help="peer halo: cuStreamWaitValue32 or the bounded k_halo_wait")
    ap.add_argument("--tb", type=int, default=1, help="MADGPU_GS_TB: Gauss-Seidel sweeps fused per pass.  1 (no temporal blocking) keeps the slab solve bit-identical "
                    "to the single-context solve; with > 1 the tile grids of a slab and of the whole volume differ, so the images agree to the solver tolerance")
    ap.add_argument("--drop", default="", help="rank:seq -- from that sequence number on the rank's arrival signals are not sent (needs --wait kernel); every rank must report the time-out")
    a = ap.parse_args()
    import hostlib
    os.environ["MADGPU_NCCL_LIB"] = hostlib.FAKE_NCCL
    os.environ["MADGPU_FAST_MIN_NX"] = "8"  # streaming kernels on these narrow volumes (the peer halo needs them)
    os.environ["MADGPU_AGGLOMERATE_VOXELS"] = str(a.agglomerate_voxels)
    os.environ["MADGPU_P2P_WAIT"] = a.wait
    os.environ["MADGPU_P2P_TIMEOUT_MS"] = "400"
    os.environ["MADGPU_GS_TB"] = str(a.tb)
    if a.drop:
        os.environ["MADGPU_P2P_TEST_DROP_SIGNAL"] = a.drop
    L = hostlib.load()
    hostlib.bind(L)
    from multigridanisotropicdiffusion_b200 import MadSolver, slabs
    from util import random_image, random_spd_tensor, rel_l2
    world = a.world
    shape = tuple(int(x) for x in a.shape.split(",")) if a.shape else (8 * world, 24, 24)
    sp = (0.3125, 0.3125, 0.5)
    T = random_spd_tensor(shape, seed=2)
    img = random_image(shape, seed=5)
    raw = C.create_string_buffer(128)
    assert L.madgpu_nccl_unique_id(raw) == 0  # also loads the NCCL binding before the rank threads start
    kw = dict(time_step=0.1, smoother=MadSolver.GS if a.smoother == "gs" else MadSolver.WJ, iterations_per_grid=a.nu, tolerance=1e-9,
              max_cycles=40, number_of_steps=a.steps, cycle=MadSolver.FMG if a.cycle == "fmg" else MadSolver.VCYCLE)
    print("plan", slabs.plan(shape[::-1], world), flush=True)
    outs, stats, errs, blobs = [None] * world, [None] * world, [], [None] * world
    bar = threading.Barrier(world)

    def run(r):
        try:
            s = MadSolver(shape, sp, rank=r, world_size=world, nccl_id=raw.raw, **kw)
            if a.peer:
                blobs[r] = s.ipc_export()
                bar.wait()
                s.ipc_import(blobs[r - 1] if r > 0 else None, blobs[r + 1] if r < world - 1 else None)
                bar.wait()
            s.set_tensor(slabs.cut(T, r, world))
            outs[r] = s.solve(slabs.cut(img, r, world), out_dtype=np.float64)
            stats[r] = s.last_stats
            bar.wait()
            s.close()
        except Exception as e:  # noqa: BLE001
            errs.append((r, repr(e)))
            if not a.drop:
                bar.abort()

    single = {}

    def run_single():  # the whole volume in one context, concurrently with the ranks
        s = MadSolver(shape, sp, **kw)
        s.set_tensor(T)
        single["out"] = s.solve(img, out_dtype=np.float64)
        single["stats"] = s.last_stats
        s.close()

    t0 = time.time()
    ts = threading.Thread(target=run_single)
    ts.start()
    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join() for t in th]
    print(f"slabs done after {time.time() - t0:.1f}s", flush=True)
    ts.join()
    if a.drop:  # a lost signal: nobody hangs, every rank returns the same error
        ok = len(errs) == world and all("timed out" in e for _, e in errs)
        print("ERRORS", sorted(errs)[:2], flush=True)
        print("SLAB_EMULATION_TIMEOUT_REPORTED" if ok else "SLAB_EMULATION_FAILED", flush=True)
        return 0 if ok else 1
    if errs:
        print("ERRORS", errs, flush=True)
        return 1
    full = np.concatenate(outs, axis=0)
    print(f"slabs: cycles {stats[0]['cycles_per_step']} relres {stats[0]['final_relres']}", flush=True)
    ref, rst = single["out"], single["stats"]
    err = rel_l2(full, ref)
    print(f"single: cycles {rst['cycles_per_step']}  rel-L2 slabs vs single {err:.3e}", flush=True)
    if a.tb > 1 and a.smoother == "gs":  # different tile grids: same fixed point, not the same iterates
        ok = err < 2e-8 and all(abs(x - y) <= 1 for x, y in zip(stats[0]["cycles_per_step"], rst["cycles_per_step"])) and max(stats[0]["final_relres"]) <= 1e-9
    else:
        ok = err < 1e-10 and stats[0]["cycles_per_step"] == rst["cycles_per_step"] and max(stats[0]["final_relres"]) <= 1e-9
    ok = ok and all(st["cycles_per_step"] == stats[0]["cycles_per_step"] for st in stats) and L.mad_host_live_allocs() == 0
    print("SLAB_EMULATION_OK" if ok else "SLAB_EMULATION_FAILED", flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
