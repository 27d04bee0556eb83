"""Run under torchrun with >= 2 GPUs (tests/test_gpu_multi.py does):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py

Every rank owns one z-slab of the volume (madgpu_create_slab: NCCL halo exchange inside libmadgpu.so, coarse levels
agglomerated on rank 0).  The distributed solve is compared with the single-GPU solve of the whole volume on rank 0:
weighted Jacobi is the same iteration (same cycle counts, per-cycle residuals, image); Gauss-Seidel relaxes slab faces
Jacobi-style like tile faces, so it is compared on the converged image."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from multigridanisotropicdiffusion_b200 import MadSolver, phantom, slabs  # noqa: E402
from util import rel_l2  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ok = True
    for shape in ((128, 96, 160), (256, 128, 128)):
        plan = slabs.plan(shape[::-1], world)
        img_t, D = phantom.vessel_phantom(shape)  # whole volume on the CPU, identical on every rank
        img = img_t.numpy()
        T = phantom.planes_to_aos(D).numpy()
        for smoother, name, peer in ((MadSolver.WJ, "wj", False), (MadSolver.GS, "gs", False), (MadSolver.WJ, "wj", True), (MadSolver.GS, "gs", True)):
            uid = slabs.create_unique_id()
            s = MadSolver(shape, phantom.VED_SPACING, time_step=0.1, smoother=smoother, iterations_per_grid=3, tolerance=1e-9,
                          max_cycles=40, number_of_steps=2, device=local_rank, rank=rank, world_size=world, nccl_id=uid)
            assert s.shape == (shape[0] // world,) + shape[1:], s.shape
            if peer:  # halo through peer stores + stream memory operations instead of NCCL send/recv
                got = slabs.enable_peer_halo(s)
                assert got, "peer-memory halo was declined"
            name = name + ("+peer" if peer else "+nccl")
            s.set_tensor(slabs.cut(T, rank, world))
            out_local = s.solve(slabs.cut(img, rank, world), out_dtype=np.float64)
            st = s.last_stats
            hist = s.relres_history().reshape(2, 40)
            full = slabs.gather_volume(out_local)
            s.close()
            if rank == 0:
                r = MadSolver(shape, phantom.VED_SPACING, time_step=0.1, smoother=smoother, iterations_per_grid=3, tolerance=1e-9,
                              max_cycles=40, number_of_steps=2, device=local_rank)
                r.set_tensor(T)
                ref = r.solve(img, out_dtype=np.float64)
                rst = r.last_stats
                rhist = r.relres_history().reshape(2, 40)
                r.close()
                err = rel_l2(full, ref)
                print(f"[{name} {shape} world {world} agglomeration level {plan['agglomeration_level']}] cycles slab {st['cycles_per_step'][:2]} "
                      f"single {rst['cycles_per_step'][:2]}  rel-L2 {err:.3e}  relres {st['final_relres'][:2]}", flush=True)
                if smoother == MadSolver.WJ:
                    good = st["cycles_per_step"][:2] == rst["cycles_per_step"][:2] and err < 1e-9
                    n = st["cycles_per_step"][0]
                    good = good and np.allclose(hist[0][:n], rhist[0][:n], rtol=1e-6, atol=1e-14)
                else:
                    good = err < 1e-6 and all(abs(a - b) <= 2 for a, b in zip(st["cycles_per_step"][:2], rst["cycles_per_step"][:2]))
                good = good and max(st["final_relres"][:2]) <= 1e-9
                ok = ok and good
                if not good:
                    print("   MISMATCH", flush=True)
            dist.barrier()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    if rank == 0:
        print("MULTI_GPU_OK" if ok else "MULTI_GPU_FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
