"""Host-side logic of the z-slab path on the CPU: the slab plan (mirror of plan_slabs() in csrc/madgpu.cu), volume
cutting / reassembly and the small collectives (unique-id broadcast, max-over-ranks timing) over torch.distributed
with the gloo backend and world_size 2.  The GPU side (NCCL halo exchange inside libmadgpu.so) is checked by
tests/test_gpu_multi.py on a multi-GPU box."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_plan():
    from multigridanisotropicdiffusion_b200 import slabs
    p = slabs.plan((512, 512, 512), 8)
    assert p["agglomeration_level"] == 3 and p["planes_per_rank"] == [64, 32, 16, 8]
    assert p["levels"][3] == (64, 64, 64)  # levels of 64^3 or fewer voxels go to rank 0
    p = slabs.plan((1024, 1024, 1024), 8)
    assert p["agglomeration_level"] == 4 and p["planes_per_rank"] == [128, 64, 32, 16, 8]
    assert slabs.plan((512, 512, 512), 2)["planes_per_rank"] == [256, 128, 64, 32]
    # odd plane counts cannot be cut with a one-plane halo
    assert slabs.plan((160, 96, 136), 2)["agglomeration_level"] == 1  # 68 planes per rank, 34 on level 1: level 1 (80x48x68) is small
    with pytest.raises(ValueError):
        slabs.plan((69, 77, 69), 2)
    with pytest.raises(ValueError):
        slabs.plan((64, 64, 36), 8)   # 36 planes do not divide by 8
    with pytest.raises(ValueError):
        slabs.plan((40, 40, 10), 2)   # a single level: nothing to distribute
    with pytest.raises(ValueError):
        slabs.plan((64, 64, 12), 4)   # 3 planes per rank
    assert slabs.plan((64, 64, 12), 2)["planes_per_rank"] == [6, 3]
    with pytest.raises(ValueError):
        slabs.plan((512, 512), 2)


def test_level_schedule_matches_the_oracle():
    from multigridanisotropicdiffusion_b200 import slabs
    from oracle import oracle as O
    for size in ((512, 512, 512), (134, 140, 119), (69, 77, 69), (256, 128, 64)):
        assert slabs.level_schedule(size) == [s for s, _ in O.level_schedule(size)]


def test_cut_and_ranges():
    from multigridanisotropicdiffusion_b200 import slabs
    v = np.arange(8 * 3 * 2).reshape(8, 3, 2)
    assert slabs.slab_range(8, 1, 4) == (2, 4)
    parts = [slabs.cut(v, r, 4) for r in range(4)]
    assert all(p.shape == (2, 3, 2) for p in parts)
    np.testing.assert_array_equal(np.concatenate(parts, 0), v)
    with pytest.raises(ValueError):
        slabs.slab_range(9, 0, 2)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from multigridanisotropicdiffusion_b200 import slabs
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        # the unique id travels as 128 bytes from rank 0 (here a stand-in: no NCCL on the CPU)
        buf = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
        got = slabs.broadcast_bytes(buf)
        assert got == bytes(range(128))
        # slabs cut from the same volume reassemble to it
        vol = np.random.default_rng(0).standard_normal((8, 5, 6))
        full = slabs.gather_volume(slabs.cut(vol, rank, world).copy())
        np.testing.assert_array_equal(full, vol)
        # benchmark contract: device time = max over ranks
        assert slabs.max_over_ranks(10.0 + rank) == 10.0 + world - 1
        q.put((rank, "ok"))
    except Exception as e:  # noqa
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_gloo_world_size_2_plumbing():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: "ok", 1: "ok"}, res


def test_slab_context_needs_nccl_id_and_gpu():
    import multigridanisotropicdiffusion_b200 as M
    with pytest.raises(M.MadGpuError):
        M.MadSolver((64, 64, 64), world_size=2, rank=0)  # no unique id
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(M.MadGpuError):
            M.MadSolver((64, 64, 64), world_size=2, rank=1, nccl_id=bytes(128))
