"""Host-side plumbing of the z-slab decomposition (one process per GPU, torch.distributed for the plumbing).

The solve itself is collective inside libmadgpu.so (NCCL send/recv of halo planes on the solver's stream, see
include/madgpu.h); this module only (1) mirrors the library's slab plan so that callers can cut inputs before creating a
context, (2) moves the NCCL unique id and small host values between ranks with torch.distributed (any backend: the
CPU tests use gloo), (3) cuts / reassembles volumes.  The reference has no counterpart: it is single-process.
"""
from __future__ import annotations

import numpy as np


def level_schedule(size_xyz):
    """GridsHierarchy level schedule (mad/itkGridsHierarchy.hxx:36-106) -> list of sizes (x, y, z)."""
    g = list(size_xyz)
    out = [tuple(g)]
    while True:
        g = [n // 2 if n % 2 == 0 else (n - 1) // 2 + 1 for n in g]
        if any(n < 6 for n in g):
            break
        out.append(tuple(g))
    return out


def plan(size_xyz, world_size):
    """Mirror of plan_slabs() in csrc/madgpu.cu.  Returns dict(agglomeration_level, planes_per_rank=[...per distributed
    level and the agglomeration level]) or raises ValueError with the library's reason."""
    sizes = level_schedule(size_xyz)
    if len(size_xyz) != 3:
        raise ValueError("z-slab decomposition needs a 3-D volume")
    if len(sizes) < 2:
        raise ValueError("the volume has a single level, nothing to distribute")
    if sizes[0][2] % world_size:
        raise ValueError("size[2] must be divisible by world_size")
    import os
    lnz = sizes[0][2] // world_size
    planes = [lnz]
    la = 0
    small = int(os.environ.get("MADGPU_AGGLOMERATE_VOXELS", 64 ** 3))  # the library's tuning hook
    for l in range(len(sizes) - 1):
        ok = sizes[l][2] % 2 == 0 and lnz % 2 == 0 and lnz >= 4
        vox = sizes[l][0] * sizes[l][1] * sizes[l][2]
        if not ok or (l > 0 and vox <= small):
            break
        lnz //= 2
        planes.append(lnz)
        la = l + 1
    if la < 1:
        raise ValueError("planes per rank must be even and >= 4 on the finest level")
    return dict(agglomeration_level=la, planes_per_rank=planes, levels=sizes)


def slab_range(nz, rank, world_size):
    """Planes [z0, z1) of a level with nz planes owned by `rank`."""
    if nz % world_size:
        raise ValueError("nz must be divisible by world_size")
    c = nz // world_size
    return rank * c, (rank + 1) * c


def cut(volume, rank, world_size):
    """Local slab (a view) of a (nz, ny, nx[, ncomp]) array."""
    z0, z1 = slab_range(volume.shape[0], rank, world_size)
    return volume[z0:z1]


def create_unique_id(group=None):
    """128-byte NCCL unique id, made on rank 0 and broadcast with torch.distributed (works on any backend)."""
    import ctypes as C

    import torch
    import torch.distributed as dist
    from . import _lib
    buf = torch.zeros(128, dtype=torch.uint8)
    if dist.get_rank(group) == 0:
        raw = C.create_string_buffer(128)
        rc = _lib.load().madgpu_nccl_unique_id(raw)
        if rc != 0:
            raise RuntimeError("madgpu_nccl_unique_id failed: " + (_lib.load().madgpu_last_error(None) or b"").decode())
        buf = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone()
    return broadcast_bytes(buf, group)


def broadcast_bytes(buf, group=None):
    """Broadcast a uint8 CPU tensor from rank 0; NCCL groups stage through the current device."""
    import torch
    import torch.distributed as dist
    if dist.get_backend(group) == "nccl":
        d = buf.cuda()
        dist.broadcast(d, src=0, group=group)
        return bytes(d.cpu().numpy().tobytes())
    dist.broadcast(buf, src=0, group=group)
    return bytes(buf.numpy().tobytes())


def gather_volume(local, group=None):
    """Reassemble the slabs on every rank (testing / small volumes): numpy (lnz, ny, nx) -> (nz, ny, nx)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    t = torch.from_numpy(np.ascontiguousarray(local))
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t, group=group)
    return torch.cat(parts, dim=0).cpu().numpy()


def max_over_ranks(value, group=None):
    """Timing reduction of the benchmark contract: device time = max over ranks."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def enable_peer_halo(solver, group=None):
    """Switch a z-slab MadSolver to the peer-memory halo: all-gather the ranks' CUDA IPC blobs (torch.distributed, any
    backend) and hand every rank its two neighbours'.  Collective.  Returns False (and leaves the NCCL exchange in use)
    when the library declines, e.g. a slab level narrower than the streaming kernels need."""
    import torch
    import torch.distributed as dist
    from .solver import MadGpuError
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    blob = solver.ipc_export()
    t = torch.frombuffer(bytearray(blob), dtype=torch.uint8).clone()
    nccl = dist.get_backend(group) == "nccl"
    if nccl:
        t = t.cuda()
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t, group=group)
    blobs = [bytes(p.cpu().numpy().tobytes()) for p in parts]
    ok = 1
    try:
        solver.ipc_import(blobs[rank - 1] if rank > 0 else None, blobs[rank + 1] if rank < world - 1 else None)
    except MadGpuError:
        ok = 0
    f = torch.tensor([ok], dtype=torch.int32)
    if nccl:
        f = f.cuda()
    dist.all_reduce(f, op=dist.ReduceOp.MIN, group=group)
    if int(f.item()) == 0:
        solver.ipc_disable()  # all or none: a rank that imported must not signal neighbours that did not
        return False
    return True
