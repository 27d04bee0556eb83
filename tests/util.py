"""Shared helpers of the test-suite: seeded inputs handed identically to the oracle and the GPU path."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def random_spd_tensor(shape, seed=0, smooth=True, lo=0.05, hi=2.0):
    """Random, smoothly varying SPD tensor field, AoS float32 (so oracle and GPU see identical values)."""
    rng = np.random.default_rng(seed)
    dim = len(shape)
    ncomp = 3 if dim == 2 else 6
    # random rotation field from smooth angles + random positive eigenvalues
    def field():
        f = rng.standard_normal(shape)
        if smooth:
            for ax in range(dim):
                f = (np.roll(f, 1, ax) + 2 * f + np.roll(f, -1, ax)) / 4
                f = (np.roll(f, 1, ax) + 2 * f + np.roll(f, -1, ax)) / 4
        return f
    if dim == 2:
        th = 3.0 * field()
        l1 = lo + (hi - lo) * (0.5 + 0.5 * np.tanh(2 * field()))
        l2 = lo + (hi - lo) * (0.5 + 0.5 * np.tanh(2 * field()))
        c, s = np.cos(th), np.sin(th)
        T = np.empty(shape + (3,), dtype=np.float32)
        T[..., 0] = l1 * c * c + l2 * s * s
        T[..., 1] = (l1 - l2) * c * s
        T[..., 2] = l1 * s * s + l2 * c * c
        return T
    A = np.stack([field() for _ in range(9)], axis=-1).reshape(shape + (3, 3))
    Q, _ = np.linalg.qr(A)
    lam = np.stack([lo + (hi - lo) * (0.5 + 0.5 * np.tanh(2 * field())) for _ in range(3)], axis=-1)
    M = np.einsum("...ik,...k,...jk->...ij", Q, lam, Q)
    T = np.empty(shape + (6,), dtype=np.float32)
    T[..., 0] = M[..., 0, 0]; T[..., 1] = M[..., 0, 1]; T[..., 2] = M[..., 0, 2]
    T[..., 3] = M[..., 1, 1]; T[..., 4] = M[..., 1, 2]; T[..., 5] = M[..., 2, 2]
    assert T.shape[-1] == ncomp
    return T


def random_image(shape, seed=0, scale=100.0):
    rng = np.random.default_rng(seed + 1000)
    img = rng.standard_normal(shape)
    for ax in range(len(shape)):
        img = (np.roll(img, 1, ax) + 2 * img + np.roll(img, -1, ax)) / 4
    return (scale * (1.0 + img)).astype(np.float32)


def load_lena():
    """512x512 uint8 decode of the reference's test/test_data/lena.jpg (fixture made by tests/golden/make_fixtures.py)."""
    return np.load(os.path.join(GOLDEN, "lena_512_u8.npz"))["image"]


def load_ved_test():
    """69x77x69 int16 volume of the reference's test/test_data/ved_test.mhd/.zraw, spacing (.3125,.3125,.5)."""
    z = np.load(os.path.join(GOLDEN, "ved_test_i16.npz"))
    return z["image"], tuple(float(s) for s in z["spacing"])


def gs_leg_model(S, u, f, plan):
    """CPU model of a Gauss-Seidel leg as the library plans it (MadSolver.gs_leg_plan): every pass runs `fused` sweeps of
    gs_tile_sweep on its (possibly shifted) tile grid with the values outside a tile frozen at those the pass started from."""
    for p in plan:
        start = u
        for _ in range(p["fused"]):
            u = gs_tile_sweep(S, u, f, p["tile"], outside=start, shift=p["shift"])
    return u


def gs_tile_sweep(S, u, f, tile, outside=None, shift=(0, 0)):
    """CPU model (numpy, explicit operator rows `S` from the oracle) of the fused GPU Gauss-Seidel sweep:
    tiles of (tx, ty, tz) voxels; inside a tile planes in z order, each plane as even rows (even x, odd x)
    then odd rows; values outside the tile are those of `outside` (default: `u`, the previous sweep; the kernel
    that fuses two sweeps in one pass reads them from the array the pass started from).  3-D only."""
    TX, TY, TZ = tile
    nz, ny, nx = u.shape
    offs = [(ox, oy, oz) for oz in (-1, 0, 1) for oy in (-1, 0, 1) for ox in (-1, 0, 1)]
    diag = S[..., 13]
    old = np.zeros((nz + 2, ny + 2, nx + 2))
    old[1:-1, 1:-1, 1:-1] = u if outside is None else outside
    w = np.zeros((nz + 2, ny + 2, nx + 2))
    w[1:-1, 1:-1, 1:-1] = u
    tx = np.arange(-1, nx + 1) // TX
    ty = (np.arange(-1, ny + 1) + shift[0]) // TY  # shift: the tile grid starts at (-shift y, -shift z)
    tz = (np.arange(-1, nz + 1) + shift[1]) // TZ
    for z in range(nz):
        for (cy, cx) in ((0, 0), (0, 1), (1, 0), (1, 1)):
            ys = np.arange(cy, ny, 2)
            xs = np.arange(cx, nx, 2)
            if len(ys) == 0 or len(xs) == 0:
                continue
            acc = np.zeros((len(ys), len(xs)))
            Sz = S[z][np.ix_(ys, xs)]
            for k, (ox, oy, oz) in enumerate(offs):
                if (ox, oy, oz) == (0, 0, 0):
                    continue
                a = Sz[..., k]
                if not a.any():
                    continue
                zz = z + oz
                same = (tz[1 + zz] == tz[1 + z]) & (ty[1 + ys + oy] == ty[1 + ys])[:, None] & (tx[1 + xs + ox] == tx[1 + xs])[None, :]
                vw = w[1 + zz][np.ix_(1 + ys + oy, 1 + xs + ox)]
                vo = old[1 + zz][np.ix_(1 + ys + oy, 1 + xs + ox)]
                acc += a * np.where(same, vw, vo)
            w[1 + z][np.ix_(1 + ys, 1 + xs)] = (f[z][np.ix_(ys, xs)] - acc) / diag[z][np.ix_(ys, xs)]
    return w[1:-1, 1:-1, 1:-1].copy()
