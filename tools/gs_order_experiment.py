#!/usr/bin/env python
"""Design experiment (CPU, numpy): V-cycle convergence of Gauss-Seidel ORDERINGS at level 0.

Compares, inside the same V(nu,nu) cycle (coarse part = the oracle's own levels):
  lex      the reference's lexicographic sweep (oracle)
  color4   global 4-colour multicolour GS (v1 kernels)
  tile     z-lexicographic planes, in-plane even rows (even x, odd x) then odd rows, inside
           (TX,TY,TZ) tiles; values outside the tile are read from the previous sweep (the fused
           B200 sweep kernel)
"""
import sys, os, itertools
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from util import load_ved_test
from multigridanisotropicdiffusion_b200 import phantom

OFFS = [(ox, oy, oz) for oz in (-1, 0, 1) for oy in (-1, 0, 1) for ox in (-1, 0, 1)]


def offdiag_sum(S, w, sel=None, getter=None):
    """sum_k a_k u_k over off-diagonal entries; w padded by 1."""
    nz, ny, nx = S.shape[:3]
    acc = np.zeros((nz, ny, nx))
    for k, (ox, oy, oz) in enumerate(OFFS):
        if (ox, oy, oz) == (0, 0, 0):
            continue
        a = S[..., k]
        if not a.any():
            continue
        acc += a * getter(ox, oy, oz)
    return acc


def sweep_color4(S, u, f):
    nz, ny, nx = u.shape
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    p = (x & 1) | ((y & 1) << 1) | ((z & 1) << 2)
    col = np.minimum(p, 7 - p)
    w = np.zeros((nz + 2, ny + 2, nx + 2)); w[1:-1, 1:-1, 1:-1] = u
    diag = S[..., 13]
    for c in range(4):
        get = lambda ox, oy, oz: w[1 + oz:nz + 1 + oz, 1 + oy:ny + 1 + oy, 1 + ox:nx + 1 + ox]
        off = offdiag_sum(S, w, getter=get)
        new = (f - off) / diag
        m = col == c
        w[1:-1, 1:-1, 1:-1][m] = new[m]
    return w[1:-1, 1:-1, 1:-1].copy()


def sweep_tile(S, u, f, tile):
    TX, TY, TZ = tile
    nz, ny, nx = u.shape
    diag = S[..., 13]
    old = np.zeros((nz + 2, ny + 2, nx + 2)); old[1:-1, 1:-1, 1:-1] = u
    w = old.copy()
    tx = np.arange(-1, nx + 1) // TX; ty = np.arange(-1, ny + 1) // TY; tz = np.arange(-1, nz + 1) // TZ
    for z in range(nz):
        for (cy, cx) in ((0, 0), (0, 1), (1, 0), (1, 1)):
            ys = np.arange(cy, ny, 2); xs = np.arange(cx, nx, 2)
            acc = np.zeros((len(ys), len(xs)))
            for k, (ox, oy, oz) in enumerate(OFFS):
                if (ox, oy, oz) == (0, 0, 0):
                    continue
                a = S[z][np.ix_(ys, xs)][..., k]
                if not a.any():
                    continue
                zz = z + oz
                same = (tz[1 + zz] == tz[1 + z]) & (ty[1 + ys + oy] == ty[1 + ys])[:, None] & (tx[1 + xs + ox] == tx[1 + xs])[None, :]
                vw = w[1 + zz][np.ix_(1 + ys + oy, 1 + xs + ox)]
                vo = old[1 + zz][np.ix_(1 + ys + oy, 1 + xs + ox)]
                acc += a * np.where(same, vw, vo)
            new = (f[z][np.ix_(ys, xs)] - acc) / diag[z][np.ix_(ys, xs)]
            w[1 + z][np.ix_(1 + ys, 1 + xs)] = new
    return w[1:-1, 1:-1, 1:-1].copy()


def run(shape, T, sp, img, nu, label, smoother, ncyc=12):
    o = O.Oracle(shape, sp, T, 0.1, smoother=0, nu=nu)
    S = o.stencil(0)
    f = img.astype(np.float64)
    u = f.copy()
    fn = np.linalg.norm(f)
    cent = o.levels[1]["centering"]
    hist = []
    for c in range(ncyc):
        if smoother == "lex":
            u = o.vcycle(u, f)
        else:
            for _ in range(nu):
                u = smoother(S, u, f)
            r = o.residual(0, u, f)
            rc = O.restrict(r, cent)
            ec = o.vcycle(np.zeros_like(rc), rc, level=1)
            u = u + O.interpolate(ec, cent, None)[tuple(slice(0, s) for s in shape)]
            for _ in range(nu):
                u = smoother(S, u, f)
        rr = np.linalg.norm(o.residual(0, u, f)) / fn
        hist.append(rr)
        if rr < 1e-10:
            break
    print(f"{label:28s} cycles={len(hist):2d}  " + " ".join(f"{h:.1e}" for h in hist), flush=True)
    return u


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "ved"
    if which == "ved":
        img, sp = load_ved_test()
        shape = img.shape
        _, D = phantom.vessel_phantom(shape)
        T = phantom.planes_to_aos(D).numpy().astype(np.float64)
    else:
        n = int(which)
        shape = (n, n, n); sp = phantom.VED_SPACING
        im, D = phantom.vessel_phantom(shape)
        img = im.numpy(); T = phantom.planes_to_aos(D).numpy().astype(np.float64)
    nu = 3
    ref = run(shape, T, sp, img, nu, "lex (reference order)", "lex")
    a = run(shape, T, sp, img, nu, "color4 (global)", sweep_color4)
    print("   rel diff vs lex:", np.linalg.norm(a - ref) / np.linalg.norm(ref))
    for tile in ((128, 8, 32), (32, 8, 16), (16, 4, 8), (8, 2, 4)):
        b = run(shape, T, sp, img, nu, f"tile {tile}", lambda S, u, f: sweep_tile(S, u, f, tile))
        print("   rel diff vs lex:", np.linalg.norm(b - ref) / np.linalg.norm(ref))
