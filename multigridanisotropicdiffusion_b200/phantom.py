"""Synthetic inputs of the shapes BASELINE.json names.

3-D: a tubular-vessel phantom (three helices of different calibre + Gaussian noise, statistics
close to the reference's test/test_data/ved_test.mhd) and an analytic diffusion tensor in the
form VEDMultigridImageFilter::GenerateDiffusionTensor builds
(/root/reference/include/itkVEDMultigridImageFilter.hxx:327-346):
    D = (1 + (eps-1) V) (I - t t^T) + (1 + (omega-1) V) t t^T,   identity where V = 0 (:357-366)
with V = vesselness^(1/sensitivity) and t the centre-line tangent, so all six components are
non-zero and vary in space (cross terms and first-derivative terms of the operator are exercised).

2-D: a rotating anisotropic tensor (D_xy != 0), and the constant tensor of the reference's 2-D tests
(test/itk2DDiffusionTest_WJ.cxx:66-73).

Everything is written with torch so the same code generates on the CPU for the parity tests and
directly in HBM for the benchmark (torch is plumbing here: memory and a Philox generator).
"""
from __future__ import annotations

import math

import torch

VED_SPACING = (0.3125, 0.3125, 0.5)  # test/test_data/ved_test.mhd ElementSpacing


def vessel_phantom(shape, device="cpu", seed=1234, spacing=VED_SPACING, eps=0.01, omega=1.5, sensitivity=10.0,
                   noise_sigma=20.0, dtype=torch.float32, z_range=None):
    """Returns (image[nz,ny,nx] float32, tensor planes[6,nz,ny,nx] float32: xx,xy,xz,yy,yz,zz).
    z_range=(z0, z1): only that slab of planes of the `shape` volume (z-slab decomposition; the noise of a slab is drawn
    from its own stream, the vessels and the tensor are identical to the whole volume's)."""
    nz_global, ny, nx = shape
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    z0, z1 = (0, nz_global) if z_range is None else z_range
    gen.manual_seed(seed if z_range is None else seed + 7919 * (z0 + 1))
    nz = z1 - z0
    shape = (nz, ny, nx)
    z = torch.arange(z0, z1, device=dev, dtype=torch.float32).view(nz, 1, 1) * spacing[2]
    y = torch.arange(ny, device=dev, dtype=torch.float32).view(1, ny, 1) * spacing[1]
    x = torch.arange(nx, device=dev, dtype=torch.float32).view(1, 1, nx) * spacing[0]
    Lx, Ly, Lz = nx * spacing[0], ny * spacing[1], nz_global * spacing[2]
    cx, cy = 0.5 * Lx, 0.5 * Ly
    R = 0.25 * min(Lx, Ly)
    turns = 2.0
    dth = 2.0 * math.pi * turns / max(Lz, 1e-9)  # d(theta)/dz
    h = min(spacing[0], spacing[1])
    best = torch.zeros(shape, device=dev, dtype=torch.float32)
    tx = torch.zeros(shape, device=dev, dtype=torch.float32)
    ty = torch.zeros(shape, device=dev, dtype=torch.float32)
    tz = torch.ones(shape, device=dev, dtype=torch.float32)
    img = torch.zeros(shape, device=dev, dtype=torch.float32)
    for k, rr in enumerate((2.0, 4.0, 6.0)):
        r = rr * h
        th = dth * z + 2.0 * math.pi * k / 3.0
        px, py = cx + R * torch.cos(th), cy + R * torch.sin(th)
        d2 = (x - px) ** 2 + (y - py) ** 2
        v = torch.exp(-d2 / (2.0 * r * r))
        img += 150.0 * v
        # unit tangent of the helix at this z
        ax, ay = -R * dth * torch.sin(th), R * dth * torch.cos(th)
        nrm = torch.sqrt(ax * ax + ay * ay + 1.0)
        sel = v > best
        best = torch.where(sel, v, best)
        tx = torch.where(sel, (ax / nrm).expand(shape), tx)
        ty = torch.where(sel, (ay / nrm).expand(shape), ty)
        tz = torch.where(sel, (1.0 / nrm).expand(shape), tz)
        del d2, v, sel
    img += 30.0
    img += noise_sigma * torch.randn(shape, device=dev, dtype=torch.float32, generator=gen)
    V = torch.where(best > 1e-3, best.clamp(max=1.0) ** (1.0 / sensitivity), torch.zeros_like(best))
    lam_perp = 1.0 + (eps - 1.0) * V
    dl = (1.0 + (omega - 1.0) * V) - lam_perp
    D = torch.empty((6,) + tuple(shape), device=dev, dtype=torch.float32)
    D[0] = lam_perp + dl * tx * tx
    D[1] = dl * tx * ty
    D[2] = dl * tx * tz
    D[3] = lam_perp + dl * ty * ty
    D[4] = dl * ty * tz
    D[5] = lam_perp + dl * tz * tz
    return img.to(dtype), D


def planes_to_aos(D: torch.Tensor) -> torch.Tensor:
    """[ncomp, ...] SoA planes -> ITK AoS buffer [..., ncomp]."""
    return D.movedim(0, -1).contiguous()


def rotating_tensor_2d(shape, lam1=50.0, lam2=5.0, device="cpu"):
    """2-D tensor planes [3, ny, nx] (xx, xy, yy) whose principal axis rotates over the image."""
    ny, nx = shape
    dev = torch.device(device)
    y = torch.arange(ny, device=dev, dtype=torch.float32).view(ny, 1) / max(ny - 1, 1)
    x = torch.arange(nx, device=dev, dtype=torch.float32).view(1, nx) / max(nx - 1, 1)
    th = math.pi * (0.75 * x + 0.5 * y) + 0.3 * torch.sin(2.0 * math.pi * x * y)
    c, s = torch.cos(th), torch.sin(th)
    l1 = lam1 * (1.0 + 0.3 * torch.cos(3.0 * math.pi * y))
    D = torch.empty((3, ny, nx), device=dev, dtype=torch.float32)
    D[0] = l1 * c * c + lam2 * s * s
    D[1] = (l1 - lam2) * c * s
    D[2] = l1 * s * s + lam2 * c * c
    return D


def constant_tensor_2d(shape, dxx=50.0, dyy=30.0, dxy=0.0, device="cpu"):
    """The reference's 2-D test tensor (test/itk2DDiffusionTest_WJ.cxx:66-73)."""
    D = torch.empty((3,) + tuple(shape), device=device, dtype=torch.float32)
    D[0] = dxx
    D[1] = dxy
    D[2] = dyy
    return D
