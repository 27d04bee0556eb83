"""ctypes binding of the CPU oracle (oracle/mad_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

Arrays are numpy float64, shape (nz, ny, nx) in 3-D / (ny, nx) in 2-D (x fastest, as ITK).
Tensors are AoS: shape (..., ncomp) with ncomp = 6 (xx,xy,xz,yy,yz,zz) or 3 (xx,xy,yy).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmadoracle.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, "mad_oracle.c"), os.path.join(_HERE, "ved_oracle.c")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.mo_create.restype = C.c_void_p
        L.mo_create.argtypes = [C.c_int, _ip, _dp, C.c_double, _dp, C.c_int, C.c_double, C.c_int, C.c_int]
        L.mo_destroy.argtypes = [C.c_void_p]
        L.mo_nlevels.argtypes = [C.c_void_p]
        L.mo_level_info.argtypes = [C.c_void_p, C.c_int, _ip, _dp, _ip]
        L.mo_level_stencil.restype = _dp
        L.mo_level_stencil.argtypes = [C.c_void_p, C.c_int]
        L.mo_level_tensor.restype = _dp
        L.mo_level_tensor.argtypes = [C.c_void_p, C.c_int]
        L.mo_set_smoother.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int]
        L.mo_smooth.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp]
        L.mo_residual.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp]
        L.mo_l2norm.restype = C.c_double
        L.mo_l2norm.argtypes = [_dp, C.c_int64]
        L.mo_restrict.argtypes = [C.c_int, _ip, _ip, _dp, _ip, _dp]
        L.mo_interpolate.argtypes = [C.c_int, _ip, _ip, _dp, _ip, _dp]
        L.mo_direct_solve.argtypes = [C.c_void_p, _dp, _dp]
        L.mo_vcycle.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, C.c_int]
        L.mo_fmg.argtypes = [C.c_void_p, _dp, _dp, C.c_int]
        L.mo_solve.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int, _dp, _ip, _dp, C.c_int]
        L.mo_set_verbose.argtypes = [C.c_void_p, C.c_int]
        L.mo_level_schedule.argtypes = [C.c_int, _ip, _ip, _ip]
        L.mo_generate_dca.argtypes = [C.c_int, _ip, _dp, C.c_double, _dp, _dp]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i3(v):
    return (C.c_int * 3)(*v)


def _xyz(shape, dim):
    """numpy shape (z,y,x) / (y,x) -> [nx, ny, nz]."""
    s = list(shape)[::-1]
    return s + [1] * (3 - len(s))


def level_schedule(size_xyz):
    dim = len(size_xyz)
    sizes = (C.c_int * (32 * 3))()
    cent = (C.c_int * (32 * 3))()
    n0 = (C.c_int * 3)(*(list(size_xyz) + [1] * (3 - dim)))
    nl = lib().mo_level_schedule(dim, n0, sizes, cent)
    return [(tuple(sizes[l * 3 + d] for d in range(dim)), tuple(cent[l * 3 + d] for d in range(dim))) for l in range(nl)]


def restrict(fine: np.ndarray, centering_xyz) -> np.ndarray:
    dim = fine.ndim
    fine = np.ascontiguousarray(fine, dtype=np.float64)
    nf = _xyz(fine.shape, dim)
    cent = list(centering_xyz) + [0] * (3 - dim)
    nc = [(nf[d] // 2 if cent[d] else (nf[d] - 1) // 2 + 1) if d < dim else 1 for d in range(3)]
    out = np.empty(nc[:dim][::-1], dtype=np.float64)
    lib().mo_restrict(dim, _i3(nf), _i3(cent), _d(fine), _i3(nc), _d(out))
    return out


def interpolate(coarse: np.ndarray, centering_xyz, fine_shape=None) -> np.ndarray:
    dim = coarse.ndim
    coarse = np.ascontiguousarray(coarse, dtype=np.float64)
    nc = _xyz(coarse.shape, dim)
    cent = list(centering_xyz) + [0] * (3 - dim)
    nf = [(nc[d] * 2 if cent[d] else (nc[d] - 1) * 2 + 1) if d < dim else 1 for d in range(3)]
    out = np.empty(nf[:dim][::-1], dtype=np.float64)
    lib().mo_interpolate(dim, _i3(nc), _i3(cent), _d(coarse), _i3(nf), _d(out))
    return out


class Oracle:
    """One GridsHierarchy + DirectSolver + smoother configuration (reference: the state built by
    MultigridAnisotropicDiffusionImageFilter::GenerateData, .hxx:131-156)."""

    GS, WJ = 0, 1
    VCYCLE, FMG, SMOOTHER = 0, 1, 2

    def __init__(self, shape, spacing_xyz, tensor_aos, time_step, smoother=0, omega=2.0 / 3.0, nu=2, max_coarse=0):
        self.dim = len(shape)
        self.shape = tuple(shape)
        n = _xyz(shape, self.dim)
        h = list(spacing_xyz) + [1.0] * (3 - self.dim)
        t = np.ascontiguousarray(tensor_aos, dtype=np.float64)
        ncomp = 3 if self.dim == 2 else 6
        assert t.shape == self.shape + (ncomp,), (t.shape, self.shape)
        self._h = lib().mo_create(self.dim, _i3(n), (C.c_double * 3)(*h), float(time_step), _d(t), int(smoother),
                                  float(omega), int(nu), int(max_coarse))
        if not self._h:
            raise RuntimeError("mo_create failed")
        self.nlevels = lib().mo_nlevels(self._h)
        self.levels = []
        for l in range(self.nlevels):
            nn, hh, cc = (C.c_int * 3)(), (C.c_double * 3)(), (C.c_int * 3)()
            lib().mo_level_info(self._h, l, nn, hh, cc)
            self.levels.append(dict(n=tuple(nn)[: self.dim], h=tuple(hh)[: self.dim], centering=tuple(cc)[: self.dim],
                                    shape=tuple(nn)[: self.dim][::-1]))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().mo_destroy(self._h)
                self._h = None
        except Exception:  # interpreter shutdown
            pass

    def set_smoother(self, smoother, omega=2.0 / 3.0, nu=2):
        lib().mo_set_smoother(self._h, int(smoother), float(omega), int(nu))

    def stencil(self, l) -> np.ndarray:
        shp = self.levels[l]["shape"]
        ns = 9 if self.dim == 2 else 27
        n = int(np.prod(shp)) * ns
        p = lib().mo_level_stencil(self._h, l)
        return np.ctypeslib.as_array(p, shape=(n,)).reshape(shp + (ns,)).copy()

    def tensor(self, l) -> np.ndarray:
        """SoA planes (ncomp, ...)."""
        shp = self.levels[l]["shape"]
        nc = 3 if self.dim == 2 else 6
        n = int(np.prod(shp)) * nc
        p = lib().mo_level_tensor(self._h, l)
        return np.ctypeslib.as_array(p, shape=(n,)).reshape((nc,) + shp).copy()

    def _chk(self, a, l):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.shape == self.levels[l]["shape"], (a.shape, self.levels[l]["shape"])
        return a

    def smooth(self, l, u, f):
        u, f = self._chk(u, l), self._chk(f, l)
        out = np.empty_like(u)
        lib().mo_smooth(self._h, l, _d(u), _d(f), _d(out))
        return out

    def residual(self, l, u, f):
        u, f = self._chk(u, l), self._chk(f, l)
        out = np.empty_like(u)
        lib().mo_residual(self._h, l, _d(u), _d(f), _d(out))
        return out

    def direct_solve(self, f):
        l = self.nlevels - 1
        f = self._chk(f, l)
        out = np.empty_like(f)
        lib().mo_direct_solve(self._h, _d(f), _d(out))
        return out

    def vcycle(self, u, f, level=0, faithful=False):
        u, f = self._chk(u, level), self._chk(f, level)
        out = np.empty_like(u)
        lib().mo_vcycle(self._h, level, _d(u), _d(f), _d(out), int(faithful))
        return out

    def fmg(self, f, faithful=False):
        f = self._chk(f, 0)
        out = np.empty_like(f)
        lib().mo_fmg(self._h, _d(f), _d(out), int(faithful))
        return out

    def solve(self, image, cycle=0, tolerance=1e-6, max_cycles=100, number_of_steps=1, faithful=False, verbose=False):
        img = np.array(image, dtype=np.float64, copy=True, order="C")
        assert img.shape == self.shape
        cyc = (C.c_int * max(number_of_steps, 1))()
        hist = np.full((max(number_of_steps, 1), max_cycles), np.nan)
        lib().mo_set_verbose(self._h, int(verbose))
        rc = lib().mo_solve(self._h, int(cycle), float(tolerance), int(max_cycles), int(number_of_steps), _d(img), cyc,
                            _d(hist), int(faithful))
        if rc != 0:
            raise RuntimeError(f"mo_solve failed: {rc}")
        return img, list(cyc)[:number_of_steps], hist


def l2norm(a) -> float:
    a = np.ascontiguousarray(a, dtype=np.float64)
    return float(lib().mo_l2norm(_d(a), a.size))
