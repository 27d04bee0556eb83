"""ctypes binding of oracle/_ref/libmadref.so: the UNMODIFIED reference headers (/root/reference/include)
compiled against the stand-in ITK of oracle/shim (oracle/Makefile, `make ref`).

TEST INFRASTRUCTURE ONLY -- used to pin the C restatement (oracle/mad_oracle.c) and to generate the golden
vectors under tests/golden (tests/golden/make_golden.py).  The library is built in the authoring container
(where /root/reference exists) and travels to the GPU box as a built artefact; `available()` says whether
it is there.  Arrays: numpy float64, (nz, ny, nx) / (ny, nx), x fastest; tensors AoS (..., ncomp).
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libmadref.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def build() -> str:
    """Only possible where /root/reference exists."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.mr_create.restype = C.c_void_p
        L.mr_create.argtypes = [C.c_int, _ip, _dp, C.c_double, _dp]
        L.mr_destroy.argtypes = [C.c_void_p]
        L.mr_nlevels.argtypes = [C.c_void_p]
        L.mr_level_info.argtypes = [C.c_void_p, C.c_int, _ip, _dp, _ip]
        L.mr_level_stencil.argtypes = [C.c_void_p, C.c_int, _dp, _ip]
        L.mr_smooth.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp]
        L.mr_residual.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp]
        L.mr_direct_solve.argtypes = [C.c_void_p, _dp, _dp]
        L.mr_transfer.argtypes = [C.c_int, C.c_int, _ip, _ip, _dp, _dp, _ip]
        L.mr_filter.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _dp, _dp, _dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int,
                                C.c_int, C.c_int, _dp, C.c_char_p, C.c_int]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i3(v):
    return (C.c_int * 3)(*v)


def _xyz(shape):
    s = list(shape)[::-1]
    return s + [1] * (3 - len(s))


def _transfer(what, a, centering_xyz):
    dim = a.ndim
    a = np.ascontiguousarray(a, dtype=np.float64)
    n = _i3(_xyz(a.shape))
    cent = _i3(list(centering_xyz) + [0] * (3 - dim))
    nout = (C.c_int * 3)()
    lib().mr_transfer(dim, what, n, cent, _d(a), None, nout)
    out = np.empty(tuple(nout)[:dim][::-1], dtype=np.float64)
    lib().mr_transfer(dim, what, n, cent, _d(a), _d(out), nout)
    return out


def restrict(fine, centering_xyz):
    """itk::mad::InterGridOperators::Restriction"""
    return _transfer(0, fine, centering_xyz)


def interpolate(coarse, centering_xyz):
    """itk::mad::InterGridOperators::Interpolation"""
    return _transfer(1, coarse, centering_xyz)


class Reference:
    """itk::mad::GridsHierarchy (+ DirectSolver on demand) built by the reference's own constructor."""

    GS, WJ = 0, 1

    def __init__(self, shape, spacing_xyz, tensor_aos, time_step):
        self.dim = len(shape)
        self.shape = tuple(shape)
        t = np.ascontiguousarray(tensor_aos, dtype=np.float64)
        h = list(spacing_xyz) + [1.0] * (3 - self.dim)
        self._h = lib().mr_create(self.dim, _i3(_xyz(shape)), (C.c_double * 3)(*h), float(time_step), _d(t))
        if not self._h:
            raise RuntimeError("mr_create failed")
        self.nlevels = lib().mr_nlevels(self._h)
        self.levels = []
        for l in range(self.nlevels):
            nn, hh, cc = (C.c_int * 3)(), (C.c_double * 3)(), (C.c_int * 3)()
            lib().mr_level_info(self._h, l, nn, hh, cc)
            self.levels.append(dict(n=tuple(nn)[: self.dim], h=tuple(hh)[: self.dim], centering=tuple(cc)[: self.dim],
                                    shape=tuple(nn)[: self.dim][::-1]))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().mr_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def stencil(self, l):
        """(rows [..., 3^dim] in Neighborhood raster order, active[3^dim]: 1-based position in the active-offset list or 0)"""
        shp = self.levels[l]["shape"]
        ns = 3 ** self.dim
        out = np.empty(shp + (ns,), dtype=np.float64)
        act = (C.c_int * ns)()
        lib().mr_level_stencil(self._h, l, _d(out), act)
        return out, np.array(list(act))

    def _chk(self, a, l):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.shape == self.levels[l]["shape"]
        return a

    def smooth(self, l, u, f, smoother):
        u, f = self._chk(u, l), self._chk(f, l)
        out = np.empty_like(u)
        lib().mr_smooth(self._h, l, int(smoother), _d(u), _d(f), _d(out))
        return out

    def residual(self, l, u, f, smoother=0):
        u, f = self._chk(u, l), self._chk(f, l)
        out = np.empty_like(u)
        lib().mr_residual(self._h, l, int(smoother), _d(u), _d(f), _d(out))
        return out

    def direct_solve(self, f):
        f = self._chk(f, self.nlevels - 1)
        out = np.empty_like(f)
        if lib().mr_direct_solve(self._h, _d(f), _d(out)) != 0:
            raise RuntimeError("mr_direct_solve failed")
        return out


_PIXEL = {"double": 0, "float": 1, "short": 2, "uchar": 3}


def run_filter(image, spacing_xyz, tensor_aos, smoother=0, cycle=0, nu=2, time_step=0.01, tolerance=1e-6, max_cycles=100,
               number_of_steps=1, pixel="double", verbose=True):
    """itk::MultigridAnisotropicDiffusionImageFilter<Image<pixel>, Image<pixel>, smoother>::Update(), driven like the
    reference's test programs.  Returns (output as float64, cycles per time step, relres per cycle per step, log)."""
    img = np.ascontiguousarray(image, dtype=np.float64)
    dim = img.ndim
    t = np.ascontiguousarray(tensor_aos, dtype=np.float64)
    h = list(spacing_xyz) + [1.0] * (3 - dim)
    out = np.empty_like(img)
    cap = 1 << 24
    buf = C.create_string_buffer(cap)
    rc = lib().mr_filter(dim, _PIXEL[pixel], int(smoother), _i3(_xyz(img.shape)), (C.c_double * 3)(*h), _d(t), _d(img), int(cycle),
                         int(nu), float(time_step), float(tolerance), int(max_cycles), int(number_of_steps), int(verbose), _d(out), buf,
                         cap)
    if rc < 0:
        raise RuntimeError(f"mr_filter failed: {rc}")
    log = buf.value.decode()
    # one outer iteration = one "|--- VCycle n. k ---|" (or "Smoother iteration n. k") line; time steps are separated by
    # "------------ Time step n." lines when there are several
    cycles, cur = [], 0
    for ln in log.splitlines():
        if ln.startswith("------------ Time step"):
            if cur:
                cycles.append(cur)
            cur = 0
        elif ln.startswith("|--- VCycle n.") or ln.startswith("Smoother iteration n."):
            cur += 1
    cycles.append(cur)
    return out, cycles, log


def relres_per_cycle(log, smoother_mode=False):
    """Stop-test relative residual after every outer iteration, from the verbose log: in SMOOTHER mode the
    "Smoother iteration" lines; otherwise the last "Level 0, iteration" line of each V-cycle (its norm and its
    right-hand side are the ones of the stop test, …Filter.hxx:239 vs :464-466)."""
    steps, cur, last, in_cycle = [], [], None, smoother_mode
    for ln in log.splitlines():
        if ln.startswith("------------ Time step"):
            if last is not None:
                cur.append(last)
            if cur:
                steps.append(cur)
            cur, last, in_cycle = [], None, smoother_mode
        elif ln.startswith("Smoother iteration n."):
            cur.append(float(ln.rsplit("=", 1)[1]))
        elif ln.startswith("|--- VCycle n."):
            if last is not None:
                cur.append(last)
            last, in_cycle = None, True
        elif in_cycle and (ln.startswith(" Level 0, iteration") or ln.startswith(" Level 0, direct solver")):
            last = float(ln.rsplit("=", 1)[1])
    if last is not None:
        cur.append(last)
    steps.append(cur)
    return steps


# ---- VED filter (itkVEDMultigridImageFilter.{h,hxx} compiled unmodified; Hessian filter and eigen-solver underneath are
# ---- the stand-ins of oracle/shim/mini_itk_ved.h, i.e. oracle/ved_oracle.c's vo_hessian / vo_eig3) ----
_ved_ready = False


def _ved_lib():
    global _ved_ready
    L = lib()
    if not _ved_ready:
        L.mrv_vesselness.restype = C.c_double
        L.mrv_vesselness.argtypes = [_dp, C.c_double, C.c_double, C.c_double]
        L.mrv_tensor_from_hessians.argtypes = [_ip, _dp, _dp, C.c_int, _dp, _dp, _dp, _dp, _dp]
        L.mrv_filter.argtypes = [C.c_int, C.c_int, _ip, _dp, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                 C.c_int, _dp, _dp]
        _ved_ready = True
    return L


def ved_available() -> bool:
    return available() and hasattr(lib(), "mrv_filter")


def ved_vesselness(e_sorted_by_magnitude, alpha=0.5, beta=0.5, gamma=5.0) -> float:
    """VEDMultigridImageFilter::VesselnessFunction"""
    e = np.ascontiguousarray(e_sorted_by_magnitude, dtype=np.float64)
    return float(_ved_lib().mrv_vesselness(_d(e), alpha, beta, gamma))


def ved_tensor_from_hessians(hessians, spacing_xyz, alpha=0.5, beta=0.5, gamma=5.0, epsilon=0.01, omega=5.0, sensitivity=10.0):
    """UpdateVesselness once per Hessian (list of (nz, ny, nx, 6) arrays), then GenerateDiffusionTensor.
    Returns dict(response, eigenvalues, eigenvectors, tensor)."""
    hs = np.ascontiguousarray(np.stack([np.asarray(h, dtype=np.float64) for h in hessians]))
    shape = hs.shape[1:4]
    p = (C.c_double * 6)(alpha, beta, gamma, epsilon, omega, sensitivity)
    out = dict(response=np.empty(shape), eigenvalues=np.empty(shape + (3,)), eigenvectors=np.empty(shape + (3, 3)),
               tensor=np.empty(shape + (6,)))
    rc = _ved_lib().mrv_tensor_from_hessians(_i3(_xyz(shape)), (C.c_double * 3)(*spacing_xyz), _d(hs), hs.shape[0], p, _d(out["response"]),
                                             _d(out["eigenvalues"]), _d(out["eigenvectors"]), _d(out["tensor"]))
    if rc:
        raise RuntimeError(f"mrv_tensor_from_hessians failed: {rc}")
    return out


def run_ved_filter(image, spacing_xyz, scales, alpha=0.5, beta=0.5, gamma=5.0, epsilon=0.01, omega=5.0, sensitivity=10.0, iterations=1,
                   diffusion_iterations=5, smoother=0, cycle=0, time_step=0.1, tolerance=1e-6, iterations_per_grid=2, pixel="double"):
    """itk::VEDMultigridImageFilter<Image<pixel,3>, Image<pixel,3>, smoother>::Update(), driven like test/itkVEDTest_GS.cxx.
    Returns (output as float64, diffusion tensor of the last outer iteration)."""
    img = np.ascontiguousarray(image, dtype=np.float64)
    out = np.empty_like(img)
    T = np.empty(img.shape + (6,))
    p = (C.c_double * 6)(alpha, beta, gamma, epsilon, omega, sensitivity)
    sc = (C.c_double * len(scales))(*scales)
    rc = _ved_lib().mrv_filter(_PIXEL[pixel], int(smoother), _i3(_xyz(img.shape)), (C.c_double * 3)(*spacing_xyz), _d(img), p, sc, len(scales),
                               int(iterations), int(diffusion_iterations), int(cycle), float(time_step), float(tolerance),
                               int(iterations_per_grid), _d(out), _d(T))
    if rc:
        raise RuntimeError(f"mrv_filter failed: {rc}")
    return out, T
