"""ctypes binding of the VED tensor front-end oracle (oracle/ved_oracle.c) and a Python restatement of
VEDMultigridImageFilter::GenerateData (itkVEDMultigridImageFilter.hxx:63-155) on top of it.

TEST INFRASTRUCTURE ONLY (see the header of ved_oracle.c for what is pinned and what is not).
Arrays are numpy float64, shape (nz, ny, nx); Hessians / tensors are AoS (nz, ny, nx, 6).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import oracle as O

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

DEFAULT_SCALES = (0.300, 0.482, 0.775, 1.245, 2.000)  # itkVEDMultigridImageFilter.hxx:52-58


class RgCoefs(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("N0", "N1", "N2", "N3", "D1", "D2", "D3", "D4", "M1", "M2", "M3", "M4",
                                          "BN1", "BN2", "BN3", "BN4", "BM1", "BM2", "BM3", "BM4")]


_ready = False


def lib():
    global _ready
    L = O.lib()
    if not _ready:
        L.vo_rg_setup.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int, C.POINTER(RgCoefs)]
        L.vo_rg_filter_line.argtypes = [C.POINTER(RgCoefs), _dp, _dp, _dp, C.c_int]
        L.vo_rg_filter_axis.argtypes = [_ip, C.c_int, C.POINTER(RgCoefs), _dp, _dp]
        L.vo_hessian.argtypes = [_ip, _dp, C.c_double, C.c_int, _dp, _dp]
        L.vo_eig3.argtypes = [_dp, _dp, _dp]
        L.vo_vesselness.restype = C.c_double
        L.vo_vesselness.argtypes = [_dp, C.c_double, C.c_double, C.c_double]
        L.vo_update_vesselness.argtypes = [C.c_int64, _dp, C.c_int, C.c_double, C.c_double, C.c_double, _dp, _dp, _dp]
        L.vo_generate_tensor.argtypes = [C.c_int64, _dp, _dp, C.c_double, C.c_double, C.c_double, _dp]
        _ready = True
    return L


def _d(a):
    return a.ctypes.data_as(_dp)


def rg_coefs(sigma, spacing, order, normalize_across_scale=True) -> RgCoefs:
    c = RgCoefs()
    lib().vo_rg_setup(float(sigma), float(spacing), int(order), int(bool(normalize_across_scale)), C.byref(c))
    return c


def rg_filter_line(coefs: RgCoefs, data) -> np.ndarray:
    data = np.ascontiguousarray(data, dtype=np.float64)
    out = np.empty_like(data)
    scratch = np.empty(2 * data.size)
    lib().vo_rg_filter_line(C.byref(coefs), _d(data), _d(out), _d(scratch), data.size)
    return out


def rg_filter_axis(volume, axis_xyz, coefs: RgCoefs) -> np.ndarray:
    """axis_xyz: 0 = x (fastest) ... 2 = z, volume shape (nz, ny, nx)."""
    v = np.ascontiguousarray(volume, dtype=np.float64)
    out = np.empty_like(v)
    n = (C.c_int * 3)(*v.shape[::-1])
    rc = lib().vo_rg_filter_axis(n, int(axis_xyz), C.byref(coefs), _d(v), _d(out))
    if rc:
        raise RuntimeError(f"vo_rg_filter_axis: {rc} (lines shorter than 4 samples are refused)")
    return out


def hessian(image, spacing_xyz, sigma, normalize_across_scale=True) -> np.ndarray:
    """ComputeHessian (itkVEDMultigridImageFilter.hxx:158-173) -> (nz, ny, nx, 6)."""
    img = np.ascontiguousarray(image, dtype=np.float64)
    out = np.empty(img.shape + (6,))
    n = (C.c_int * 3)(*img.shape[::-1])
    h = (C.c_double * 3)(*spacing_xyz)
    rc = lib().vo_hessian(n, h, float(sigma), int(bool(normalize_across_scale)), _d(img), _d(out))
    if rc:
        raise RuntimeError(f"vo_hessian: {rc}")
    return out


def eig3(a6):
    """a6 = (xx, xy, xz, yy, yz, zz) -> (w ascending, V with column k the eigenvector of w[k])."""
    a = np.ascontiguousarray(a6, dtype=np.float64)
    w, V = np.empty(3), np.empty(9)
    lib().vo_eig3(_d(a), _d(w), _d(V))
    return w, V.reshape(3, 3)


def vesselness(e_sorted_by_magnitude, alpha=0.5, beta=0.5, gamma=5.0) -> float:
    e = np.ascontiguousarray(e_sorted_by_magnitude, dtype=np.float64)
    return float(lib().vo_vesselness(_d(e), alpha, beta, gamma))


class VesselnessState:
    """m_MaxVesselnessResponse / m_MaxVesselnessEigenValues / m_MaxVesselnessEigenVectors
    (itkVEDMultigridImageFilter.h:136-139)."""

    def __init__(self, shape):
        self.shape = tuple(shape)
        self.response = None
        self.eigenvalues = np.zeros(self.shape + (3,))
        self.eigenvectors = np.zeros(self.shape + (3, 3))

    def update(self, hessian_aos, alpha=0.5, beta=0.5, gamma=5.0):
        h = np.ascontiguousarray(hessian_aos, dtype=np.float64)
        assert h.shape == self.shape + (6,)
        first = self.response is None
        if first:
            self.response = np.empty(self.shape)
        lib().vo_update_vesselness(int(np.prod(self.shape)), _d(h), int(first), alpha, beta, gamma, _d(self.response),
                                   _d(self.eigenvalues), _d(self.eigenvectors))

    def tensor(self, sensitivity=10.0, epsilon=0.01, omega=5.0) -> np.ndarray:
        out = np.empty(self.shape + (6,))
        lib().vo_generate_tensor(int(np.prod(self.shape)), _d(self.response), _d(self.eigenvectors), sensitivity, epsilon, omega,
                                 _d(out))
        return out


def ved_tensor(image, spacing_xyz, scales=DEFAULT_SCALES, alpha=0.5, beta=0.5, gamma=5.0, epsilon=0.01, omega=5.0,
               sensitivity=10.0, hessians=None):
    """One pass of the scale loop + GenerateDiffusionTensor (itkVEDMultigridImageFilter.hxx:110-120).
    Returns (tensor_aos, state)."""
    st = VesselnessState(np.shape(image))
    for i, s in enumerate(scales):
        H = hessians[i] if hessians is not None else hessian(image, spacing_xyz, s)
        st.update(H, alpha, beta, gamma)
    return st.tensor(sensitivity, epsilon, omega), st


def ved_filter(image, spacing_xyz, scales=DEFAULT_SCALES, alpha=0.5, beta=0.5, gamma=5.0, epsilon=0.01, omega=5.0,
               sensitivity=10.0, iterations=1, diffusion_iterations=5, smoother=0, cycle=0, time_step=0.1, tolerance=1e-6,
               iterations_per_grid=2, out_dtype=None):
    """VEDMultigridImageFilter::GenerateData (itkVEDMultigridImageFilter.hxx:63-155); defaults :33-58.
    Returns (output, info) with info = dict(cycles=[per outer iteration: cycles per time step], tensors=[...])."""
    img = np.array(image, dtype=np.float64)  # :70-100 cast to the internal pixel type
    info = dict(cycles=[], tensors=[])
    for _ in range(iterations):  # :105
        T, _st = ved_tensor(img, spacing_xyz, scales, alpha, beta, gamma, epsilon, omega, sensitivity)  # :108-120
        info["tensors"].append(T)
        # DiffusionStep, :381-402: MaxCycles = 100
        o = O.Oracle(img.shape, spacing_xyz, T, time_step, smoother=smoother, nu=iterations_per_grid)
        img, cyc, _ = o.solve(img, cycle=cycle, tolerance=tolerance, max_cycles=100, number_of_steps=diffusion_iterations)
        info["cycles"].append(cyc)
    if out_dtype is not None and np.issubdtype(np.dtype(out_dtype), np.integer):
        out = np.trunc(img).astype(out_dtype)  # static_cast< OutputPixelType >, :141
    elif out_dtype is not None:
        out = img.astype(out_dtype)
    else:
        out = img
    return out, info
