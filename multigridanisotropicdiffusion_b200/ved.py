"""Host-side handle on one VED front-end context of libmadgpu.so (C-ABI in include/madved.h): Hessian at several
scales, vesselness, diffusion tensor -- the steps of itk::VEDMultigridImageFilter that precede the multigrid solve
(/root/reference/include/itkVEDMultigridImageFilter.hxx:158-378).  Every method is one C-ABI call.  numpy arrays are
(nz, ny, nx), x fastest; Hessians / tensors are AoS (nz, ny, nx, 6).  No CPU fallback."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as B
from .solver import _PIX, MadGpuError, MadSolver, _ptr

DEFAULT_SCALES = (0.300, 0.482, 0.775, 1.245, 2.000)  # itkVEDMultigridImageFilter.hxx:52-58


class MadVed:
    def __init__(self, shape, spacing_xyz=None, alpha=0.5, beta=0.5, gamma=5.0, epsilon=0.01, omega=5.0, sensitivity=10.0, device=0):
        self._lib = B.load()
        self._ctx = C.c_void_p()
        self.shape = tuple(int(s) for s in shape)
        if len(self.shape) != 3:
            raise MadGpuError("the VED filter is 3-D (itkVEDMultigridImageFilter.h:45)")
        p = B.VedParams()
        self._lib.madved_params_default(C.byref(p))
        sp = list(spacing_xyz) if spacing_xyz is not None else [1.0, 1.0, 1.0]
        for d in range(3):
            p.size[d] = self.shape[::-1][d]
            p.spacing[d] = float(sp[d])
        p.alpha, p.beta, p.gamma = float(alpha), float(beta), float(gamma)
        p.epsilon, p.omega, p.sensitivity = float(epsilon), float(omega), float(sensitivity)
        p.device = int(device)
        self.params = p
        rc = self._lib.madved_create(C.byref(p), C.byref(self._ctx))
        if rc != 0:
            msg = self._lib.madved_last_error(None)
            self._ctx = C.c_void_p()
            raise MadGpuError(f"madved_create failed ({rc}): {msg.decode() if msg else ''}")

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._lib.madved_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc, what):
        if rc != 0:
            msg = self._lib.madved_last_error(self._ctx)
            raise MadGpuError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def set_params(self, alpha, beta, gamma, epsilon, omega, sensitivity):
        self._check(self._lib.madved_set_params(self._ctx, alpha, beta, gamma, epsilon, omega, sensitivity), "set_params")

    def set_image(self, image: np.ndarray):
        img = np.ascontiguousarray(image)
        if img.shape != self.shape:
            raise MadGpuError(f"image shape {img.shape} != {self.shape}")
        if img.dtype not in _PIX:
            raise MadGpuError(f"unsupported pixel type {img.dtype}")
        self._check(self._lib.madved_set_image(self._ctx, _PIX[img.dtype], _ptr(img)), "set_image")

    def set_image_device(self, d_image: int):
        self._check(self._lib.madved_set_image_device_f32(self._ctx, C.c_void_p(int(d_image))), "set_image_device")

    def image_device(self) -> int:
        p = C.c_void_p()
        self._check(self._lib.madved_image_device(self._ctx, C.byref(p)), "image_device")
        return int(p.value)

    def begin(self):
        self._check(self._lib.madved_begin(self._ctx), "begin")

    def hessian(self, sigma):
        """ComputeHessian (hxx:158-173) of the current image into the context's Hessian planes."""
        self._check(self._lib.madved_hessian(self._ctx, float(sigma)), "hessian")

    def update_vesselness(self, hessian_aos=None):
        """UpdateVesselness (hxx:215-299): on the context's Hessian planes, or on a host Hessian (nz, ny, nx, 6) float64."""
        if hessian_aos is None:
            self._check(self._lib.madved_update_vesselness(self._ctx), "update_vesselness")
        else:
            h = np.ascontiguousarray(hessian_aos, dtype=np.float64)
            if h.shape != self.shape + (6,):
                raise MadGpuError(f"Hessian shape {h.shape} != {self.shape + (6,)}")
            self._check(self._lib.madved_update_vesselness_host_f64(self._ctx, _ptr(h)), "update_vesselness_host")

    def add_scale(self, sigma):
        self._check(self._lib.madved_add_scale(self._ctx, float(sigma)), "add_scale")

    def tensor_planes(self):
        """Six device pointers (ints) to the dense fp32 tensor planes: MadSolver.set_tensor_device takes them as they are."""
        arr = (C.c_void_p * 6)()
        self._check(self._lib.madved_tensor_planes(self._ctx, arr), "tensor_planes")
        return [int(a) for a in arr]

    def get_tensor(self) -> np.ndarray:
        out = np.empty(self.shape + (6,), dtype=np.float64)
        self._check(self._lib.madved_get_tensor_f64(self._ctx, _ptr(out)), "get_tensor")
        return out

    def get_response(self) -> np.ndarray:
        out = np.empty(self.shape, dtype=np.float64)
        self._check(self._lib.madved_get_response_f64(self._ctx, _ptr(out)), "get_response")
        return out

    def get_hessian(self) -> np.ndarray:
        out = np.empty(self.shape + (6,), dtype=np.float64)
        self._check(self._lib.madved_get_hessian_f64(self._ctx, _ptr(out)), "get_hessian")
        return out

    def stats(self) -> dict:
        st = B.VedStats()
        st.struct_size = C.sizeof(B.VedStats)
        self._check(self._lib.madved_get_stats(self._ctx, C.byref(st)), "get_stats")
        return dict(scales=st.scales, hessian_ms=st.hessian_ms, vesselness_ms=st.vesselness_ms, h2d_ms=st.h2d_ms, d2h_ms=st.d2h_ms,
                    diffusion_ms=st.diffusion_ms, kernel_launches=st.kernel_launches)

    def run(self, solver: MadSolver, image: np.ndarray, scales=DEFAULT_SCALES, iterations=1, out_dtype=None, out: np.ndarray = None):
        """VEDMultigridImageFilter::GenerateData (hxx:63-155) on the device; `solver` carries the DiffusionStep settings."""
        img = np.ascontiguousarray(image)
        if img.shape != self.shape or solver.shape != self.shape:
            raise MadGpuError(f"image shape {img.shape} / solver shape {solver.shape} != {self.shape}")
        if img.dtype not in _PIX:
            raise MadGpuError(f"unsupported pixel type {img.dtype}")
        odt = np.dtype(out_dtype) if out_dtype is not None else (out.dtype if out is not None else img.dtype)
        if out is None:
            out = np.empty(self.shape, dtype=odt)
        elif out.shape != self.shape or out.dtype != odt or not out.flags.c_contiguous:
            raise MadGpuError("out must be C-contiguous with the image shape and the output pixel type")
        sc = (C.c_double * len(scales))(*[float(s) for s in scales])
        st = B.Stats()
        st.struct_size = C.sizeof(B.Stats)
        self._check(self._lib.madved_run(self._ctx, solver._ctx, _PIX[img.dtype], _ptr(img), _PIX[odt], _ptr(out), sc, len(scales),
                                         int(iterations), C.byref(st)), "run")
        solver._stats(st)
        return out
