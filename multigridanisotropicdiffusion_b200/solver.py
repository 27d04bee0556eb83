"""Host-side handle on one libmadgpu context (one GridsHierarchy + DirectSolver on one B200).

`MadSolver` is the thin object the filter classes (filter.py) and the parity tests drive; every
method is one C-ABI call (include/madgpu.h).  numpy arrays are (nz, ny, nx) / (ny, nx), x fastest,
exactly the ITK buffer order.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as B


class MadGpuError(RuntimeError):
    pass


_PIX = {np.dtype(np.uint8): B.PIX_U8, np.dtype(np.int16): B.PIX_I16, np.dtype(np.float32): B.PIX_F32,
        np.dtype(np.float64): B.PIX_F64}


def _ptr(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


class MadSolver:
    GS, WJ = B.SMOOTHER_GS, B.SMOOTHER_WJ
    VCYCLE, FMG, SMOOTHER = B.CYCLE_V, B.CYCLE_FMG, B.CYCLE_SMOOTHER

    def __init__(self, shape, spacing_xyz=None, time_step=0.01, smoother=B.SMOOTHER_GS, omega=2.0 / 3.0,
                 iterations_per_grid=2, cycle=B.CYCLE_V, tolerance=1e-6, max_cycles=100, number_of_steps=1,
                 verbose=False, gs_colors=4, device=0, rank=0, world_size=1, nccl_id=None):
        """shape: the (global) image shape (nz, ny, nx) / (ny, nx).  rank / world_size / nccl_id: this context owns one
        z-slab of the volume (madgpu_create_slab; see slabs.py); images and tensors passed to it are the local slab."""
        self._lib = B.load()
        self._ctx = C.c_void_p()
        self.global_shape = tuple(int(s) for s in shape)
        self.shape = self.global_shape
        self.rank, self.world_size = int(rank), int(world_size)
        self.dim = len(self.shape)
        if self.dim not in (2, 3):
            raise MadGpuError("images must be 2-D or 3-D")
        p = B.Params()
        self._lib.madgpu_params_default(C.byref(p))
        p.dim = self.dim
        size = list(self.shape[::-1]) + [1] * (3 - self.dim)
        sp = list(spacing_xyz) if spacing_xyz is not None else [1.0] * self.dim
        sp = sp + [1.0] * (3 - self.dim)
        for d in range(3):
            p.size[d] = size[d]
            p.spacing[d] = float(sp[d])
        p.time_step = float(time_step)
        p.smoother = int(smoother)
        p.omega = float(omega)
        p.iterations_per_grid = int(iterations_per_grid)
        p.cycle = int(cycle)
        p.tolerance = float(tolerance)
        p.max_cycles = int(max_cycles)
        p.number_of_steps = int(number_of_steps)
        p.verbose = int(bool(verbose))
        p.gs_colors = int(gs_colors)
        p.device = int(device)
        p.rank, p.world_size = self.rank, self.world_size
        self.params = p
        if self.world_size > 1:
            if nccl_id is None or len(nccl_id) != 128:
                raise MadGpuError("world_size > 1 needs the 128-byte NCCL unique id (slabs.create_unique_id)")
            self._nccl_id = C.create_string_buffer(bytes(nccl_id), 128)
            rc = self._lib.madgpu_create_slab(C.byref(p), self._nccl_id, C.byref(self._ctx))
        else:
            rc = self._lib.madgpu_create(C.byref(p), C.byref(self._ctx))
        if rc != 0:
            msg = self._lib.madgpu_last_error(None)
            self._ctx = C.c_void_p()
            raise MadGpuError(f"madgpu_create failed ({rc}): {msg.decode() if msg else ''}")
        self.nlevels = self._lib.madgpu_num_levels(self._ctx)
        self.levels = []
        for l in range(self.nlevels):
            n, h, c = (C.c_int32 * 3)(), (C.c_double * 3)(), (C.c_int32 * 3)()
            self._lib.madgpu_level_info(self._ctx, l, n, h, c)
            self.levels.append(dict(n=tuple(n)[: self.dim], h=tuple(h)[: self.dim], centering=tuple(c)[: self.dim],
                                    shape=tuple(n)[: self.dim][::-1]))
        if self.world_size > 1:
            zb, zc, gz = C.c_int32(), C.c_int32(), C.c_int32()
            self._check(self._lib.madgpu_slab(self._ctx, 0, C.byref(zb), C.byref(zc), C.byref(gz)), "slab")
            self.z_begin, self.z_count = zb.value, zc.value
            self.shape = (zc.value,) + self.global_shape[1:]
        else:
            self.z_begin, self.z_count = 0, self.global_shape[0] if self.dim == 3 else 1
        self.ncomp = 3 if self.dim == 2 else 6
        self.ns = 9 if self.dim == 2 else 27
        self.last_stats = None

    # ------------------------------------------------------------------ lifetime / errors
    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._lib.madgpu_destroy(self._ctx)
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
            msg = self._lib.madgpu_last_error(self._ctx)
            raise MadGpuError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    # ------------------------------------------------------------------ configuration
    def set_solver(self, smoother=None, omega=None, iterations_per_grid=None, cycle=None, tolerance=None,
                   max_cycles=None, number_of_steps=None, verbose=None):
        p = self.params
        if smoother is not None: p.smoother = int(smoother)
        if omega is not None: p.omega = float(omega)
        if iterations_per_grid is not None: p.iterations_per_grid = int(iterations_per_grid)
        if cycle is not None: p.cycle = int(cycle)
        if tolerance is not None: p.tolerance = float(tolerance)
        if max_cycles is not None: p.max_cycles = int(max_cycles)
        if number_of_steps is not None: p.number_of_steps = int(number_of_steps)
        if verbose is not None: p.verbose = int(bool(verbose))
        self._check(self._lib.madgpu_set_solver(self._ctx, p.smoother, p.omega, p.iterations_per_grid, p.cycle,
                                                p.tolerance, p.max_cycles, p.number_of_steps, p.verbose), "set_solver")

    def set_profiling(self, on=True):
        """on: True / False, or a list of kernel-class names (see _lib.K_NAMES) to time with CUDA events."""
        if isinstance(on, (list, tuple)):
            mask = 0
            for k in on:
                mask |= 1 << B.K_NAMES.index(k)
        else:
            mask = -1 if on else 0
        self._check(self._lib.madgpu_set_profiling(self._ctx, mask), "set_profiling")

    def set_tensor(self, tensor_aos: np.ndarray):
        """ITK tensor buffer: shape image.shape + (ncomp,), float32 or float64."""
        t = np.ascontiguousarray(tensor_aos)
        if t.shape != self.shape + (self.ncomp,):
            raise MadGpuError(f"tensor shape {t.shape} != {self.shape + (self.ncomp,)}")
        if t.dtype == np.float32:
            self._check(self._lib.madgpu_set_tensor_f32(self._ctx, _ptr(t)), "set_tensor_f32")
        elif t.dtype == np.float64:
            self._check(self._lib.madgpu_set_tensor_f64(self._ctx, _ptr(t)), "set_tensor_f64")
        else:
            raise MadGpuError("tensor dtype must be float32 or float64")

    def set_tensor_device(self, plane_ptrs):
        """ncomp device pointers (ints) to dense fp32 planes already resident in HBM."""
        arr = (C.c_void_p * self.ncomp)(*[C.c_void_p(int(p)) for p in plane_ptrs])
        self._check(self._lib.madgpu_set_tensor_device_f32(self._ctx, arr), "set_tensor_device_f32")

    # ------------------------------------------------------------------ solve
    def _stats(self, st: B.Stats):
        d = dict(steps=st.steps, cycles_per_step=list(st.cycles_per_step)[: st.steps],
                 final_relres=list(st.final_relres)[: st.steps], total_cycles=st.total_cycles, levels=st.levels,
                 setup_ms=st.setup_ms, h2d_ms=st.h2d_ms, d2h_ms=st.d2h_ms, solve_ms=st.solve_ms, fmg_ms=st.fmg_ms,
                 kernel_launches=st.kernel_launches,
                 prof_ms={k: st.prof_ms[i] for i, k in enumerate(B.K_NAMES)},
                 prof_launches={k: st.prof_launches[i] for i, k in enumerate(B.K_NAMES)}, graph_launches=int(st.graph_launches))
        self.last_stats = d
        return d

    def solve(self, image: np.ndarray, out_dtype=None, out: np.ndarray = None) -> np.ndarray:
        """GenerateData(): host image in, host image out (pixel type preserved unless out_dtype).
        `out` may be a preallocated (e.g. pinned) C-contiguous array of the output type."""
        img = np.ascontiguousarray(image)
        if img.shape != self.shape:
            raise MadGpuError(f"image shape {img.shape} != {self.shape}")
        if img.dtype not in _PIX:
            raise MadGpuError(f"unsupported pixel type {img.dtype}")
        odt = np.dtype(out_dtype) if out_dtype is not None else (out.dtype if out is not None else img.dtype)
        if out is None:
            out = np.empty(self.shape, dtype=odt)
        elif out.shape != self.shape or out.dtype != odt or not out.flags.c_contiguous:
            raise MadGpuError("out must be C-contiguous with the image shape and the output pixel type")
        st = B.Stats()
        st.struct_size = C.sizeof(B.Stats)
        self._check(self._lib.madgpu_solve_cast(self._ctx, _PIX[img.dtype], _ptr(img), _PIX[odt], _ptr(out), C.byref(st)),
                    "solve")
        self._stats(st)
        return out

    def solve_device(self, d_in, d_out: int):
        """Device-resident dense fp32 image in / out (raw device pointers).  d_in = None continues from the fp64 result of the previous
        solve of this context (the carrier between the VED filter's outer iterations)."""
        st = B.Stats()
        st.struct_size = C.sizeof(B.Stats)
        self._check(self._lib.madgpu_solve_device_f32(self._ctx, C.c_void_p(int(d_in)) if d_in is not None else None, C.c_void_p(int(d_out)), C.byref(st)),
                    "solve_device")
        return self._stats(st)

    # cycle-level driving (benchmarks, per-V-cycle parity)
    def cycles_begin(self, image=None, d_in=None):
        if d_in is not None:
            self._check(self._lib.madgpu_cycles_begin_device_f32(self._ctx, C.c_void_p(int(d_in))), "cycles_begin_device")
        else:
            img = np.ascontiguousarray(image, dtype=np.float32)
            if img.shape != self.shape:
                raise MadGpuError(f"image shape {img.shape} != {self.shape}")
            self._check(self._lib.madgpu_cycles_begin_f32(self._ctx, _ptr(img)), "cycles_begin")

    def cycles_run(self, n):
        """n outer iterations; returns (relres[n], device_ms, stats)."""
        rr = np.empty(max(n, 1), dtype=np.float64)
        ms = C.c_float()
        st = B.Stats()
        st.struct_size = C.sizeof(B.Stats)
        self._check(self._lib.madgpu_cycles_run(self._ctx, int(n), rr.ctypes.data_as(C.POINTER(C.c_double)), C.byref(ms),
                                                C.byref(st)), "cycles_run")
        return rr[:n], float(ms.value), self._stats(st)

    def cycles_end(self, d_out=None):
        if d_out is not None:
            self._check(self._lib.madgpu_cycles_end_device_f32(self._ctx, C.c_void_p(int(d_out))), "cycles_end_device")
            return None
        out = np.empty(self.shape, dtype=np.float64)
        self._check(self._lib.madgpu_cycles_end_f64(self._ctx, _ptr(out)), "cycles_end")
        return out

    # peer-memory halo of a z-slab context (see slabs.enable_peer_halo)
    def ipc_export(self) -> bytes:
        n = C.c_size_t()
        self._check(self._lib.madgpu_ipc_export(self._ctx, None, 0, C.byref(n)), "ipc_export")
        buf = C.create_string_buffer(n.value)
        self._check(self._lib.madgpu_ipc_export(self._ctx, buf, n.value, C.byref(n)), "ipc_export")
        return buf.raw

    def ipc_import(self, blob_lower, blob_upper):
        self._check(self._lib.madgpu_ipc_import(self._ctx, blob_lower, blob_upper), "ipc_import")

    def ipc_disable(self):
        self._check(self._lib.madgpu_ipc_disable(self._ctx), "ipc_disable")

    def gs_tile(self, level=0):
        """(tx, ty, tz) of the fused Gauss-Seidel sweep on `level`, or None for one pass per colour."""
        t = (C.c_int32 * 3)()
        self._check(self._lib.madgpu_gs_tile(self._ctx, int(level), t), "gs_tile")
        return None if t[0] == 0 else (t[0], t[1], t[2])

    def gs_leg_plan(self, level=0, n_iter=1):
        """The passes the next n_iter Gauss-Seidel sweeps on `level` run as: list of dicts(fused, tile=(tx, ty, tz), shift=(oy, oz));
        a tile of (0, 0, 0) is one pass per colour over the whole level (madgpu_gs_leg_plan)."""
        buf = (C.c_int32 * (6 * 64))()
        n = self._lib.madgpu_gs_leg_plan(self._ctx, int(level), int(n_iter), buf, 64)
        if n < 0:
            self._check(n, "gs_leg_plan")
        return [dict(fused=buf[6 * i], tile=(buf[6 * i + 1], buf[6 * i + 2], buf[6 * i + 3]), shift=(buf[6 * i + 4], buf[6 * i + 5])) for i in range(n)]

    def relres_history(self) -> np.ndarray:
        p = self.params
        h = np.full(p.number_of_steps * p.max_cycles, np.nan)
        self._lib.madgpu_get_relres_history(self._ctx, h.ctypes.data_as(C.POINTER(C.c_double)), h.size)
        return h.reshape(p.number_of_steps, p.max_cycles)

    # ------------------------------------------------------------------ per-operator entry points (tests)
    def _f32(self, a, level):
        a = np.ascontiguousarray(a, dtype=np.float32)
        if a.shape != self.levels[level]["shape"]:
            raise MadGpuError(f"array shape {a.shape} != level {level} shape {self.levels[level]['shape']}")
        return a

    def op_get_tensor(self, level):
        out = np.empty((self.ncomp,) + self.levels[level]["shape"], dtype=np.float32)
        self._check(self._lib.madgpu_op_get_tensor(self._ctx, level, _ptr(out)), "op_get_tensor")
        return out

    def op_assemble(self, level):
        out = np.empty(self.levels[level]["shape"] + (self.ns,), dtype=np.float32)
        self._check(self._lib.madgpu_op_assemble(self._ctx, level, _ptr(out)), "op_assemble")
        return out

    def op_smooth(self, level, u, f, smoother=None, n_iter=1):
        u, f = self._f32(u, level), self._f32(f, level)
        out = np.empty_like(u)
        sm = self.params.smoother if smoother is None else int(smoother)
        self._check(self._lib.madgpu_op_smooth(self._ctx, level, sm, n_iter, _ptr(u), _ptr(f), _ptr(out)), "op_smooth")
        return out

    def op_residual(self, level, u, f):
        u, f = self._f32(u, level), self._f32(f, level)
        out = np.empty_like(u)
        nrm = C.c_double()
        self._check(self._lib.madgpu_op_residual(self._ctx, level, _ptr(u), _ptr(f), _ptr(out), C.byref(nrm)), "op_residual")
        return out, nrm.value

    def op_residual_f64(self, u, f, norm_only=False):
        """norm_only: run the kernel of the solve loop (fp32 residual kept on the device, fp64 norm returned)."""
        u = np.ascontiguousarray(u, dtype=np.float64)
        f = np.ascontiguousarray(f, dtype=np.float64)
        assert u.shape == self.shape and f.shape == self.shape
        out = None if norm_only else np.empty_like(u)
        nrm = C.c_double()
        self._check(self._lib.madgpu_op_residual_f64(self._ctx, _ptr(u), _ptr(f), _ptr(out) if out is not None else None,
                                                     C.byref(nrm)), "op_residual_f64")
        return out, nrm.value

    def op_restrict(self, fine_level, fine):
        fine = self._f32(fine, fine_level)
        out = np.empty(self.levels[fine_level + 1]["shape"], dtype=np.float32)
        self._check(self._lib.madgpu_op_restrict(self._ctx, fine_level, _ptr(fine), _ptr(out)), "op_restrict")
        return out

    def op_prolong(self, fine_level, coarse):
        coarse = self._f32(coarse, fine_level + 1)
        out = np.empty(self.levels[fine_level]["shape"], dtype=np.float32)
        self._check(self._lib.madgpu_op_prolong(self._ctx, fine_level, _ptr(coarse), _ptr(out)), "op_prolong")
        return out

    def op_coarse_solve(self, f):
        f = self._f32(f, self.nlevels - 1)
        out = np.empty_like(f)
        self._check(self._lib.madgpu_op_coarse_solve(self._ctx, _ptr(f), _ptr(out)), "op_coarse_solve")
        return out

    def op_vcycle(self, level, u, f):
        u, f = self._f32(u, level), self._f32(f, level)
        out = np.empty_like(u)
        self._check(self._lib.madgpu_op_vcycle(self._ctx, level, _ptr(u), _ptr(f), _ptr(out)), "op_vcycle")
        return out
