"""Host-side mirror of the reference's filter API for the multigrid diffusion path.

Same names, argument meaning and defaults as itk::MultigridAnisotropicDiffusionImageFilter
(/root/reference/include/itkMultigridAnisotropicDiffusionImageFilter.h:89-171, defaults .hxx:38-49)
and the diffusion half of itk::VEDMultigridImageFilter (DiffusionStep,
/root/reference/include/itkVEDMultigridImageFilter.hxx:381-402), so the parity tests read like
the reference's own test programs (test/itk2DDiffusionTest_WJ.cxx:88-109).  The C++ drop-in for an
ITK tree is include/itkMultigridAnisotropicDiffusionImageFilter.h; both sit on the same C-ABI.

The smoother is a constructor argument here because the reference selects it with a template
argument (`TSmootherType`); "gs" (default, as in the reference) or "wj".
"""
from __future__ import annotations

import numpy as np

from .solver import MadGpuError, MadSolver


class MultigridGaussSeidelSmoother:  # tag types, mad/itkMultigridGaussSeidelSmoother.h
    tag = MadSolver.GS


class MultigridWeightedJacobiSmoother:  # mad/itkMultigridWeightedJacobiSmoother.h
    tag = MadSolver.WJ


def _smoother_tag(s):
    if s is None:
        return MadSolver.GS
    if isinstance(s, str):
        return {"gs": MadSolver.GS, "wj": MadSolver.WJ}[s.lower()]
    if hasattr(s, "tag"):
        return s.tag
    return int(s)


class MultigridAnisotropicDiffusionImageFilter:
    VCYCLE, FMG, SMOOTHER = MadSolver.VCYCLE, MadSolver.FMG, MadSolver.SMOOTHER

    def __init__(self, smoother=None, device=0):
        self._smoother = _smoother_tag(smoother)
        self._device = device
        # defaults: …Filter.hxx:38-49
        self._time_step = 0.01
        self._number_of_steps = 1
        self._cycle = self.VCYCLE
        self._iterations_per_grid = 2
        self._tolerance = 1e-6
        self._max_cycles = 100
        self._verbose = False
        self._input = None
        self._spacing = None
        self._tensor = None
        self._output = None
        self._solver = None
        self._solver_key = None
        self._tensor_dirty = True
        self.stats = None

    # itkSetMacro setters (…Filter.h:133-156)
    def SetCycle(self, c): self._cycle = int(c)
    def SetIterationsPerGrid(self, n): self._iterations_per_grid = int(n)
    def SetMaxCycles(self, n): self._max_cycles = int(n)
    def SetNumberOfSteps(self, n): self._number_of_steps = int(n)
    def SetTimeStep(self, dt): self._time_step = float(dt)
    def SetTolerance(self, t): self._tolerance = float(t)
    def SetVerbose(self, v): self._verbose = bool(v)

    def SetDiffusionTensor(self, tensor_aos):
        """…Filter.h:160 -- ITK tensor buffer, shape image.shape + (3|6,); deep-copied like the reference."""
        t = np.array(tensor_aos, copy=True)
        if t.dtype not in (np.float32, np.float64):
            t = t.astype(np.float64)
        self._tensor = t
        self._tensor_dirty = True

    def SetInput(self, image, spacing=None):
        """image: numpy array (uint8 / int16 / float32 / float64); spacing: (sx, sy[, sz]), default 1."""
        self._input = np.asarray(image)
        self._spacing = tuple(spacing) if spacing is not None else (1.0,) * self._input.ndim

    def Update(self):
        if self._input is None:
            raise MadGpuError("no input image")
        if self._tensor is None:
            raise MadGpuError("no diffusion tensor (SetDiffusionTensor)")
        key = (self._input.shape, self._spacing, self._time_step, self._device)
        if self._solver is None or key != self._solver_key:
            if self._solver is not None:
                self._solver.close()
            self._solver = MadSolver(self._input.shape, self._spacing, time_step=self._time_step, device=self._device)
            self._solver_key = key
            self._tensor_dirty = True
        s = self._solver
        s.set_solver(smoother=self._smoother, iterations_per_grid=self._iterations_per_grid, cycle=self._cycle,
                     tolerance=self._tolerance, max_cycles=self._max_cycles, number_of_steps=self._number_of_steps,
                     verbose=self._verbose)
        if self._tensor_dirty:
            s.set_tensor(self._tensor)
            self._tensor_dirty = False
        self._output = s.solve(self._input)
        self.stats = s.last_stats
        return self

    def GetOutput(self):
        if self._output is None:
            self.Update()
        return self._output

    def close(self):
        if self._solver is not None:
            self._solver.close()
            self._solver = None


class VEDMultigridImageFilter:
    """itk::VEDMultigridImageFilter (/root/reference/include/itkVEDMultigridImageFilter.h:41-170): same setters, defaults
    (VED.hxx:33-58) and GenerateData flow (VED.hxx:63-155) -- per outer iteration the Hessian at every scale, vesselness,
    tensor, then DiffusionStep -- all of it on the device (include/madved.h).  A tensor handed in with SetDiffusionTensor
    (not part of the reference's API) replaces the front-end, which is how the solve path alone is driven."""

    VCYCLE, FMG, SMOOTHER = MadSolver.VCYCLE, MadSolver.FMG, MadSolver.SMOOTHER

    def __init__(self, smoother=None, device=0):
        self._smoother = smoother
        self._device = device
        # defaults: VED.hxx:33-58
        self._alpha, self._beta, self._gamma = 0.5, 0.5, 5.0
        self._epsilon, self._omega, self._sensitivity = 0.01, 5.0, 10.0
        self._scales = [0.300, 0.482, 0.775, 1.245, 2.000]
        self._iterations = 1
        self._diffusion_iterations = 5
        self._cycle = self.VCYCLE
        self._time_step = 0.1
        self._tolerance = 1e-6
        self._diffusion_iterations_per_grid = 2
        self._verbose = False
        self._tensor = None
        self._input = None
        self._spacing = None
        self._output = None
        self.stats = None
        self.ved_stats = None

    # VED parameters (VED.h:88-96)
    def SetAlpha(self, v): self._alpha = float(v)
    def SetBeta(self, v): self._beta = float(v)
    def SetGamma(self, v): self._gamma = float(v)
    def SetEpsilon(self, v): self._epsilon = float(v)
    def SetOmega(self, v): self._omega = float(v)
    def SetSensitivity(self, v): self._sensitivity = float(v)
    def SetScales(self, v): self._scales = [float(s) for s in v]
    def SetIterations(self, n): self._iterations = int(n)
    def SetDiffusionIterations(self, n): self._diffusion_iterations = int(n)
    # MAD parameters (VED.h:99-106)
    def SetCycle(self, c): self._cycle = int(c)
    def SetTimeStep(self, dt): self._time_step = float(dt)
    def SetTolerance(self, t): self._tolerance = float(t)
    def SetDiffusionIterationsPerGrid(self, n): self._diffusion_iterations_per_grid = int(n)
    def SetVerbose(self, v): self._verbose = bool(v)
    def SetDiffusionTensor(self, t): self._tensor = t

    def SetInput(self, image, spacing=None):
        self._input = np.asarray(image)
        self._spacing = tuple(spacing) if spacing is not None else (1.0,) * self._input.ndim

    def DiffusionStep(self, image):
        """VED.hxx:381-402 with a host tensor (SetDiffusionTensor)."""
        f = MultigridAnisotropicDiffusionImageFilter(self._smoother, self._device)
        f.SetVerbose(self._verbose)
        f.SetDiffusionTensor(self._tensor)
        f.SetInput(image, self._spacing)
        f.SetTimeStep(self._time_step)
        f.SetTolerance(self._tolerance)
        f.SetNumberOfSteps(self._diffusion_iterations)
        f.SetIterationsPerGrid(self._diffusion_iterations_per_grid)
        f.SetCycle(self._cycle)
        f.SetMaxCycles(100)
        f.Update()
        self.stats = f.stats
        out = f.GetOutput()
        f.close()
        return out

    def Update(self):
        if self._input is None:
            raise MadGpuError("no input image")
        if self._tensor is not None:
            self._output = self.DiffusionStep(self._input)
            return self
        from .ved import MadVed
        shape = self._input.shape
        # DiffusionStep's parameter mapping (VED.hxx:386-397)
        import time
        t0 = time.perf_counter()
        solver = MadSolver(shape, self._spacing, time_step=self._time_step, smoother=_smoother_tag(self._smoother),
                           iterations_per_grid=self._diffusion_iterations_per_grid, cycle=self._cycle, tolerance=self._tolerance,
                           max_cycles=100, number_of_steps=self._diffusion_iterations, verbose=self._verbose, device=self._device)
        t1 = time.perf_counter()
        try:
            ved = MadVed(shape, self._spacing, self._alpha, self._beta, self._gamma, self._epsilon, self._omega, self._sensitivity,
                         device=self._device)
            t2 = time.perf_counter()
            try:
                self._output = ved.run(solver, self._input, self._scales, self._iterations)
                # wall clock of the three parts of the call: the two contexts (device allocations) and madved_run
                self.timing = {"solver_create_s": t1 - t0, "ved_create_s": t2 - t1, "run_s": time.perf_counter() - t2}
                self.ved_stats = ved.stats()
                self.stats = solver.last_stats
            finally:
                ved.close()
        finally:
            solver.close()
        return self

    def GetOutput(self):
        if self._output is None:
            self.Update()
        return self._output

    def close(self):
        """Nothing to release: Update() creates and destroys its device contexts (kept for symmetry with the solver filter)."""

