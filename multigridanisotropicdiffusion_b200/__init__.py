"""B200-native multigrid solver for the implicit anisotropic-diffusion step of
itk::MultigridAnisotropicDiffusionImageFilter / itk::VEDMultigridImageFilter.

The product is libmadgpu.so (hand-written sm_100a CUDA behind the C-ABI of include/madgpu.h);
this package is its host-side mirror of the reference's filter interface.  Importing the package
does not load CUDA; the first solver does, and it raises if the library or the GPU is missing.
"""
from .filter import (MultigridAnisotropicDiffusionImageFilter, MultigridGaussSeidelSmoother,
                     MultigridWeightedJacobiSmoother, VEDMultigridImageFilter)
from .solver import MadGpuError, MadSolver
from .ved import MadVed

__all__ = ["MultigridAnisotropicDiffusionImageFilter", "VEDMultigridImageFilter", "MultigridGaussSeidelSmoother",
           "MultigridWeightedJacobiSmoother", "MadSolver", "MadVed", "MadGpuError"]
