set(DOCUMENTATION "B200 (sm_100a) drop-in for the multigrid anisotropic-diffusion module: the filters keep their names and
template signatures, GenerateData() runs in libmadgpu (CUDA) through the C-ABI of include/madgpu.h / include/madved.h.")

# Same dependencies as the reference's itk-module.cmake minus ITKImageFeature: the Hessian of the VED filter is computed by
# libmadgpu's own recursive-Gaussian kernels instead of itk::HessianRecursiveGaussianImageFilter.
itk_module(MultigridAnisotropicDiffusion
  DEPENDS
    ITKCommon
    ITKIOImageBase
    ITKImageFilterBase
    ITKImageGrid
  TEST_DEPENDS
    ITKTestKernel
  EXCLUDE_FROM_DEFAULT
  DESCRIPTION
    "${DOCUMENTATION}"
)
