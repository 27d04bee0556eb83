# ITK remote-module declaration of the B200 drop-in (read by ITK's module system when this directory is placed under
# Modules/Remote or Modules/External; the standalone build in CMakeLists.txt does not use it).
#
# Dependencies: what the drop-in headers touch of ITK -- images, regions, iterators, ImageToImageFilter, image I/O for the tests.
# ITKImageFeature, which the reference needs for itk::HessianRecursiveGaussianImageFilter, is NOT a dependency here: the Hessian of
# the VED filter is computed by libmadgpu's own recursive-Gaussian kernels (include/madved.h).
set(MADGPU_MODULE_DESCRIPTION
    "Multigrid anisotropic diffusion and vessel enhancing diffusion on NVIDIA B200 (sm_100a): the filters keep their names and template signatures, GenerateData() runs in libmadgpu through the C-ABI of include/madgpu.h and include/madved.h.")

itk_module(MultigridAnisotropicDiffusion
  ENABLE_SHARED
  DEPENDS ITKCommon ITKImageFilterBase ITKImageGrid ITKIOImageBase
  TEST_DEPENDS ITKTestKernel
  DESCRIPTION "${MADGPU_MODULE_DESCRIPTION}"
  EXCLUDE_FROM_DEFAULT)
