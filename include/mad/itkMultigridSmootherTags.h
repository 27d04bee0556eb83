/*
 * mad/itkMultigridSmootherTags.h -- the reference selects the smoother with a template argument
 * (TSmootherType of itk::MultigridAnisotropicDiffusionImageFilter, default
 * mad::MultigridGaussSeidelSmoother<Dim>; /root/reference/include/itkMultigridAnisotropicDiffusionImageFilter.h:89-92)
 * and default-constructs it.  In the B200 drop-in the smoothers live in libmadgpu.so, so the two class names
 * survive as tag types carrying the run-time selector of include/madgpu.h.
 *
 *   mad::MultigridGaussSeidelSmoother<D>      replaces mad/itkMultigridGaussSeidelSmoother.h (lexicographic sweep ->
 *                                             fused / multicolour Gauss-Seidel on the GPU, same fixed point)
 *   mad::MultigridWeightedJacobiSmoother<D>   replaces mad/itkMultigridWeightedJacobiSmoother.h (omega = 2/3 when
 *                                             default-constructed, .hxx:186-191)
 */
#ifndef __itkMultigridSmootherTags_h
#define __itkMultigridSmootherTags_h

#include "madgpu.h"

namespace itk
{
namespace mad
{
template <unsigned int VDimension>
struct MultigridGaussSeidelSmoother {
  static int MadgpuSmoother() { return MADGPU_SMOOTHER_GS; }
  static double Omega() { return 1.0; }
};

template <unsigned int VDimension>
struct MultigridWeightedJacobiSmoother {
  static int MadgpuSmoother() { return MADGPU_SMOOTHER_WJ; }
  static double Omega() { return 2.0 / 3.0; }
};
}  // namespace mad
}  // namespace itk
#endif
