/*
 * itkVEDMultigridImageFilter.h -- B200 drop-in for the reference filter of the same name
 * (/root/reference/include/itkVEDMultigridImageFilter.h:41-170, .hxx:33-402): Manniesing's vessel enhancing diffusion with the
 * multigrid solver for the diffusion steps.
 *
 * Same class name, template signature, setters and defaults (.hxx:33-58).  GenerateData() (.hxx:63-155) no longer runs on the
 * host: the input buffer goes to libmadgpu.so once, and per outer iteration the Hessian at every scale (ComputeHessian,
 * .hxx:158-173), the vesselness update (.hxx:215-299), the tensor (.hxx:302-378) and DiffusionStep (.hxx:381-402: NumberOfSteps =
 * DiffusionIterations, IterationsPerGrid = DiffusionIterationsPerGrid, MaxCycles = 100) all run on the device through
 * madved_run (include/madved.h); the tensor never leaves HBM.  Output: static_cast to the output pixel type, spacing and origin
 * copied, direction not propagated (as in the reference, .hxx:132-137).
 *
 * The Hessian is computed by the library's own recursive-Gaussian kernels (the reference calls
 * itk::HessianRecursiveGaussianImageFilter, third-party; parity with ITK's filter is unpinned, see include/madved.h).  In an
 * ITK tree a caller who wants ITK's Hessian can drive madved_update_vesselness_host_f64 directly.
 * Header-only; link with -lmadgpu.
 */
#ifndef __itkVEDMultigridImageFilter_h
#define __itkVEDMultigridImageFilter_h

#include <vector>

#include "itkImageToImageFilter.h"
#include "itkMacro.h"
#include "itkMultigridAnisotropicDiffusionImageFilter.h"
#include "madved.h"

namespace itk
{
template <class TInputImage, class TOutputImage, class TSmootherType = mad::MultigridGaussSeidelSmoother<TInputImage::ImageDimension> >
class VEDMultigridImageFilter
  : public ImageToImageFilter<Image<typename TInputImage::PixelType, 3>, Image<typename TOutputImage::PixelType, 3> >
{
public:
  typedef VEDMultigridImageFilter Self;
  typedef ImageToImageFilter<TInputImage, TOutputImage> SuperClass;
  typedef SmartPointer<Self> Pointer;
  typedef SmartPointer<const Self> ConstPointer;
  typedef TInputImage InputImageType;
  typedef typename TInputImage::PixelType InputPixelType;
  typedef TOutputImage OutputImageType;
  typedef typename TOutputImage::PixelType OutputPixelType;
  typedef double InternalPixelType;
  typedef Image<InternalPixelType, 3> InternalImageType;
  typedef InternalPixelType Precision;
  typedef MultigridAnisotropicDiffusionImageFilter<InternalImageType, InternalImageType, TSmootherType> MADFilterType;
  typedef typename MADFilterType::CycleType CycleType;

  itkNewMacro(Self);
  itkTypeMacro(VEDMultigridImageFilter, ImageToImageFilter);

  /** VED parameters (reference .h:88-96). */
  itkSetMacro(Alpha, Precision);
  itkSetMacro(Beta, Precision);
  itkSetMacro(Gamma, Precision);
  itkSetMacro(Epsilon, Precision);
  itkSetMacro(Omega, Precision);
  itkSetMacro(Sensitivity, Precision);
  itkSetMacro(Scales, std::vector<Precision>);
  itkSetMacro(Iterations, unsigned int);
  itkSetMacro(DiffusionIterations, unsigned int);
  /** MAD parameters (reference .h:99-102). */
  itkSetMacro(Cycle, CycleType);
  itkSetMacro(TimeStep, Precision);
  itkSetMacro(Tolerance, Precision);
  itkSetMacro(DiffusionIterationsPerGrid, unsigned int);
  itkSetMacro(Verbose, bool);
  /** CUDA device ordinal (new; default 0). */
  itkSetMacro(Device, int);

  /** Statistics of the last Update(): the last DiffusionStep (madgpu.h) and the tensor front-end (madved.h). */
  const madgpu_stats& GetStatistics() const { return m_Stats; }
  const madved_stats& GetFrontEndStatistics() const { return m_VedStats; }

protected:
  VEDMultigridImageFilter()
    : m_Alpha(0.5), m_Beta(0.5), m_Gamma(5.0), m_Epsilon(0.01), m_Omega(5.), m_Sensitivity(10.), m_Iterations(1), m_DiffusionIterations(5),
      m_Cycle(MADFilterType::VCYCLE), m_TimeStep(0.1), m_Tolerance(1e-6), m_DiffusionIterationsPerGrid(2), m_Verbose(false), m_Device(0)
  {
    m_Scales.resize(5);
    m_Scales[0] = 0.300;
    m_Scales[1] = 0.482;
    m_Scales[2] = 0.775;
    m_Scales[3] = 1.245;
    m_Scales[4] = 2.000;
    m_Stats = madgpu_stats();
    m_VedStats = madved_stats();
  }
  ~VEDMultigridImageFilter() {}

  virtual void GenerateData()
  {
    typedef Image<InputPixelType, 3> InImage;
    typedef Image<OutputPixelType, 3> OutImage;
    const InImage* input = this->GetInput();
    if (!input) itkExceptionMacro(<< "no input image");
    if (m_Scales.empty()) itkExceptionMacro(<< "no scales");
    const typename InImage::RegionType region = input->GetLargestPossibleRegion();

    madved_params vp;
    madved_params_default(&vp);
    madgpu_params sp;
    madgpu_params_default(&sp);
    sp.dim = 3;
    for (unsigned int d = 0; d < 3; ++d) {
      vp.size[d] = sp.size[d] = static_cast<int32_t>(region.GetSize(d));
      vp.spacing[d] = sp.spacing[d] = input->GetSpacing()[d];
    }
    vp.alpha = m_Alpha; vp.beta = m_Beta; vp.gamma = m_Gamma;
    vp.epsilon = m_Epsilon; vp.omega = m_Omega; vp.sensitivity = m_Sensitivity;
    vp.device = sp.device = m_Device;
    // DiffusionStep's parameter mapping (reference .hxx:386-397)
    sp.verbose = m_Verbose ? 1 : 0;
    sp.time_step = m_TimeStep;
    sp.tolerance = m_Tolerance;
    sp.number_of_steps = static_cast<int32_t>(m_DiffusionIterations);
    sp.iterations_per_grid = static_cast<int32_t>(m_DiffusionIterationsPerGrid);
    sp.cycle = static_cast<int32_t>(m_Cycle);
    sp.max_cycles = 100;
    sp.smoother = TSmootherType::MadgpuSmoother();

    struct Guard {
      madgpu_ctx* s;
      madved_ctx* v;
      ~Guard() { madved_destroy(v); madgpu_destroy(s); }
    } g = {nullptr, nullptr};
    if (madgpu_create(&sp, &g.s) != MADGPU_OK) itkExceptionMacro(<< "madgpu_create: " << madgpu_last_error(nullptr));
    if (madved_create(&vp, &g.v) != MADGPU_OK) itkExceptionMacro(<< "madved_create: " << madved_last_error(nullptr));

    typename OutImage::Pointer outputImage = OutImage::New();
    outputImage->SetRegions(region);
    outputImage->Allocate();
    outputImage->SetSpacing(input->GetSpacing());
    outputImage->SetOrigin(input->GetOrigin());

    m_Stats.struct_size = static_cast<int32_t>(sizeof(madgpu_stats));
    const int rc = madved_run(g.v, g.s, madgpu_detail::PixelTag<InputPixelType>::value, input->GetBufferPointer(),
                              madgpu_detail::PixelTag<OutputPixelType>::value, outputImage->GetBufferPointer(), m_Scales.data(),
                              static_cast<int32_t>(m_Scales.size()), static_cast<int32_t>(m_Iterations), &m_Stats);
    if (rc != MADGPU_OK) itkExceptionMacro(<< "madved_run: " << madved_last_error(g.v));
    m_VedStats.struct_size = static_cast<int32_t>(sizeof(madved_stats));
    madved_get_stats(g.v, &m_VedStats);

    this->AllocateOutputs();
    this->GraftOutput(outputImage);
  }

private:
  Precision m_Alpha, m_Beta, m_Gamma, m_Epsilon, m_Omega, m_Sensitivity;
  std::vector<Precision> m_Scales;
  unsigned int m_Iterations, m_DiffusionIterations;
  CycleType m_Cycle;
  Precision m_TimeStep, m_Tolerance;
  unsigned int m_DiffusionIterationsPerGrid;
  bool m_Verbose;
  int m_Device;
  madgpu_stats m_Stats;
  madved_stats m_VedStats;

  VEDMultigridImageFilter(const Self&);
  void operator=(const Self&);
};
}  // namespace itk
#endif
