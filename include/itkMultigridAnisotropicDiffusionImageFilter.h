/*
 * itkMultigridAnisotropicDiffusionImageFilter.h -- B200 drop-in for the reference filter of the same name
 * (/root/reference/include/itkMultigridAnisotropicDiffusionImageFilter.h:89-171, .hxx:38-297).
 *
 * Same class name, template signature, setters, defaults and output semantics; GenerateData() no longer runs the
 * multigrid on the host but hands the raw ITK buffers to libmadgpu.so through the C-ABI of madgpu.h:
 *
 *   reference                                              this header
 *   -----------------------------------------------------  ---------------------------------------------
 *   SetDiffusionTensor(): deep copy + cast to double         keeps the pointer; the buffer goes to
 *     (.hxx:66-101)                                          madgpu_set_tensor_f32/_f64 as is (AoS, ITK order)
 *   GenerateData(): cast, GridsHierarchy, DirectSolver,      madgpu_create + madgpu_set_tensor_* + madgpu_solve_cast
 *     time-step loop, V-cycle/FMG/smoother, cast (.hxx:104-297)
 *   output: static_cast to the pixel type, spacing + origin  same (direction is not propagated, as in the reference)
 *     copied (.hxx:267-289)
 *   errors: none raised                                      non-zero C-ABI codes become itkExceptionMacro
 *
 * Pixel types: unsigned char, short, float, double (MADGPU_PIX_*); the tensor image has the input pixel type
 * (…Filter.h:111-112); integer tensors are converted to double on the host first (they are tiny in practice:
 * the reference's users pass float or double tensors).  Header-only; link with -lmadgpu.
 */
#ifndef __itkMultigridAnisotropicDiffusionImageFilter_h
#define __itkMultigridAnisotropicDiffusionImageFilter_h

#include <string>
#include <vector>

#include "itkImage.h"
#include "itkImageToImageFilter.h"
#include "itkMacro.h"
#include "itkSymmetricSecondRankTensor.h"
#include "mad/itkMultigridSmootherTags.h"
#include "madgpu.h"

#ifndef itkExceptionMacro  /* stand-in ITK used by the compile test has no exception machinery */
#include <sstream>
#include <stdexcept>
#define itkExceptionMacro(x) { std::ostringstream m_; m_ << "itk::ERROR: " x; throw std::runtime_error(m_.str()); }
#define MADGPU_PLAIN_EXCEPTION 1
#endif

namespace itk
{
namespace madgpu_detail
{
template <typename T> struct PixelTag;
template <> struct PixelTag<unsigned char> { enum { value = MADGPU_PIX_U8 }; };
template <> struct PixelTag<short> { enum { value = MADGPU_PIX_I16 }; };
template <> struct PixelTag<float> { enum { value = MADGPU_PIX_F32 }; };
template <> struct PixelTag<double> { enum { value = MADGPU_PIX_F64 }; };
}  // namespace madgpu_detail

template <class TInputImage, class TOutputImage, class TSmootherType = mad::MultigridGaussSeidelSmoother<TInputImage::ImageDimension> >
class MultigridAnisotropicDiffusionImageFilter : public ImageToImageFilter<TInputImage, TOutputImage>
{
public:
  typedef MultigridAnisotropicDiffusionImageFilter Self;
  typedef ImageToImageFilter<TInputImage, TOutputImage> SuperClass;
  typedef SmartPointer<Self> Pointer;
  typedef SmartPointer<const Self> ConstPointer;
  typedef TInputImage InputImageType;
  typedef typename TInputImage::PixelType InputPixelType;
  typedef TOutputImage OutputImageType;
  typedef typename TOutputImage::PixelType OutputPixelType;
  typedef double InternalPixelType;
  typedef Image<SymmetricSecondRankTensor<InputPixelType, TInputImage::ImageDimension>, TInputImage::ImageDimension> InputTensorImageType;
  typedef typename InputTensorImageType::PixelType TensorPixelType;
  typedef InternalPixelType Precision;

  enum CycleType { VCYCLE, FMG, SMOOTHER };  // == MADGPU_CYCLE_V / _FMG / _SMOOTHER

  itkNewMacro(Self);
  itkTypeMacro(MultigridAnisotropicDiffusionImageFilter, ImageToImageFilter);

  itkSetMacro(Cycle, CycleType);
  itkSetMacro(IterationsPerGrid, unsigned int);
  itkSetMacro(MaxCycles, unsigned int);
  itkSetMacro(NumberOfSteps, unsigned int);
  itkSetMacro(TimeStep, Precision);
  itkSetMacro(Tolerance, Precision);
  itkSetMacro(Verbose, bool);

  /** Deep copy at Set time, as the reference does (.hxx:66-101): the caller may release, reuse or modify its tensor image
   *  between this call and Update().  Only the pixel buffer is kept; the tensor image's spacing is ignored (.hxx:131). */
  void SetDiffusionTensor(const InputTensorImageType* inputTensor)
  {
    m_TensorPixels = 0;
    m_TensorCopy.clear();
    if (inputTensor) {
      m_TensorPixels = static_cast<size_t>(inputTensor->GetLargestPossibleRegion().GetNumberOfPixels());
      const TensorPixelType* b = inputTensor->GetBufferPointer();
      m_TensorCopy.assign(b, b + m_TensorPixels);
    }
    this->Modified();
  }

  /** CUDA device ordinal (new; default 0). */
  itkSetMacro(Device, int);

  /** Statistics of the last Update(): cycles per time step, final relative residuals, timings (madgpu.h). */
  const madgpu_stats& GetStatistics() const { return m_Stats; }

protected:
  MultigridAnisotropicDiffusionImageFilter()
    : m_TimeStep(0.01), m_NumberOfSteps(1), m_Cycle(VCYCLE), m_IterationsPerGrid(2), m_Tolerance(1e-6), m_MaxCycles(100), m_Verbose(false),
      m_Device(0), m_TensorPixels(0)
  {
    m_Stats = madgpu_stats();
  }
  ~MultigridAnisotropicDiffusionImageFilter() {}

  virtual void GenerateData()
  {
    const unsigned int Dim = TInputImage::ImageDimension;
    const InputImageType* input = this->GetInput();
    if (!input) itkExceptionMacro(<< "no input image");
    if (m_TensorCopy.empty()) itkExceptionMacro(<< "no diffusion tensor (SetDiffusionTensor)");
    const typename InputImageType::RegionType region = input->GetLargestPossibleRegion();
    if (m_TensorPixels != static_cast<size_t>(region.GetNumberOfPixels()))
      itkExceptionMacro(<< "the diffusion tensor image has " << m_TensorPixels << " pixels, the input image " << region.GetNumberOfPixels());

    madgpu_params p;
    madgpu_params_default(&p);
    p.dim = static_cast<int32_t>(Dim);
    for (unsigned int d = 0; d < Dim; ++d) {
      p.size[d] = static_cast<int32_t>(region.GetSize(d));
      p.spacing[d] = input->GetSpacing()[d];  // the tensor image's spacing is ignored, as in the reference (.hxx:131)
    }
    p.time_step = m_TimeStep;
    p.number_of_steps = static_cast<int32_t>(m_NumberOfSteps);
    p.cycle = static_cast<int32_t>(m_Cycle);
    p.iterations_per_grid = static_cast<int32_t>(m_IterationsPerGrid);
    p.tolerance = m_Tolerance;
    p.max_cycles = static_cast<int32_t>(m_MaxCycles);
    p.verbose = m_Verbose ? 1 : 0;
    p.smoother = TSmootherType::MadgpuSmoother();
    p.device = m_Device;

    madgpu_ctx* ctx = nullptr;
    if (madgpu_create(&p, &ctx) != MADGPU_OK) itkExceptionMacro(<< "madgpu_create: " << madgpu_last_error(nullptr));
    struct Guard {
      madgpu_ctx* c;
      ~Guard() { madgpu_destroy(c); }
    } guard = {ctx};

    if (this->SetTensor(ctx, m_TensorCopy.data(), m_TensorPixels) != MADGPU_OK)
      itkExceptionMacro(<< "madgpu_set_tensor: " << madgpu_last_error(ctx));

    typename OutputImageType::Pointer outputImage = OutputImageType::New();
    outputImage->SetRegions(region);
    outputImage->Allocate();
    outputImage->SetSpacing(input->GetSpacing());
    outputImage->SetOrigin(input->GetOrigin());

    m_Stats.struct_size = static_cast<int32_t>(sizeof(madgpu_stats));
    const int rc = madgpu_solve_cast(ctx, madgpu_detail::PixelTag<InputPixelType>::value, input->GetBufferPointer(),
                                     madgpu_detail::PixelTag<OutputPixelType>::value, outputImage->GetBufferPointer(), &m_Stats);
    if (rc != MADGPU_OK) itkExceptionMacro(<< "madgpu_solve: " << madgpu_last_error(ctx));

    this->AllocateOutputs();
    this->GraftOutput(outputImage);
  }

private:
  // ITK stores SymmetricSecondRankTensor<T, D> as D(D+1)/2 contiguous scalars in upper-triangular row-major order,
  // which is exactly the layout madgpu_set_tensor_* reads.
  int SetTensor(madgpu_ctx* ctx, const SymmetricSecondRankTensor<float, TInputImage::ImageDimension>* t, size_t) const
  {
    return madgpu_set_tensor_f32(ctx, reinterpret_cast<const float*>(t));
  }
  int SetTensor(madgpu_ctx* ctx, const SymmetricSecondRankTensor<double, TInputImage::ImageDimension>* t, size_t) const
  {
    return madgpu_set_tensor_f64(ctx, reinterpret_cast<const double*>(t));
  }
  template <typename TP>
  int SetTensor(madgpu_ctx* ctx, const SymmetricSecondRankTensor<TP, TInputImage::ImageDimension>* t, size_t nvox) const
  {
    const unsigned int nc = TInputImage::ImageDimension * (TInputImage::ImageDimension + 1) / 2;
    std::vector<double> tmp(nvox * nc);
    for (size_t v = 0; v < nvox; ++v)
      for (unsigned int k = 0; k < nc; ++k) tmp[v * nc + k] = static_cast<double>(t[v][k]);
    return madgpu_set_tensor_f64(ctx, tmp.data());
  }

  Precision m_TimeStep;
  unsigned int m_NumberOfSteps;
  CycleType m_Cycle;
  unsigned int m_IterationsPerGrid;
  Precision m_Tolerance;
  unsigned int m_MaxCycles;
  bool m_Verbose;
  int m_Device;
  std::vector<TensorPixelType> m_TensorCopy;  // SetDiffusionTensor's deep copy (the reference keeps an InternalTensorImage)
  size_t m_TensorPixels;
  madgpu_stats m_Stats;

  MultigridAnisotropicDiffusionImageFilter(const Self&);
  void operator=(const Self&);
};
}  // namespace itk
#endif
