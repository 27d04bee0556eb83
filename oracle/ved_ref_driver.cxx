/*
 * ved_ref_driver.cxx -- C entry points around the UNMODIFIED /root/reference/include/itkVEDMultigridImageFilter.{h,hxx},
 * compiled against the stand-in ITK/vnl of oracle/shim into oracle/_ref/libmadref.so (oracle/Makefile, `make ref`).
 *
 * TEST INFRASTRUCTURE ONLY.  It pins the VED part of the oracle (oracle/ved_oracle.c: vo_vesselness,
 * vo_update_vesselness, vo_generate_tensor; oracle/ved.py: ved_filter) against the reference's own code.  The Hessian
 * filter and the eigen-solver underneath are stand-ins (see oracle/shim/mini_itk_ved.h): third-party arithmetic that is
 * not part of the reference.  The filter's private methods are reached by compiling this one translation unit with
 * `private` spelt `public` -- nothing of the reference is edited or copied.
 */
#include <cmath>
#include <cstdio>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

#include "mini_itk_ved.h"

#define private public
#include "itkVEDMultigridImageFilter.h"
#include "mad/itkMultigridWeightedJacobiSmoother.h"
#undef private

namespace
{
typedef itk::Image<double, 3> ImageD;
typedef itk::Image<short, 3> ImageS;

template <typename TImage>
typename TImage::Pointer alloc_image3(const int* n, const double* h)
{
  typename TImage::Pointer img = TImage::New();
  typename TImage::IndexType idx;
  typename TImage::SizeType size;
  typename TImage::SpacingType sp;
  idx.Fill(0);
  for (unsigned int d = 0; d < 3; ++d) { size[d] = n[d]; sp[d] = h[d]; }
  img->SetRegions(typename TImage::RegionType(idx, size));
  img->Allocate();
  img->SetSpacing(sp);
  return img;
}

template <typename TImage>
typename TImage::Pointer make_image3(const int* n, const double* h, const double* data)
{
  typename TImage::Pointer img = TImage::New();
  typename TImage::IndexType idx;
  typename TImage::SizeType size;
  typename TImage::SpacingType sp;
  idx.Fill(0);
  size_t nv = 1;
  for (unsigned int d = 0; d < 3; ++d) { size[d] = n[d]; sp[d] = h[d]; nv *= n[d]; }
  img->SetRegions(typename TImage::RegionType(idx, size));
  img->Allocate();
  img->SetSpacing(sp);
  if (data)
    for (size_t v = 0; v < nv; ++v) img->GetBufferPointer()[v] = static_cast<typename TImage::PixelType>(data[v]);
  return img;
}

struct CoutCapture {
  std::ostringstream cap;
  std::streambuf* old;
  CoutCapture() : old(std::cout.rdbuf(cap.rdbuf())) {}
  ~CoutCapture() { std::cout.rdbuf(old); }
};

template <typename TFilter>
void set_params(TFilter* f, const double* p)
{
  f->SetAlpha(p[0]); f->SetBeta(p[1]); f->SetGamma(p[2]); f->SetEpsilon(p[3]); f->SetOmega(p[4]); f->SetSensitivity(p[5]);
}

template <typename TPixel, typename TSmoother>
int run_ved(const int* n, const double* h, const double* image, const double* params, const double* scales, int nscales, int iterations,
            int diffusion_iterations, int cycle, double dt, double tol, int nu, double* out, double* tensor_out)
{
  typedef itk::Image<TPixel, 3> ImageType;
  typedef itk::VEDMultigridImageFilter<ImageType, ImageType, TSmoother> FilterType;
  typename ImageType::Pointer img = make_image3<ImageType>(n, h, image);
  typename FilterType::Pointer f = FilterType::New();
  set_params(f.GetPointer(), params);
  f->SetScales(std::vector<double>(scales, scales + nscales));
  f->SetIterations(iterations);
  f->SetDiffusionIterations(diffusion_iterations);
  f->SetCycle(static_cast<typename FilterType::CycleType>(cycle));
  f->SetTimeStep(dt);
  f->SetTolerance(tol);
  f->SetDiffusionIterationsPerGrid(nu);
  f->SetVerbose(false);
  f->SetInput(img);
  CoutCapture quiet;
  try {
    f->Update();
  } catch (...) {
    return -1;
  }
  const size_t nv = static_cast<size_t>(n[0]) * n[1] * n[2];
  for (size_t v = 0; v < nv; ++v) out[v] = static_cast<double>(f->GetOutput()->GetBufferPointer()[v]);
  if (tensor_out)  // the tensor of the LAST outer iteration
    for (size_t v = 0; v < nv; ++v)
      for (unsigned int k = 0; k < 6; ++k) tensor_out[v * 6 + k] = f->m_DiffusionTensor->GetBufferPointer()[v][k];
  return 0;
}
}  // namespace

extern "C" {

/* VesselnessFunction (private, itkVEDMultigridImageFilter.hxx:176-212) on eigenvalues sorted by magnitude */
double mrv_vesselness(const double* e, double alpha, double beta, double gamma)
{
  typedef itk::VEDMultigridImageFilter<ImageD, ImageD> FilterType;
  FilterType::Pointer f = FilterType::New();
  f->SetAlpha(alpha); f->SetBeta(beta); f->SetGamma(gamma);
  vnl_vector<double> ev(3);
  ev[0] = e[0]; ev[1] = e[1]; ev[2] = e[2];
  return f->VesselnessFunction(ev);
}

/* The scale loop body + GenerateDiffusionTensor on caller-supplied Hessians (UpdateVesselness :215-299 once per Hessian,
 * then GenerateDiffusionTensor :302-378).  hessians: nscales blocks of nvox*6 doubles.  params = alpha, beta, gamma,
 * epsilon, omega, sensitivity.  Outputs (each may be NULL): response nvox, eigenvalues nvox*3, eigenvectors nvox*9
 * (row-major Matrix), tensor nvox*6. */
int mrv_tensor_from_hessians(const int* n, const double* h, const double* hessians, int nscales, const double* params, double* response,
                             double* eigenvalues, double* eigenvectors, double* tensor)
{
  typedef itk::VEDMultigridImageFilter<ImageD, ImageD> FilterType;
  typedef FilterType::TensorImageType TensorImageType;
  FilterType::Pointer f = FilterType::New();
  set_params(f.GetPointer(), params);
  const size_t nv = static_cast<size_t>(n[0]) * n[1] * n[2];
  CoutCapture quiet;
  try {
    for (int s = 0; s < nscales; ++s) {
      TensorImageType::Pointer H = alloc_image3<TensorImageType>(n, h);
      for (size_t v = 0; v < nv; ++v)
        for (unsigned int k = 0; k < 6; ++k) H->GetBufferPointer()[v][k] = hessians[(static_cast<size_t>(s) * nv + v) * 6 + k];
      f->UpdateVesselness(H);
    }
    f->GenerateDiffusionTensor();
  } catch (...) {
    return -1;
  }
  for (size_t v = 0; v < nv; ++v) {
    if (response) response[v] = f->m_MaxVesselnessResponse->GetBufferPointer()[v];
    if (eigenvalues)
      for (unsigned int d = 0; d < 3; ++d) eigenvalues[v * 3 + d] = f->m_MaxVesselnessEigenValues->GetBufferPointer()[v][d];
    if (eigenvectors)
      for (unsigned int r = 0; r < 3; ++r)
        for (unsigned int c = 0; c < 3; ++c) eigenvectors[v * 9 + r * 3 + c] = f->m_MaxVesselnessEigenVectors->GetBufferPointer()[v](r, c);
    if (tensor)
      for (unsigned int k = 0; k < 6; ++k) tensor[v * 6 + k] = f->m_DiffusionTensor->GetBufferPointer()[v][k];
  }
  return 0;
}

/* The whole filter as test/itkVEDTest_GS.cxx drives it (:64-101).  pixel: 0 double, 2 short; smoother: 0 GS, 1 WJ. */
int mrv_filter(int pixel, int smoother, const int* n, const double* h, const double* image, const double* params, const double* scales, int nscales,
               int iterations, int diffusion_iterations, int cycle, double dt, double tol, int nu, double* out, double* tensor_out)
{
#define MRV_RUN(P)                                                                                                                          \
  (smoother == 0 ? run_ved<P, itk::mad::MultigridGaussSeidelSmoother<3> >(n, h, image, params, scales, nscales, iterations, diffusion_iterations, \
                                                                          cycle, dt, tol, nu, out, tensor_out)                                  \
                 : run_ved<P, itk::mad::MultigridWeightedJacobiSmoother<3> >(n, h, image, params, scales, nscales, iterations,                   \
                                                                             diffusion_iterations, cycle, dt, tol, nu, out, tensor_out))
  switch (pixel) {
    case 0: return MRV_RUN(double);
    case 2: return MRV_RUN(short);
    default: return -2;
  }
#undef MRV_RUN
}

}  // extern "C"
