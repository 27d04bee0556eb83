// Compile-and-run test of the drop-in ITK header (include/itkMultigridAnisotropicDiffusionImageFilter.h) against
// the stand-in ITK of oracle/shim, written the way the reference's test programs drive the filter
// (test/itk2DDiffusionTest_WJ.cxx:61-109, test/itkVEDTest_GS.cxx:46-92).
//
//   dropin_test <dim> <smoother gs|wj> <cycle v|fmg|s> <nu> <dt> <tol> <steps> <pixel f32|f64|i16|u8> <in.raw> <tensor.raw> <out.raw>
//               nx ny [nz] sx sy [sz]
// in.raw / out.raw hold the pixel type; tensor.raw holds the tensor in the pixel type for f32/f64 and in f64 for
// the integer pixel types (it is cast to the pixel type, as an ITK user of a short image would have to).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <vector>

#include "itkMultigridAnisotropicDiffusionImageFilter.h"

template <typename T>
static std::vector<T> read_raw(const char* path, size_t n)
{
  std::vector<T> v(n);
  std::ifstream f(path, std::ios::binary);
  f.read(reinterpret_cast<char*>(v.data()), static_cast<std::streamsize>(n * sizeof(T)));
  if (!f) { std::fprintf(stderr, "short read: %s\n", path); std::exit(3); }
  return v;
}

template <typename TPixel, unsigned int D, typename TSmoother>
static int run(char** a, const int* n, const double* h)
{
  typedef itk::Image<TPixel, D> ImageType;
  typedef itk::MultigridAnisotropicDiffusionImageFilter<ImageType, ImageType, TSmoother> FilterType;
  typedef typename FilterType::InputTensorImageType TensorImageType;
  typename ImageType::IndexType idx;
  typename ImageType::SizeType size;
  typename ImageType::SpacingType sp;
  idx.Fill(0);
  size_t nv = 1;
  for (unsigned int d = 0; d < D; ++d) { size[d] = n[d]; sp[d] = h[d]; nv *= n[d]; }
  typename ImageType::Pointer img = ImageType::New();
  img->SetRegions(typename ImageType::RegionType(idx, size));
  img->Allocate();
  img->SetSpacing(sp);
  const std::vector<TPixel> in = read_raw<TPixel>(a[9], nv);
  std::memcpy(img->GetBufferPointer(), in.data(), nv * sizeof(TPixel));
  const unsigned int nc = D * (D + 1) / 2;
  typename TensorImageType::Pointer tensor = TensorImageType::New();
  tensor->SetRegions(typename ImageType::RegionType(idx, size));
  tensor->Allocate();
  if (sizeof(TPixel) >= 4) {
    const std::vector<TPixel> t = read_raw<TPixel>(a[10], nv * nc);
    for (size_t v = 0; v < nv; ++v)
      for (unsigned int k = 0; k < nc; ++k) tensor->GetBufferPointer()[v][k] = t[v * nc + k];
  } else {
    const std::vector<double> t = read_raw<double>(a[10], nv * nc);
    for (size_t v = 0; v < nv; ++v)
      for (unsigned int k = 0; k < nc; ++k) tensor->GetBufferPointer()[v][k] = static_cast<TPixel>(t[v * nc + k]);
  }
  typename FilterType::Pointer filter = FilterType::New();
  filter->SetInput(img);
  filter->SetDiffusionTensor(tensor);
  // the reference deep-copies the tensor at Set time (.hxx:66-101): scribbling over the caller's image and dropping the last
  // reference before Update() is valid usage and must not change the result
  tensor->FillBuffer(typename TensorImageType::PixelType());
  tensor = nullptr;
  filter->SetIterationsPerGrid(static_cast<unsigned int>(std::atoi(a[4])));
  filter->SetTimeStep(std::atof(a[5]));
  filter->SetTolerance(std::atof(a[6]));
  filter->SetNumberOfSteps(static_cast<unsigned int>(std::atoi(a[7])));
  filter->SetMaxCycles(100);
  filter->SetVerbose(false);
  const std::string cyc = a[3];
  filter->SetCycle(cyc == "fmg" ? FilterType::FMG : cyc == "s" ? FilterType::SMOOTHER : FilterType::VCYCLE);
  try {
    filter->Update();
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 2;
  }
  const madgpu_stats& st = filter->GetStatistics();
  std::printf("steps %d cycles", st.steps);
  for (int s = 0; s < st.steps; ++s) std::printf(" %d", st.cycles_per_step[s]);
  std::printf(" relres");
  for (int s = 0; s < st.steps; ++s) std::printf(" %.3e", st.final_relres[s]);
  std::printf("\n");
  std::ofstream o(a[11], std::ios::binary);
  o.write(reinterpret_cast<const char*>(filter->GetOutput()->GetBufferPointer()), static_cast<std::streamsize>(nv * sizeof(TPixel)));
  return 0;
}

template <typename TPixel, unsigned int D>
static int pick_smoother(char** a, const int* n, const double* h)
{
  if (std::string(a[2]) == "wj") return run<TPixel, D, itk::mad::MultigridWeightedJacobiSmoother<D> >(a, n, h);
  return run<TPixel, D, itk::mad::MultigridGaussSeidelSmoother<D> >(a, n, h);
}

template <unsigned int D>
static int pick_pixel(char** a, const int* n, const double* h)
{
  const std::string p = a[8];
  if (p == "f32") return pick_smoother<float, D>(a, n, h);
  if (p == "f64") return pick_smoother<double, D>(a, n, h);
  if (p == "i16") return pick_smoother<short, D>(a, n, h);
  if (p == "u8") return pick_smoother<unsigned char, D>(a, n, h);
  return 4;
}

int main(int argc, char** argv)
{
  if (argc < 12) { std::fprintf(stderr, "usage: see header comment\n"); return 1; }
  const int dim = std::atoi(argv[1]);
  if (argc != 12 + 2 * dim) { std::fprintf(stderr, "expected %d size/spacing arguments\n", 2 * dim); return 1; }
  int n[3] = {1, 1, 1};
  double h[3] = {1, 1, 1};
  for (int d = 0; d < dim; ++d) { n[d] = std::atoi(argv[12 + d]); h[d] = std::atof(argv[12 + dim + d]); }
  return dim == 2 ? pick_pixel<2>(argv, n, h) : pick_pixel<3>(argv, n, h);
}
