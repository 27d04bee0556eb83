// Compile-and-run test of the drop-in VED filter header (include/itkVEDMultigridImageFilter.h) against the stand-in ITK of
// oracle/shim, written the way the reference's test program drives the original filter (test/itkVEDTest_GS.cxx:46-101).
//
//   ved_dropin_test <cycle v|fmg|s> <pixel i16|f64> <in.raw> <out.raw> nx ny nz sx sy sz
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "itkVEDMultigridImageFilter.h"

template <typename TPixel>
static int run(char** a)
{
  typedef itk::Image<TPixel, 3> ImageType;
  typedef itk::mad::MultigridGaussSeidelSmoother<ImageType::ImageDimension> smootherType;
  typedef itk::VEDMultigridImageFilter<ImageType, ImageType, smootherType> filterType;
  typename ImageType::IndexType idx;
  typename ImageType::SizeType size;
  typename ImageType::SpacingType sp;
  idx.Fill(0);
  size_t nv = 1;
  for (unsigned int d = 0; d < 3; ++d) { size[d] = std::atoi(a[5 + d]); sp[d] = std::atof(a[8 + d]); nv *= size[d]; }
  typename ImageType::Pointer input = ImageType::New();
  input->SetRegions(typename ImageType::RegionType(idx, size));
  input->Allocate();
  input->SetSpacing(sp);
  {
    std::ifstream f(a[3], std::ios::binary);
    f.read(reinterpret_cast<char*>(input->GetBufferPointer()), static_cast<std::streamsize>(nv * sizeof(TPixel)));
    if (!f) { std::fprintf(stderr, "short read: %s\n", a[3]); return 3; }
  }
  typename filterType::Pointer filter = filterType::New();
  typedef typename filterType::CycleType cycleType;
  cycleType cycle = filterType::MADFilterType::VCYCLE;
  if (std::strcmp(a[1], "fmg") == 0) cycle = filterType::MADFilterType::FMG;
  else if (std::strcmp(a[1], "s") == 0) cycle = filterType::MADFilterType::SMOOTHER;
  filter->SetCycle(cycle);
  filter->SetDiffusionIterationsPerGrid(3);
  filter->SetInput(input);
  filter->SetVerbose(false);
  std::vector<double> sigmaValues(5);
  sigmaValues[0] = 0.300;
  sigmaValues[1] = 0.482;
  sigmaValues[2] = 0.775;
  sigmaValues[3] = 1.245;
  sigmaValues[4] = 2.000;
  filter->SetScales(sigmaValues);
  filter->SetAlpha(0.5);
  filter->SetBeta(0.5);
  filter->SetGamma(5.);
  filter->SetEpsilon(0.01);
  filter->SetSensitivity(10.);
  filter->SetIterations(1);
  filter->SetTolerance(1e-10);
  filter->SetTimeStep(0.1);
  filter->SetDiffusionIterations(4);
  filter->SetOmega(1.5);
  try {
    filter->Update();
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 2;
  }
  const madgpu_stats& st = filter->GetStatistics();
  std::printf("steps %d cycles", st.steps);
  for (int s = 0; s < st.steps; ++s) std::printf(" %d", st.cycles_per_step[s]);
  std::printf(" scales %d hessian_ms %.3f vesselness_ms %.3f\n", filter->GetFrontEndStatistics().scales, filter->GetFrontEndStatistics().hessian_ms,
              filter->GetFrontEndStatistics().vesselness_ms);
  std::ofstream o(a[4], std::ios::binary);
  o.write(reinterpret_cast<const char*>(filter->GetOutput()->GetBufferPointer()), static_cast<std::streamsize>(nv * sizeof(TPixel)));
  return 0;
}

int main(int argc, char** argv)
{
  if (argc != 11) { std::fprintf(stderr, "usage: see header comment\n"); return 1; }
  return std::string(argv[2]) == "i16" ? run<short>(argv) : run<double>(argv);
}
