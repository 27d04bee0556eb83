/*
 * ref_driver.cxx -- C entry points around the UNMODIFIED reference headers
 * (/root/reference/include/itkMultigridAnisotropicDiffusionImageFilter.{h,hxx}, include/mad/*), compiled
 * against the stand-in ITK/vnl of oracle/shim into oracle/_ref/libmadref.so (oracle/Makefile, `make ref`).
 *
 * TEST INFRASTRUCTURE ONLY.  It lets tests/test_oracle_vs_ref.py pin the C restatement (mad_oracle.c)
 * against the reference's own code on every routine of the path, and tests/golden/make_golden.py record
 * golden vectors produced BY THE REFERENCE CODE.  Nothing from /root/reference is copied: the headers are
 * included from where they lie.  Arrays are x fastest; tensors are the ITK AoS buffer.
 */
#include <cstdio>
#include <cstring>
#include <sstream>
#include <string>

#include "itkMultigridAnisotropicDiffusionImageFilter.h"
#include "mad/itkMultigridWeightedJacobiSmoother.h"

namespace
{
template <unsigned int D>
struct Hier {
  typedef itk::mad::GridsHierarchy<D> GridsType;
  typedef itk::mad::DirectSolver<D> SolverType;
  typedef itk::Image<double, D> ImageType;
  typedef typename GridsType::TensorImageType TensorImageType;
  typename TensorImageType::Pointer tensor;
  GridsType* grids;
  SolverType* solver;
  Hier() : grids(nullptr), solver(nullptr) {}
  ~Hier() { delete solver; delete grids; }
};

struct Handle {
  int dim;
  Hier<2>* h2;
  Hier<3>* h3;
};

template <unsigned int D>
typename itk::Image<double, D>::Pointer make_image(const int* n, const double* h, const double* data)
{
  typedef itk::Image<double, D> ImageType;
  typename ImageType::Pointer img = ImageType::New();
  typename ImageType::IndexType idx;
  typename ImageType::SizeType size;
  typename ImageType::SpacingType sp;
  idx.Fill(0);
  for (unsigned int d = 0; d < D; ++d) { size[d] = n[d]; sp[d] = h ? h[d] : 1.0; }
  img->SetRegions(typename ImageType::RegionType(idx, size));
  img->Allocate();
  img->SetSpacing(sp);
  if (data) std::memcpy(img->GetBufferPointer(), data, sizeof(double) * img->GetLargestPossibleRegion().GetNumberOfPixels());
  return img;
}

template <typename TPixel, unsigned int D>
typename itk::Image<itk::SymmetricSecondRankTensor<TPixel, D>, D>::Pointer make_tensor(const int* n, const double* h, const double* aos)
{
  typedef itk::Image<itk::SymmetricSecondRankTensor<TPixel, D>, D> TensorImageType;
  typename TensorImageType::Pointer t = TensorImageType::New();
  typename TensorImageType::IndexType idx;
  typename TensorImageType::SizeType size;
  typename TensorImageType::SpacingType sp;
  idx.Fill(0);
  for (unsigned int d = 0; d < D; ++d) { size[d] = n[d]; sp[d] = h[d]; }
  t->SetRegions(typename TensorImageType::RegionType(idx, size));
  t->Allocate();
  t->SetSpacing(sp);
  const unsigned int nc = D * (D + 1) / 2;
  const size_t nv = t->GetLargestPossibleRegion().GetNumberOfPixels();
  for (size_t v = 0; v < nv; ++v)
    for (unsigned int k = 0; k < nc; ++k) t->GetBufferPointer()[v][k] = static_cast<TPixel>(aos[v * nc + k]);
  return t;
}

template <unsigned int D>
Hier<D>* build(const int* n, const double* h, double dt, const double* aos)
{
  Hier<D>* H = new Hier<D>();
  H->tensor = make_tensor<double, D>(n, h, aos);
  typename Hier<D>::ImageType::Pointer img = make_image<D>(n, h, nullptr);
  H->grids = new typename Hier<D>::GridsType(img->GetLargestPossibleRegion(), img->GetSpacing(), H->tensor, dt);
  return H;
}

template <unsigned int D>
void level_info(Hier<D>* H, int l, int* n, double* h, int* cent)
{
  typename Hier<D>::GridsType::Grid* g = H->grids->GetGridAtLevel(l);
  for (unsigned int d = 0; d < D; ++d) {
    n[d] = static_cast<int>(g->g_Region.GetSize(d));
    h[d] = g->g_Spacing[d];
    cent[d] = g->g_Centering[d] == itk::mad::InterGridOperators<D>::cell ? 1 : 0;
  }
}

template <unsigned int D>
void level_stencil(Hier<D>* H, int l, double* out, int* active)
{
  typedef typename Hier<D>::GridsType::StencilImageType StencilImageType;
  typename StencilImageType::Pointer S = H->grids->GetCoarseOperatorAtLevel(l);
  const size_t nv = S->GetLargestPossibleRegion().GetNumberOfPixels();
  unsigned int ns = 1;
  for (unsigned int d = 0; d < D; ++d) ns *= 3;
  for (size_t v = 0; v < nv; ++v)
    for (unsigned int k = 0; k < ns; ++k) out[v * ns + k] = S->GetBufferPointer()[v][k];
  if (active) {
    for (unsigned int k = 0; k < ns; ++k) active[k] = 0;
    typename StencilImageType::StencilType probe;
    probe.SetRadius(1);
    typename StencilImageType::OffsetListType lst = S->GetActiveOffsetList();
    int order = 1;
    for (typename StencilImageType::OffsetListType::iterator it = lst.begin(); it != lst.end(); ++it, ++order)
      for (unsigned int k = 0; k < ns; ++k)
        if (probe.GetOffset(k) == *it) active[k] = order;  // position in the active list (1-based)
  }
}

template <unsigned int D>
void smooth_or_residual(Hier<D>* H, int l, int what, int smoother, const double* u, const double* f, double* out)
{
  typedef typename Hier<D>::ImageType ImageType;
  int n[3];
  double h[3];
  int c[3];
  level_info<D>(H, l, n, h, c);
  typename ImageType::Pointer U = make_image<D>(n, h, u), F = make_image<D>(n, h, f), R;
  itk::mad::MultigridGaussSeidelSmoother<D> gs;
  itk::mad::MultigridWeightedJacobiSmoother<D> wj;
  const itk::mad::MultigridSmoother<D>* s = smoother == 0 ? static_cast<const itk::mad::MultigridSmoother<D>*>(&gs) : &wj;
  if (what == 0) R = s->SingleIteration(U, F, H->grids->GetCoarseOperatorAtLevel(l));
  else R = s->ComputeResidual(U, F, H->grids->GetCoarseOperatorAtLevel(l));
  std::memcpy(out, R->GetBufferPointer(), sizeof(double) * R->GetLargestPossibleRegion().GetNumberOfPixels());
}

template <unsigned int D>
void transfer(int what, const int* n, const int* cent, const double* in, double* out, int* nout)
{
  typedef itk::mad::InterGridOperators<D> IGO;
  std::array<typename IGO::CoarseGridCenteringType, D> c;
  for (unsigned int d = 0; d < D; ++d) c[d] = cent[d] ? IGO::cell : IGO::vertex;
  IGO op(c);
  typename itk::Image<double, D>::Pointer I = make_image<D>(n, nullptr, in), O;
  O = what == 0 ? op.Restriction(I) : op.Interpolation(I);
  for (unsigned int d = 0; d < D; ++d) nout[d] = static_cast<int>(O->GetLargestPossibleRegion().GetSize(d));
  if (out) std::memcpy(out, O->GetBufferPointer(), sizeof(double) * O->GetLargestPossibleRegion().GetNumberOfPixels());
}

template <unsigned int D>
void direct_solve(Hier<D>* H, const double* f, double* out)
{
  const int L = H->grids->GetMaxDepth();
  if (!H->solver) H->solver = new typename Hier<D>::SolverType(H->grids->GetCoarseOperatorAtLevel(L));
  int n[3];
  double h[3];
  int c[3];
  level_info<D>(H, L, n, h, c);
  typename Hier<D>::ImageType::Pointer F = make_image<D>(n, h, f);
  typename Hier<D>::ImageType::Pointer E = H->solver->Solve(F);
  std::memcpy(out, E->GetBufferPointer(), sizeof(double) * E->GetLargestPossibleRegion().GetNumberOfPixels());
}

// The filter itself, exactly as the reference's test programs drive it (test/itk2DDiffusionTest_WJ.cxx:88-109).
template <typename TPixel, unsigned int D, typename TSmoother>
int run_filter(const int* n, const double* h, const double* aos, const double* image, int cycle, int nu, double dt, double tol, int max_cycles,
               int steps, int verbose, double* out, char* log, int logcap)
{
  typedef itk::Image<TPixel, D> ImageType;
  typedef itk::MultigridAnisotropicDiffusionImageFilter<ImageType, ImageType, TSmoother> FilterType;
  typename ImageType::Pointer img = ImageType::New();
  typename ImageType::IndexType idx;
  typename ImageType::SizeType size;
  typename ImageType::SpacingType sp;
  idx.Fill(0);
  size_t nv = 1;
  for (unsigned int d = 0; d < D; ++d) { size[d] = n[d]; sp[d] = h[d]; nv *= n[d]; }
  img->SetRegions(typename ImageType::RegionType(idx, size));
  img->Allocate();
  img->SetSpacing(sp);
  for (size_t v = 0; v < nv; ++v) img->GetBufferPointer()[v] = static_cast<TPixel>(image[v]);
  typename FilterType::InputTensorImageType::Pointer tensor = make_tensor<TPixel, D>(n, h, aos);

  typename FilterType::Pointer filter = FilterType::New();
  filter->SetInput(img);
  filter->SetDiffusionTensor(tensor);
  filter->SetIterationsPerGrid(nu);
  filter->SetTimeStep(dt);
  filter->SetNumberOfSteps(steps);
  filter->SetMaxCycles(max_cycles);
  filter->SetTolerance(tol);
  filter->SetVerbose(verbose != 0);
  filter->SetCycle(static_cast<typename FilterType::CycleType>(cycle));

  std::ostringstream cap;
  std::streambuf* old = std::cout.rdbuf(cap.rdbuf());
  const std::streamsize oldprec = std::cout.precision(17);
  try {
    filter->Update();
  } catch (...) {
    std::cout.rdbuf(old);
    std::cout.precision(oldprec);
    return -1;
  }
  std::cout.rdbuf(old);
  std::cout.precision(oldprec);
  ImageType* o = filter->GetOutput();
  for (size_t v = 0; v < nv; ++v) out[v] = static_cast<double>(o->GetBufferPointer()[v]);
  const std::string s = cap.str();
  if (log && logcap > 0) {
    const size_t m = std::min(s.size(), static_cast<size_t>(logcap - 1));
    std::memcpy(log, s.data(), m);
    log[m] = 0;
  }
  return static_cast<int>(s.size());
}
}  // namespace

extern "C" {

void* mr_create(int dim, const int* n, const double* h, double dt, const double* tensor_aos)
{
  Handle* H = new Handle();
  H->dim = dim;
  H->h2 = nullptr;
  H->h3 = nullptr;
  try {
    if (dim == 2) H->h2 = build<2>(n, h, dt, tensor_aos);
    else H->h3 = build<3>(n, h, dt, tensor_aos);
  } catch (...) {
    delete H;
    return nullptr;
  }
  return H;
}

void mr_destroy(void* p)
{
  Handle* H = static_cast<Handle*>(p);
  if (!H) return;
  delete H->h2;
  delete H->h3;
  delete H;
}

int mr_nlevels(void* p)
{
  Handle* H = static_cast<Handle*>(p);
  return 1 + static_cast<int>(H->dim == 2 ? H->h2->grids->GetMaxDepth() : H->h3->grids->GetMaxDepth());
}

void mr_level_info(void* p, int l, int* n, double* h, int* cent)
{
  Handle* H = static_cast<Handle*>(p);
  n[2] = 1; h[2] = 1.0; cent[2] = 0;
  if (H->dim == 2) level_info<2>(H->h2, l, n, h, cent);
  else level_info<3>(H->h3, l, n, h, cent);
}

void mr_level_stencil(void* p, int l, double* out, int* active)
{
  Handle* H = static_cast<Handle*>(p);
  if (H->dim == 2) level_stencil<2>(H->h2, l, out, active);
  else level_stencil<3>(H->h3, l, out, active);
}

/* smoother: 0 Gauss-Seidel (lexicographic), 1 weighted Jacobi (default-constructed: omega = 2/3) */
void mr_smooth(void* p, int l, int smoother, const double* u, const double* f, double* out)
{
  Handle* H = static_cast<Handle*>(p);
  if (H->dim == 2) smooth_or_residual<2>(H->h2, l, 0, smoother, u, f, out);
  else smooth_or_residual<3>(H->h3, l, 0, smoother, u, f, out);
}

void mr_residual(void* p, int l, int smoother, const double* u, const double* f, double* out)
{
  Handle* H = static_cast<Handle*>(p);
  if (H->dim == 2) smooth_or_residual<2>(H->h2, l, 1, smoother, u, f, out);
  else smooth_or_residual<3>(H->h3, l, 1, smoother, u, f, out);
}

int mr_direct_solve(void* p, const double* f, double* out)
{
  Handle* H = static_cast<Handle*>(p);
  try {
    if (H->dim == 2) direct_solve<2>(H->h2, f, out);
    else direct_solve<3>(H->h3, f, out);
  } catch (...) {
    return -1;
  }
  return 0;
}

/* what: 0 Restriction, 1 Interpolation; cent[d]: 0 vertex, 1 cell; nout receives the output size; out may be NULL */
void mr_transfer(int dim, int what, const int* n, const int* cent, const double* in, double* out, int* nout)
{
  nout[2] = 1;
  if (dim == 2) transfer<2>(what, n, cent, in, out, nout);
  else transfer<3>(what, n, cent, in, out, nout);
}

/* pixel: 0 double, 1 float, 2 short, 3 unsigned char (input is given as doubles and cast to the pixel type first) */
int mr_filter(int dim, int pixel, int smoother, const int* n, const double* h, const double* tensor_aos, const double* image, int cycle, int nu,
              double dt, double tol, int max_cycles, int steps, int verbose, double* out, char* log, int logcap)
{
#define MR_RUN(P, D)                                                                                                                    \
  (smoother == 0 ? run_filter<P, D, itk::mad::MultigridGaussSeidelSmoother<D> >(n, h, tensor_aos, image, cycle, nu, dt, tol, max_cycles, steps, \
                                                                                 verbose, out, log, logcap)                                \
                 : run_filter<P, D, itk::mad::MultigridWeightedJacobiSmoother<D> >(n, h, tensor_aos, image, cycle, nu, dt, tol, max_cycles,   \
                                                                                    steps, verbose, out, log, logcap))
  if (dim == 2) {
    switch (pixel) {
      case 0: return MR_RUN(double, 2);
      case 1: return MR_RUN(float, 2);
      case 3: return MR_RUN(unsigned char, 2);
      default: return -2;
    }
  }
  switch (pixel) {
    case 0: return MR_RUN(double, 3);
    case 1: return MR_RUN(float, 3);
    case 2: return MR_RUN(short, 3);
    default: return -2;
  }
#undef MR_RUN
}

}  // extern "C"
