/*
 * mini_itk_ved.h -- the additional ITK / vnl surface that /root/reference/include/itkVEDMultigridImageFilter.{h,hxx}
 * uses on top of mini_itk.h, so that the VED filter also compiles UNMODIFIED into oracle/_ref/libmadref.so.
 * TEST INFRASTRUCTURE ONLY.  Written from the public API as the reference uses it; not ITK / VXL code.
 *
 * Two of the stand-ins replace third-party ARITHMETIC that is absent from /root/reference (and from this image):
 *   - itk::HessianRecursiveGaussianImageFilter  -> vo_hessian   (oracle/ved_oracle.c)
 *   - vnl_symmetric_eigensystem<double>         -> vo_eig3      (oracle/ved_oracle.c)
 * so what this build pins is the reference's OWN code around them (vesselness function, magnitude sort, arg-max over
 * scales, eigenvector bookkeeping, Q D Q^T, DiffusionStep, casts) -- not ITK's Hessian or VXL's eigen-solver.
 */
#ifndef MINI_ITK_VED_H
#define MINI_ITK_VED_H

#include <cmath>
#include <cstdlib>

#include "mini_itk.h"

/* The reference calls abs() unqualified on doubles (itkVEDMultigridImageFilter.hxx:197,202,266-268); in an ITK build the
 * floating-point overload is visible at global scope.  Make that explicit here. */
using std::abs;

extern "C" {
int vo_hessian(const int* n, const double* h, double sigma, int normalize_across_scale, const double* image, double* hessian_aos);
void vo_eig3(const double* a, double* w, double* V);
}

namespace itk
{
template <typename T, unsigned int N>
class FixedArray
{
public:
  FixedArray() { for (unsigned int i = 0; i < N; ++i) m_v[i] = T(); }
  T& operator[](unsigned int i) { return m_v[i]; }
  const T& operator[](unsigned int i) const { return m_v[i]; }
  void Fill(const T& v) { for (unsigned int i = 0; i < N; ++i) m_v[i] = v; }
private:
  T m_v[N];
};

template <typename T, unsigned int R = 3, unsigned int C = 3>
class Matrix
{
public:
  Matrix() { Fill(T()); }
  void Fill(const T& v) { for (unsigned int r = 0; r < R; ++r) for (unsigned int c = 0; c < C; ++c) m_v[r][c] = v; }
  T& operator()(unsigned int r, unsigned int c) { return m_v[r][c]; }
  const T& operator()(unsigned int r, unsigned int c) const { return m_v[r][c]; }
  Matrix<T, C, R> GetTranspose() const
  {
    Matrix<T, C, R> t;
    for (unsigned int r = 0; r < R; ++r) for (unsigned int c = 0; c < C; ++c) t(c, r) = m_v[r][c];
    return t;
  }
  /* square product, rows times columns, summed in index order (as vnl_matrix_fixed does) */
  Matrix operator*(const Matrix& o) const
  {
    Matrix p;
    for (unsigned int r = 0; r < R; ++r)
      for (unsigned int c = 0; c < C; ++c) {
        T s = T();
        for (unsigned int k = 0; k < C; ++k) s += m_v[r][k] * o.m_v[k][c];
        p.m_v[r][c] = s;
      }
    return p;
  }
private:
  T m_v[R][C];
};

template <typename TInputImage, typename TOutputImage>
class HessianRecursiveGaussianImageFilter : public LightObject
{
public:
  typedef HessianRecursiveGaussianImageFilter Self;
  typedef SmartPointer<Self> Pointer;
  itkNewMacro(Self);
  void SetInput(const TInputImage* in) { m_Input = in; }
  void SetNormalizeAcrossScale(bool b) { m_Normalize = b; }
  void SetSigma(double s) { m_Sigma = s; }
  void Update()
  {
    const typename TInputImage::RegionType region = m_Input->GetLargestPossibleRegion();
    int n[3];
    double h[3];
    for (unsigned int d = 0; d < 3; ++d) { n[d] = static_cast<int>(region.GetSize(d)); h[d] = m_Input->GetSpacing()[d]; }
    const size_t nv = region.GetNumberOfPixels();
    std::vector<double> img(nv), hes(nv * 6);
    for (size_t v = 0; v < nv; ++v) img[v] = static_cast<double>(m_Input->GetBufferPointer()[v]);
    if (vo_hessian(n, h, m_Sigma, m_Normalize ? 1 : 0, img.data(), hes.data()) != 0) throw std::runtime_error("HessianRecursiveGaussianImageFilter stand-in failed");
    m_Output = TOutputImage::New();
    m_Output->SetRegions(region);
    m_Output->Allocate();
    m_Output->SetSpacing(m_Input->GetSpacing());
    m_Output->SetOrigin(m_Input->GetOrigin());
    for (size_t v = 0; v < nv; ++v)
      for (unsigned int k = 0; k < 6; ++k) m_Output->GetBufferPointer()[v][k] = hes[v * 6 + k];
  }
  TOutputImage* GetOutput() { return m_Output.GetPointer(); }
protected:
  HessianRecursiveGaussianImageFilter() : m_Input(nullptr), m_Normalize(false), m_Sigma(1.0) {}
private:
  const TInputImage* m_Input;
  bool m_Normalize;
  double m_Sigma;
  typename TOutputImage::Pointer m_Output;
};
}  // namespace itk

template <typename T>
class vnl_matrix
{
public:
  vnl_matrix(unsigned int r, unsigned int c) : m_r(r), m_c(c), m_d(static_cast<size_t>(r) * c) {}
  T& operator()(unsigned int r, unsigned int c) { return m_d[static_cast<size_t>(r) * m_c + c]; }
  const T& operator()(unsigned int r, unsigned int c) const { return m_d[static_cast<size_t>(r) * m_c + c]; }
  unsigned int rows() const { return m_r; }
  unsigned int cols() const { return m_c; }
private:
  unsigned int m_r, m_c;
  std::vector<T> m_d;
};

/* ascending eigenvalues, get_eigenvector(i) = unit eigenvector of get_eigenvalue(i) (the VXL contract) */
template <typename T>
class vnl_symmetric_eigensystem
{
public:
  explicit vnl_symmetric_eigensystem(const vnl_matrix<T>& M)
  {
    if (M.rows() != 3 || M.cols() != 3) throw std::runtime_error("vnl_symmetric_eigensystem stand-in: 3x3 only");
    const double a[6] = {M(0, 0), M(0, 1), M(0, 2), M(1, 1), M(1, 2), M(2, 2)};
    vo_eig3(a, m_w, m_V);
  }
  T get_eigenvalue(int i) const { return m_w[i]; }
  vnl_vector<T> get_eigenvector(int i) const
  {
    vnl_vector<T> v(3);
    for (int r = 0; r < 3; ++r) v(r) = m_V[r * 3 + i];
    return v;
  }
private:
  double m_w[3], m_V[9];
};

#endif  // MINI_ITK_VED_H
