// ved_math.h -- per-voxel and per-line arithmetic of the VED tensor front-end, shared between the
// CUDA kernels of ved.cu (device) and the host harness tests/ved_host_harness.cpp, which runs exactly these functions on the
// CPU against the oracle in the `-m "not gpu"` suite.  No CUDA runtime calls in here.
//
// Reference (paths under /root/reference/include):
//   vesselness()        VEDMultigridImageFilter::VesselnessFunction   itkVEDMultigridImageFilter.hxx:176-212
//   update_voxel()      ::UpdateVesselness (per voxel)                 :246-297   + ::GenerateDiffusionTensor :327-366
//   eig3_top()          vnl_symmetric_eigensystem<double>              (third-party, call site :259-264)
//   rg_*                itk::RecursiveGaussianImageFilter / RecursiveSeparableImageFilter (third-party, call site :164-171)
#ifndef MADGPU_VED_MATH_H
#define MADGPU_VED_MATH_H

#include <math.h>

#if defined(__CUDACC__)
#define VED_HD __host__ __device__ __forceinline__
#else
#define VED_HD inline
#endif
#if defined(__CUDA_ARCH__)  // unroll hints for the device pass only (host compilers warn about the pragma)
#define VED_UNROLL _Pragma("unroll")
#define VED_UNROLL4 _Pragma("unroll 4")
#else
#define VED_UNROLL
#define VED_UNROLL4
#endif

namespace ved
{
struct Params {  // itkSetMacro setters, itkVEDMultigridImageFilter.h:88-93
  double alpha, beta, gamma, epsilon, omega, sensitivity;
};

// ---------------------------------------------------------------------------------------------------------------------
// Recursive Gaussian of one axis: y = causal(x) + anticausal(x), 4th-order recursions (Deriche) with ITK's normalisation.
// ---------------------------------------------------------------------------------------------------------------------
struct RgCoefs {
  double N0, N1, N2, N3;  // causal numerator
  double D1, D2, D3, D4;  // common denominator
  double M1, M2, M3, M4;  // anticausal numerator
  double cdc, adc;        // response of the causal / anticausal recursion to a constant 1 (SN/SD, SM/SD)
};

namespace detail
{
inline void rg_n(double sd, double A1, double B1, double W1, double L1, double A2, double B2, double W2, double L2, double n[4], double& SN,
                 double& DN, double& EN)
{
  const double s1 = sin(W1 / sd), s2 = sin(W2 / sd), c1 = cos(W1 / sd), c2 = cos(W2 / sd), e1 = exp(L1 / sd), e2 = exp(L2 / sd);
  n[0] = A1 + A2;
  n[1] = e2 * (B2 * s2 - (A2 + 2 * A1) * c2) + e1 * (B1 * s1 - (A1 + 2 * A2) * c1);
  n[2] = 2 * e1 * e2 * ((A1 + A2) * c2 * c1 - (B1 * c2 * s1 + B2 * c1 * s2)) + A2 * e1 * e1 + A1 * e2 * e2;
  n[3] = e2 * e1 * e1 * (B2 * s2 - A2 * c2) + e1 * e2 * e2 * (B1 * s1 - A1 * c1);
  SN = n[0] + n[1] + n[2] + n[3];
  DN = n[1] + 2 * n[2] + 3 * n[3];
  EN = n[1] + 4 * n[2] + 9 * n[3];
}
}  // namespace detail

// Host-side coefficient set-up (a plain host function: called once per pass, the kernels get the result by value).
// order 0/1/2 = smoothing / first / second derivative; sigma in physical units, the recursion runs on the sample grid with sigma / spacing; with normalize_across_scale the derivative of order k carries sigma^k
// (physical; the Hessian divides by the spacings afterwards).
inline void rg_setup(double sigma, double spacing, int order, bool normalize_across_scale, RgCoefs& c)
{
  const double sd = sigma / fabs(spacing);
  const double W1 = 0.6681, L1 = -1.3932, W2 = 2.0787, L2 = -1.3732;
  const double A1[3] = {1.3530, -0.6724, -1.3563}, B1[3] = {1.8151, -3.4327, 5.2318};
  const double A2[3] = {-0.3531, 0.6724, 0.3446}, B2[3] = {0.0902, 0.6100, -2.2355};
  const double c1 = cos(W1 / sd), c2 = cos(W2 / sd), e1 = exp(L1 / sd), e2 = exp(L2 / sd);
  c.D1 = -2 * (e2 * c2 + e1 * c1);
  c.D2 = 4 * c2 * c1 * e1 * e2 + e1 * e1 + e2 * e2;
  c.D3 = -2 * c1 * e1 * e2 * e2 - 2 * c2 * e2 * e1 * e1;
  c.D4 = e1 * e1 * e2 * e2;
  const double SD = 1.0 + c.D1 + c.D2 + c.D3 + c.D4;
  const double DD = c.D1 + 2 * c.D2 + 3 * c.D3 + 4 * c.D4;
  const double ED = c.D1 + 4 * c.D2 + 9 * c.D3 + 16 * c.D4;
  double n[4], SN, DN, EN, norm;
  bool symmetric = true;
  if (order == 0) {
    detail::rg_n(sd, A1[0], B1[0], W1, L1, A2[0], B2[0], W2, L2, n, SN, DN, EN);
    norm = 1.0 / (2 * SN / SD - n[0]);
  } else if (order == 1) {
    detail::rg_n(sd, A1[1], B1[1], W1, L1, A2[1], B2[1], W2, L2, n, SN, DN, EN);
    norm = (normalize_across_scale ? sigma : 1.0) / (2 * (SN * DD - DN * SD) / (SD * SD));
    symmetric = false;
  } else {
    double n0[4], n2[4], SN0, DN0, EN0, SN2, DN2, EN2;
    detail::rg_n(sd, A1[0], B1[0], W1, L1, A2[0], B2[0], W2, L2, n0, SN0, DN0, EN0);
    detail::rg_n(sd, A1[2], B1[2], W1, L1, A2[2], B2[2], W2, L2, n2, SN2, DN2, EN2);
    const double beta = -(2 * SN2 - SD * n2[0]) / (2 * SN0 - SD * n0[0]);  // cancels the DC response
    for (int i = 0; i < 4; ++i) n[i] = n2[i] + beta * n0[i];
    SN = SN2 + beta * SN0; DN = DN2 + beta * DN0; EN = EN2 + beta * EN0;
    norm = (normalize_across_scale ? sigma * sigma : 1.0) /
           ((EN * SD * SD - ED * SN * SD - 2 * DN * DD * SD + 2 * DD * DD * SN) / (SD * SD * SD));
  }
  c.N0 = n[0] * norm; c.N1 = n[1] * norm; c.N2 = n[2] * norm; c.N3 = n[3] * norm;
  const double sgn = symmetric ? 1.0 : -1.0;
  c.M1 = sgn * (c.N1 - c.D1 * c.N0);
  c.M2 = sgn * (c.N2 - c.D2 * c.N0);
  c.M3 = sgn * (c.N3 - c.D3 * c.N0);
  c.M4 = -sgn * c.D4 * c.N0;
  c.cdc = (c.N0 + c.N1 + c.N2 + c.N3) / SD;
  c.adc = (c.M1 + c.M2 + c.M3 + c.M4) / SD;
}

// Recursion state: the last four inputs and outputs.  The border sample is taken to extend to infinity, so the recursion
// starts in its steady state for that constant (ITK's boundary coefficients BN / BM say the same thing).
struct RgState {
  double x1, x2, x3, x4;
  double y1, y2, y3, y4;
};

VED_HD void rg_causal_init(RgState& s, const RgCoefs& c, double edge)
{
  s.x1 = s.x2 = s.x3 = s.x4 = edge;
  s.y1 = s.y2 = s.y3 = s.y4 = edge * c.cdc;
}

// y[i] = N0 x[i] + N1 x[i-1] + N2 x[i-2] + N3 x[i-3] - (D1 y[i-1] + D2 y[i-2] + D3 y[i-3] + D4 y[i-4]); call with i ascending
VED_HD double rg_causal_step(RgState& s, const RgCoefs& c, double x)
{
  const double y = (c.N0 * x + c.N1 * s.x1 + c.N2 * s.x2 + c.N3 * s.x3) - (c.D1 * s.y1 + c.D2 * s.y2 + c.D3 * s.y3 + c.D4 * s.y4);
  s.x3 = s.x2; s.x2 = s.x1; s.x1 = x;
  s.y4 = s.y3; s.y3 = s.y2; s.y2 = s.y1; s.y1 = y;
  return y;
}

VED_HD void rg_anti_init(RgState& s, const RgCoefs& c, double edge)
{
  s.x1 = s.x2 = s.x3 = s.x4 = edge;
  s.y1 = s.y2 = s.y3 = s.y4 = edge * c.adc;
}

// a[i] = M1 x[i+1] + M2 x[i+2] + M3 x[i+3] + M4 x[i+4] - (D1 a[i+1] + ... + D4 a[i+4]); call with i descending, passing x[i]
VED_HD double rg_anti_step(RgState& s, const RgCoefs& c, double x)
{
  const double y = (c.M1 * s.x1 + c.M2 * s.x2 + c.M3 * s.x3 + c.M4 * s.x4) - (c.D1 * s.y1 + c.D2 * s.y2 + c.D3 * s.y3 + c.D4 * s.y4);
  s.x4 = s.x3; s.x3 = s.x2; s.x2 = s.x1; s.x1 = x;
  s.y4 = s.y3; s.y3 = s.y2; s.y2 = s.y1; s.y1 = y;
  return y;
}

// One whole line, K filters of the same input: x[i * stride] -> out[k][i * stride] (fp32 storage, fp64 recursion).  The causal
// pass stores its result, the anticausal pass adds to it and applies scale[k].  in must not alias any out[k].
// This is the body of k_rg_lines (ved.cu); k_rg_rows runs the same sequence of steps per row through shared-memory tiles.
// Loads are batched RG_BATCH elements ahead of the recursion: the outputs may alias the input as far as the compiler can tell, so a
// load placed after a store has to wait for it -- one memory latency per element, which bounded k_rg_lines at 1.2-1.4 TB/s.
constexpr int RG_BATCH = 8;
template <int K>
VED_HD void rg_line(const float* __restrict__ x, long long stride, int n, const RgCoefs* c, float* const* out, const double* scale)
{
  RgState s[K];
  const double e0 = (double)x[0];
VED_UNROLL
  for (int k = 0; k < K; ++k) rg_causal_init(s[k], c[k], e0);
  for (int i0 = 0; i0 < n; i0 += RG_BATCH) {
    float xb[RG_BATCH];
VED_UNROLL
    for (int j = 0; j < RG_BATCH; ++j) xb[j] = i0 + j < n ? x[(long long)(i0 + j) * stride] : 0.f;
VED_UNROLL
    for (int j = 0; j < RG_BATCH; ++j) {
      if (i0 + j < n) {
        const double xi = (double)xb[j];
VED_UNROLL
        for (int k = 0; k < K; ++k) out[k][(long long)(i0 + j) * stride] = (float)rg_causal_step(s[k], c[k], xi);
      }
    }
  }
  const double e1 = (double)x[(long long)(n - 1) * stride];
VED_UNROLL
  for (int k = 0; k < K; ++k) rg_anti_init(s[k], c[k], e1);
  for (int i0 = n - 1; i0 >= 0; i0 -= RG_BATCH) {
    float xb[RG_BATCH], ob[K][RG_BATCH];
VED_UNROLL
    for (int j = 0; j < RG_BATCH; ++j) {
      const bool in = i0 - j >= 0;
      xb[j] = in ? x[(long long)(i0 - j) * stride] : 0.f;
VED_UNROLL
      for (int k = 0; k < K; ++k) ob[k][j] = in ? out[k][(long long)(i0 - j) * stride] : 0.f;
    }
VED_UNROLL
    for (int j = 0; j < RG_BATCH; ++j) {
      if (i0 - j >= 0) {
        const double xi = (double)xb[j];
VED_UNROLL
        for (int k = 0; k < K; ++k) out[k][(long long)(i0 - j) * stride] = (float)(((double)ob[k][j] + rg_anti_step(s[k], c[k], xi)) * scale[k]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Symmetric 3x3 eigen-problem by cyclic Jacobi rotations (double).  h = (xx, xy, xz, yy, yz, zz).
// w: eigenvalues in ASCENDING order (the contract of vnl_symmetric_eigensystem); t: unit eigenvector of w[2].
// ---------------------------------------------------------------------------------------------------------------------
namespace detail
{
// one rotation in the (p, q) plane: app, aqq, apq the 2x2 block, arp / arq the couplings to the third index,
// (q0p, q0q), ... the two affected columns of the eigenvector matrix
VED_HD void jacobi_rotate(double& app, double& aqq, double& apq, double& arp, double& arq, double& q0p, double& q0q, double& q1p, double& q1q,
                          double& q2p, double& q2q)
{
  if (apq == 0.0) return;
  const double theta = (aqq - app) / (2.0 * apq);
  const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
  const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
  app -= t * apq;
  aqq += t * apq;
  apq = 0.0;
  const double rp = arp, rq = arq;
  arp = c * rp - s * rq;
  arq = s * rp + c * rq;
  double a = q0p, b = q0q;
  q0p = c * a - s * b; q0q = s * a + c * b;
  a = q1p; b = q1q;
  q1p = c * a - s * b; q1q = s * a + c * b;
  a = q2p; b = q2q;
  q2p = c * a - s * b; q2q = s * a + c * b;
}
}  // namespace detail

VED_HD void eig3_top(const double h[6], double w[3], double t[3])
{
  double a00 = h[0], a01 = h[1], a02 = h[2], a11 = h[3], a12 = h[4], a22 = h[5];
  double q00 = 1, q01 = 0, q02 = 0, q10 = 0, q11 = 1, q12 = 0, q20 = 0, q21 = 0, q22 = 1;
  const double frob = a00 * a00 + a11 * a11 + a22 * a22 + 2.0 * (a01 * a01 + a02 * a02 + a12 * a12);  // rotation invariant
  for (int sweep = 0; sweep < 16; ++sweep) {
    const double off = a01 * a01 + a02 * a02 + a12 * a12;
    if (off <= 1e-36 * frob) break;  // off-diagonal norm below 1e-18 of the matrix norm: converged to the last bit
    detail::jacobi_rotate(a00, a11, a01, a02, a12, q00, q01, q10, q11, q20, q21);  // (0,1), third index 2
    detail::jacobi_rotate(a00, a22, a02, a01, a12, q00, q02, q10, q12, q20, q22);  // (0,2), third index 1
    detail::jacobi_rotate(a11, a22, a12, a01, a02, q01, q02, q11, q12, q21, q22);  // (1,2), third index 0
  }
  // ascending order; the eigenvector of the largest eigenvalue
  double lo = a00, mid = a11, hi = a22;
  double t0 = q02, t1 = q12, t2 = q22;
  if (lo > hi) { const double x = lo; lo = hi; hi = x; t0 = q00; t1 = q10; t2 = q20; }  // hi <- a00
  if (mid > hi) {                                                                        // hi <- a11
    const double x = mid; mid = hi; hi = x;
    t0 = q01; t1 = q11; t2 = q21;
  }
  if (lo > mid) { const double x = lo; lo = mid; mid = x; }
  w[0] = lo; w[1] = mid; w[2] = hi;
  t[0] = t0; t[1] = t1; t[2] = t2;
}

// VesselnessFunction, itkVEDMultigridImageFilter.hxx:176-212; e sorted by increasing magnitude.
VED_HD double vesselness(const double e[3], const Params& P)
{
  if (e[1] >= 0.0 || e[2] >= 0.0) return 0.0;  // :183-186
  const double smoothC = 1e-5;                 // :190
  const double alphaNum = (e[1] * e[1]) / (e[2] * e[2]);            // :196
  const double betaNum = (e[0] * e[0]) / fabs(e[1] * e[2]);         // :197
  const double gammaNum = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];  // :198-200
  const double smooth = exp(-(2 * smoothC * smoothC) / (fabs(e[1]) * e[2] * e[2]));  // :202-203
  return smooth * (1. - exp(-alphaNum / (2.0 * P.alpha * P.alpha))) * exp(-betaNum / (2.0 * P.beta * P.beta)) *
         (1. - exp(-gammaNum / (2.0 * P.gamma * P.gamma)));  // :205-207
}

// One voxel of UpdateVesselness (:246-297) with GenerateDiffusionTensor (:327-366) folded in: the reference keeps the
// eigen-system of the best scale and builds T = Q D Q^T afterwards, with D = diag(1+(eps-1)V, 1+(eps-1)V, 1+(omega-1)V) in the
// eigen-solver's ascending order (the magnitude sort of :266-268 does NOT reorder the vectors), V = response^(1/sensitivity).
// Q is orthonormal, so T = a I + (b - a) t t^T with t the eigenvector of the LARGEST eigenvalue: only t is needed, and the
// tensor of the best scale so far can be written straight away.  first: no response stored yet (:222, :272).
// Returns true when (response, T) were replaced.
VED_HD bool update_voxel(const double h[6], bool first, const Params& P, double& response, double T[6])
{
  double w[3], t[3];
  eig3_top(h, w, t);  // :259-264
  double e[3] = {w[0], w[1], w[2]};
  double x;
  if (fabs(e[0]) > fabs(e[1])) { x = e[0]; e[0] = e[1]; e[1] = x; }  // :266
  if (fabs(e[1]) > fabs(e[2])) { x = e[1]; e[1] = e[2]; e[2] = x; }  // :267
  if (fabs(e[0]) > fabs(e[1])) { x = e[0]; e[0] = e[1]; e[1] = x; }  // :268
  const double v = vesselness(e, P);  // :270
  if (!(first || v > response)) return false;  // :272
  response = v;
  const double V = pow(v, 1.0 / P.sensitivity);  // :327
  if (V > 0.0) {
    const double a = 1.0 + (P.epsilon - 1.0) * V, b = 1.0 + (P.omega - 1.0) * V;  // :336-337
    const double d = b - a;
    T[0] = a + d * t[0] * t[0]; T[1] = d * t[0] * t[1]; T[2] = d * t[0] * t[2];
    T[3] = a + d * t[1] * t[1]; T[4] = d * t[1] * t[2];
    T[5] = a + d * t[2] * t[2];
  } else {  // :357-366
    T[0] = 1.0; T[1] = 0.0; T[2] = 0.0; T[3] = 1.0; T[4] = 0.0; T[5] = 1.0;
  }
  return true;
}
}  // namespace ved

#endif  // MADGPU_VED_MATH_H
