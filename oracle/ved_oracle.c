/*
 * ved_oracle.c -- CPU restatement (double precision, single thread) of the tensor front-end of
 * itk::VEDMultigridImageFilter: Hessian at several scales, vesselness, diffusion-tensor synthesis
 * (SURVEY.md section 8f, ranks 1 and 2 -- the caller step immediately before the multigrid solve).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's CPU
 * legs may load it.  The product path (libmadgpu.so) never calls into this file.
 *
 * What is pinned and what is not
 * ------------------------------
 *  - vo_vesselness / vo_update_vesselness / vo_generate_tensor restate the REFERENCE'S OWN CODE
 *    (/root/reference/include/itkVEDMultigridImageFilter.hxx:176-378).  They are pinned against that code:
 *    oracle/_ref/libmadref.so compiles the unmodified itkVEDMultigridImageFilter.{h,hxx} against the
 *    stand-in ITK of oracle/shim, and tests/test_oracle_vs_ref.py compares response, eigen-system and
 *    tensor bit for bit / to rounding.
 *  - vo_rg_* / vo_hessian restate THIRD-PARTY code that is absent from /root/reference:
 *    itk::HessianRecursiveGaussianImageFilter -> itk::RecursiveGaussianImageFilter ->
 *    itk::RecursiveSeparableImageFilter (ITK 4.x, version not pinned by the reference; call site
 *    itkVEDMultigridImageFilter.hxx:164-171).  The published algorithm (Deriche's recursive Gaussian in the
 *    4th-order form with the Farneback-Westin style normalisation ITK documents) is restated from its
 *    description.  PARITY UNPINNED: no reference test or fixture holds an output of it.  It is cross-checked
 *    against sampled-Gaussian convolution (scipy) in tests/test_cpu_ved.py, which bounds restatement errors
 *    but is not the same filter.
 *  - vo_eig3 stands in for vnl_symmetric_eigensystem<double> (VXL, EISPACK rs; call site .hxx:259-264):
 *    ascending eigenvalues, eigenvectors as columns.  Any exact symmetric eigen-solver agrees to rounding
 *    except for the sign of a vector (irrelevant: the tensor is Q D Q^T) and the basis of a degenerate
 *    eigen-space.  PARITY UNPINNED (benign), checked against LAPACK (numpy.linalg.eigh).
 *
 * Arrays are x fastest (ITK index[0] contiguous); tensors / Hessians are the ITK AoS buffer, six scalars
 * per voxel in the order (0,0),(0,1),(0,2),(1,1),(1,2),(2,2).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * Recursive Gaussian, one axis.  itk::RecursiveGaussianImageFilter::SetUp + ComputeNCoefficients /
 * ComputeDCoefficients / ComputeRemainingCoefficients (third-party, see header).
 * order: 0 smoothing, 1 first derivative, 2 second derivative.  sigma in physical units; the filter
 * works on the sample grid with sigmad = sigma / spacing and, with normalize_across_scale, multiplies the
 * derivative of order k by sigma^k (physical sigma: the caller divides by spacing^k afterwards,
 * HessianRecursiveGaussianImageFilter).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  double N0, N1, N2, N3;
  double D1, D2, D3, D4;
  double M1, M2, M3, M4;
  double BN1, BN2, BN3, BN4;
  double BM1, BM2, BM3, BM4;
} vo_rg_coefs;

static void rg_n_coefs(double sigmad, double A1, double B1, double W1, double L1, double A2, double B2, double W2, double L2,
                       double *N0, double *N1, double *N2, double *N3, double *SN, double *DN, double *EN)
{
  const double Sin1 = sin(W1 / sigmad), Sin2 = sin(W2 / sigmad);
  const double Cos1 = cos(W1 / sigmad), Cos2 = cos(W2 / sigmad);
  const double Exp1 = exp(L1 / sigmad), Exp2 = exp(L2 / sigmad);
  *N0 = A1 + A2;
  *N1 = Exp2 * (B2 * Sin2 - (A2 + 2 * A1) * Cos2);
  *N1 += Exp1 * (B1 * Sin1 - (A1 + 2 * A2) * Cos1);
  *N2 = (A1 + A2) * Cos2 * Cos1;
  *N2 -= B1 * Cos2 * Sin1 + B2 * Cos1 * Sin2;
  *N2 *= 2 * Exp1 * Exp2;
  *N2 += A2 * Exp1 * Exp1 + A1 * Exp2 * Exp2;
  *N3 = Exp2 * Exp1 * Exp1 * (B2 * Sin2 - A2 * Cos2);
  *N3 += Exp1 * Exp2 * Exp2 * (B1 * Sin1 - A1 * Cos1);
  *SN = *N0 + *N1 + *N2 + *N3;
  *DN = *N1 + 2 * *N2 + 3 * *N3;
  *EN = *N1 + 4 * *N2 + 9 * *N3;
}

static void rg_d_coefs(double sigmad, double W1, double L1, double W2, double L2, vo_rg_coefs *c, double *SD, double *DD, double *ED)
{
  const double Cos1 = cos(W1 / sigmad), Cos2 = cos(W2 / sigmad);
  const double Exp1 = exp(L1 / sigmad), Exp2 = exp(L2 / sigmad);
  c->D4 = Exp1 * Exp1 * Exp2 * Exp2;
  c->D3 = -2 * Cos1 * Exp1 * Exp2 * Exp2;
  c->D3 += -2 * Cos2 * Exp2 * Exp1 * Exp1;
  c->D2 = 4 * Cos2 * Cos1 * Exp1 * Exp2;
  c->D2 += Exp1 * Exp1 + Exp2 * Exp2;
  c->D1 = -2 * (Exp2 * Cos2 + Exp1 * Cos1);
  *SD = 1.0 + c->D1 + c->D2 + c->D3 + c->D4;
  *DD = c->D1 + 2 * c->D2 + 3 * c->D3 + 4 * c->D4;
  *ED = c->D1 + 4 * c->D2 + 9 * c->D3 + 16 * c->D4;
}

static void rg_remaining(vo_rg_coefs *c, int symmetric)
{
  if (symmetric) {
    c->M1 = c->N1 - c->D1 * c->N0;
    c->M2 = c->N2 - c->D2 * c->N0;
    c->M3 = c->N3 - c->D3 * c->N0;
    c->M4 = -c->D4 * c->N0;
  } else {
    c->M1 = -(c->N1 - c->D1 * c->N0);
    c->M2 = -(c->N2 - c->D2 * c->N0);
    c->M3 = -(c->N3 - c->D3 * c->N0);
    c->M4 = c->D4 * c->N0;
  }
  /* boundary coefficients: the border sample is assumed to extend to infinity */
  const double SN = c->N0 + c->N1 + c->N2 + c->N3;
  const double SM = c->M1 + c->M2 + c->M3 + c->M4;
  const double SD = 1.0 + c->D1 + c->D2 + c->D3 + c->D4;
  c->BN1 = c->D1 * SN / SD; c->BN2 = c->D2 * SN / SD; c->BN3 = c->D3 * SN / SD; c->BN4 = c->D4 * SN / SD;
  c->BM1 = c->D1 * SM / SD; c->BM2 = c->D2 * SM / SD; c->BM3 = c->D3 * SM / SD; c->BM4 = c->D4 * SM / SD;
}

void vo_rg_setup(double sigma, double spacing, int order, int normalize_across_scale, vo_rg_coefs *c)
{
  if (spacing < 0.0) spacing = -spacing;
  const double sigmad = sigma / spacing;
  double across = 1.0;
  /* Deriche's parameters for the 4th-order approximation of the Gaussian and its derivatives */
  const double W1 = 0.6681, L1 = -1.3932, W2 = 2.0787, L2 = -1.3732;
  const double A1[3] = {1.3530, -0.6724, -1.3563};
  const double B1[3] = {1.8151, -3.4327, 5.2318};
  const double A2[3] = {-0.3531, 0.6724, 0.3446};
  const double B2[3] = {0.0902, 0.6100, -2.2355};
  double SD, DD, ED, SN, DN, EN;
  rg_d_coefs(sigmad, W1, L1, W2, L2, c, &SD, &DD, &ED);
  if (order == 0) {
    rg_n_coefs(sigmad, A1[0], B1[0], W1, L1, A2[0], B2[0], W2, L2, &c->N0, &c->N1, &c->N2, &c->N3, &SN, &DN, &EN);
    const double alpha0 = 2 * SN / SD - c->N0;
    c->N0 *= across / alpha0; c->N1 *= across / alpha0; c->N2 *= across / alpha0; c->N3 *= across / alpha0;
    rg_remaining(c, 1);
  } else if (order == 1) {
    if (normalize_across_scale) across = sigma;
    rg_n_coefs(sigmad, A1[1], B1[1], W1, L1, A2[1], B2[1], W2, L2, &c->N0, &c->N1, &c->N2, &c->N3, &SN, &DN, &EN);
    const double alpha1 = 2 * (SN * DD - DN * SD) / (SD * SD);
    c->N0 *= across / alpha1; c->N1 *= across / alpha1; c->N2 *= across / alpha1; c->N3 *= across / alpha1;
    rg_remaining(c, 0);
  } else {
    if (normalize_across_scale) across = sigma * sigma;
    double N0_0, N1_0, N2_0, N3_0, N0_2, N1_2, N2_2, N3_2, SN0, DN0, EN0, SN2, DN2, EN2;
    rg_n_coefs(sigmad, A1[0], B1[0], W1, L1, A2[0], B2[0], W2, L2, &N0_0, &N1_0, &N2_0, &N3_0, &SN0, &DN0, &EN0);
    rg_n_coefs(sigmad, A1[2], B1[2], W1, L1, A2[2], B2[2], W2, L2, &N0_2, &N1_2, &N2_2, &N3_2, &SN2, &DN2, &EN2);
    /* the second-order kernel gets a multiple of the smoothing kernel added so that its DC response vanishes */
    const double beta = -(2 * SN2 - SD * N0_2) / (2 * SN0 - SD * N0_0);
    c->N0 = N0_2 + beta * N0_0; c->N1 = N1_2 + beta * N1_0; c->N2 = N2_2 + beta * N2_0; c->N3 = N3_2 + beta * N3_0;
    SN = SN2 + beta * SN0; DN = DN2 + beta * DN0; EN = EN2 + beta * EN0;
    const double alpha2 = (EN * SD * SD - ED * SN * SD - 2 * DN * DD * SD + 2 * DD * DD * SN) / (SD * SD * SD);
    c->N0 *= across / alpha2; c->N1 *= across / alpha2; c->N2 *= across / alpha2; c->N3 *= across / alpha2;
    rg_remaining(c, 1);
  }
}

/* itk::RecursiveSeparableImageFilter::FilterDataArray: causal + anti-causal pass over one line of ln >= 4 samples.
 * scratch: 2 * ln doubles. */
void vo_rg_filter_line(const vo_rg_coefs *c, const double *data, double *outs, double *scratch, int ln)
{
  double *s1 = scratch, *s2 = scratch + ln;
  const double v1 = data[0];
  s1[0] = v1 * c->N0 + v1 * c->N1 + v1 * c->N2 + v1 * c->N3;
  s1[1] = data[1] * c->N0 + v1 * c->N1 + v1 * c->N2 + v1 * c->N3;
  s1[2] = data[2] * c->N0 + data[1] * c->N1 + v1 * c->N2 + v1 * c->N3;
  s1[3] = data[3] * c->N0 + data[2] * c->N1 + data[1] * c->N2 + v1 * c->N3;
  s1[0] -= v1 * c->BN1 + v1 * c->BN2 + v1 * c->BN3 + v1 * c->BN4;
  s1[1] -= s1[0] * c->D1 + v1 * c->BN2 + v1 * c->BN3 + v1 * c->BN4;
  s1[2] -= s1[1] * c->D1 + s1[0] * c->D2 + v1 * c->BN3 + v1 * c->BN4;
  s1[3] -= s1[2] * c->D1 + s1[1] * c->D2 + s1[0] * c->D3 + v1 * c->BN4;
  for (int i = 4; i < ln; ++i) {
    s1[i] = data[i] * c->N0 + data[i - 1] * c->N1 + data[i - 2] * c->N2 + data[i - 3] * c->N3;
    s1[i] -= s1[i - 1] * c->D1 + s1[i - 2] * c->D2 + s1[i - 3] * c->D3 + s1[i - 4] * c->D4;
  }
  const double v2 = data[ln - 1];
  s2[ln - 1] = v2 * c->M1 + v2 * c->M2 + v2 * c->M3 + v2 * c->M4;
  s2[ln - 2] = data[ln - 1] * c->M1 + v2 * c->M2 + v2 * c->M3 + v2 * c->M4;
  s2[ln - 3] = data[ln - 2] * c->M1 + data[ln - 1] * c->M2 + v2 * c->M3 + v2 * c->M4;
  s2[ln - 4] = data[ln - 3] * c->M1 + data[ln - 2] * c->M2 + data[ln - 1] * c->M3 + v2 * c->M4;
  s2[ln - 1] -= v2 * c->BM1 + v2 * c->BM2 + v2 * c->BM3 + v2 * c->BM4;
  s2[ln - 2] -= s2[ln - 1] * c->D1 + v2 * c->BM2 + v2 * c->BM3 + v2 * c->BM4;
  s2[ln - 3] -= s2[ln - 2] * c->D1 + s2[ln - 1] * c->D2 + v2 * c->BM3 + v2 * c->BM4;
  s2[ln - 4] -= s2[ln - 3] * c->D1 + s2[ln - 2] * c->D2 + s2[ln - 1] * c->D3 + v2 * c->BM4;
  for (int i = ln - 4; i > 0; --i) {
    s2[i - 1] = data[i] * c->M1 + data[i + 1] * c->M2 + data[i + 2] * c->M3 + data[i + 3] * c->M4;
    s2[i - 1] -= s2[i] * c->D1 + s2[i + 1] * c->D2 + s2[i + 2] * c->D3 + s2[i + 3] * c->D4;
  }
  for (int i = 0; i < ln; ++i) outs[i] = s1[i] + s2[i];
}

/* one axis of a volume n = (nx, ny, nz); in and out may be the same array */
int vo_rg_filter_axis(const int *n, int axis, const vo_rg_coefs *c, const double *in, double *out)
{
  const int ln = n[axis];
  if (ln < 4) return -1; /* RecursiveSeparableImageFilter refuses lines shorter than 4 samples */
  const int64_t stride = axis == 0 ? 1 : (axis == 1 ? n[0] : (int64_t)n[0] * n[1]);
  double *line = (double *)malloc(sizeof(double) * 4 * (size_t)ln);
  if (!line) return -2;
  double *res = line + ln, *scratch = line + 2 * ln;
  const int na = axis == 0 ? n[1] : n[0], nb = axis == 2 ? n[1] : n[2];
  const int64_t sa = axis == 0 ? n[0] : 1, sb = axis == 2 ? n[0] : (int64_t)n[0] * n[1];
  for (int b = 0; b < nb; ++b)
    for (int a = 0; a < na; ++a) {
      const int64_t base = a * sa + b * sb;
      for (int i = 0; i < ln; ++i) line[i] = in[base + i * stride];
      vo_rg_filter_line(c, line, res, scratch, ln);
      for (int i = 0; i < ln; ++i) out[base + i * stride] = res[i];
    }
  free(line);
  return 0;
}

/* itk::HessianRecursiveGaussianImageFilter::GenerateData (third-party, see header), as configured by
 * VEDMultigridImageFilter::ComputeHessian (itkVEDMultigridImageFilter.hxx:158-173: NormalizeAcrossScale on).
 * For every pair dima <= dimb: second-order filter along dima (or first-order along dima and dimb), zero-order
 * smoothing along the remaining axes, result divided by spacing[dima] * spacing[dimb].
 * hessian_aos: 6 doubles per voxel. */
int vo_hessian(const int *n, const double *h, double sigma, int normalize_across_scale, const double *image, double *hessian_aos)
{
  const int64_t nvox = (int64_t)n[0] * n[1] * n[2];
  double *tmp = (double *)malloc(sizeof(double) * (size_t)nvox);
  if (!tmp) return -2;
  int comp = 0;
  for (int dima = 0; dima < 3; ++dima)
    for (int dimb = dima; dimb < 3; ++dimb, ++comp) {
      int order[3] = {0, 0, 0};
      if (dima == dimb) order[dima] = 2;
      else { order[dima] = 1; order[dimb] = 1; }
      /* derivative filter A runs first (along dima), then B, then the smoothing filters */
      int seq[3], k = 0;
      seq[k++] = dima;
      if (dimb != dima) seq[k++] = dimb;
      for (int d = 0; d < 3; ++d)
        if (d != dima && d != dimb) seq[k++] = d;
      const double *src = image;
      for (int s = 0; s < 3; ++s) {
        vo_rg_coefs c;
        vo_rg_setup(sigma, h[seq[s]], order[seq[s]], normalize_across_scale, &c);
        const int rc = vo_rg_filter_axis(n, seq[s], &c, src, tmp);
        if (rc) { free(tmp); return rc; }
        src = tmp;
      }
      const double factor = h[dima] * h[dimb];
      for (int64_t v = 0; v < nvox; ++v) hessian_aos[v * 6 + comp] = tmp[v] / factor;
    }
  free(tmp);
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Symmetric 3x3 eigen-system, stand-in for vnl_symmetric_eigensystem<double> (third-party, see header):
 * a = (xx, xy, xz, yy, yz, zz); w ascending; V row-major 3x3 with COLUMN k the unit eigenvector of w[k]
 * (get_eigenvector(k)(r) == V[r*3+k]).  Cyclic Jacobi rotations to machine precision.
 * ---------------------------------------------------------------------------------------------- */
void vo_eig3(const double *a, double *w, double *V)
{
  double A[3][3] = {{a[0], a[1], a[2]}, {a[1], a[3], a[4]}, {a[2], a[4], a[5]}};
  double Q[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 60; ++sweep) {
    const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
    if (off == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (A[p][q] == 0.0) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        const double app = A[p][p], aqq = A[q][q], apq = A[p][q];
        A[p][p] = app - t * apq;
        A[q][q] = aqq + t * apq;
        A[p][q] = A[q][p] = 0.0;
        const int r = 3 - p - q;
        const double arp = A[r][p], arq = A[r][q];
        A[r][p] = A[p][r] = c * arp - s * arq;
        A[r][q] = A[q][r] = s * arp + c * arq;
        for (int i = 0; i < 3; ++i) {
          const double qip = Q[i][p], qiq = Q[i][q];
          Q[i][p] = c * qip - s * qiq;
          Q[i][q] = s * qip + c * qiq;
        }
      }
  }
  int idx[3] = {0, 1, 2};
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2 - i; ++j)
      if (A[idx[j]][idx[j]] > A[idx[j + 1]][idx[j + 1]]) { int t = idx[j]; idx[j] = idx[j + 1]; idx[j + 1] = t; }
  for (int k = 0; k < 3; ++k) {
    w[k] = A[idx[k]][idx[k]];
    for (int r = 0; r < 3; ++r) V[r * 3 + k] = Q[r][idx[k]];
  }
}

/* ------------------------------------------------------------------------------------------------
 * VEDMultigridImageFilter::VesselnessFunction, itkVEDMultigridImageFilter.hxx:176-212.
 * e: eigenvalues sorted by increasing magnitude.  The reference's unqualified abs() on doubles resolves to the
 * <cmath> floating-point overload, i.e. fabs.
 * ---------------------------------------------------------------------------------------------- */
double vo_vesselness(const double *e, double alpha, double beta, double gamma)
{
  if (e[1] >= 0 || e[2] >= 0) return 0.0; /* :183-186 */
  const double smoothC = 1e-5;            /* :190 */
  const double alphaDen = 2.0 * alpha * alpha, betaDen = 2.0 * beta * beta, gammaDen = 2.0 * gamma * gamma; /* :192-194 */
  const double alphaNum = (e[1] * e[1]) / (e[2] * e[2]);             /* :196 */
  const double betaNum = (e[0] * e[0]) / fabs(e[1] * e[2]);          /* :197 */
  const double gammaNum = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];   /* :198-200 */
  const double smooth = exp(-(2 * smoothC * smoothC) / (fabs(e[1]) * e[2] * e[2])); /* :202-203 */
  return smooth * (1. - exp(-alphaNum / alphaDen)) * exp(-betaNum / betaDen) * (1. - exp(-gammaNum / gammaDen)); /* :205-207 */
}

/* VEDMultigridImageFilter::UpdateVesselness, itkVEDMultigridImageFilter.hxx:215-299.
 * first: m_MaxVesselnessResponse was null (:222) -- the response starts at 0 (:231) and every voxel is stored.
 * response: nvox; eigenvalues: 3 per voxel (sorted by magnitude, :266-268); eigenvectors: 9 per voxel, row-major
 * Matrix with column d = get_eigenvector(d) in the eigen-solver's ASCENDING order (:281-283) -- the two orders
 * are not the same, and GenerateDiffusionTensor relies on the second one. */
void vo_update_vesselness(int64_t nvox, const double *hessian_aos, int first, double alpha, double beta, double gamma,
                          double *response, double *eigenvalues, double *eigenvectors)
{
  if (first)
    for (int64_t v = 0; v < nvox; ++v) response[v] = 0.0;
  for (int64_t v = 0; v < nvox; ++v) {
    double w[3], V[9], e[3];
    vo_eig3(hessian_aos + v * 6, w, V); /* :259-264 */
    e[0] = w[0]; e[1] = w[1]; e[2] = w[2];
    double t;
    if (fabs(e[0]) > fabs(e[1])) { t = e[0]; e[0] = e[1]; e[1] = t; } /* :266 */
    if (fabs(e[1]) > fabs(e[2])) { t = e[1]; e[1] = e[2]; e[2] = t; } /* :267 */
    if (fabs(e[0]) > fabs(e[1])) { t = e[0]; e[0] = e[1]; e[1] = t; } /* :268 */
    const double vess = vo_vesselness(e, alpha, beta, gamma); /* :270 */
    if (first || vess > response[v]) {                        /* :272 */
      for (int d = 0; d < 3; ++d) {
        eigenvalues[v * 3 + d] = e[d];
        for (int d2 = 0; d2 < 3; ++d2) eigenvectors[v * 9 + d2 * 3 + d] = V[d2 * 3 + d];
      }
      response[v] = vess;
    }
  }
}

/* VEDMultigridImageFilter::GenerateDiffusionTensor, itkVEDMultigridImageFilter.hxx:302-378. */
void vo_generate_tensor(int64_t nvox, const double *response, const double *eigenvectors, double sensitivity, double epsilon,
                        double omega, double *tensor_aos)
{
  for (int64_t v = 0; v < nvox; ++v) {
    double *T = tensor_aos + v * 6;
    const double Vs = pow(response[v], 1. / sensitivity); /* :327 */
    if (Vs > 0) {
      const double *Q = eigenvectors + v * 9;
      double D[3];
      for (int d = 0; d < 3; ++d) D[d] = d == 2 ? 1. + (omega - 1.) * Vs : 1. + (epsilon - 1.) * Vs; /* :336-337 */
      double temp[9], Tm[9]; /* temp = Q * D; T = temp * Qt (:343-346) */
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
          double s = 0;
          for (int k = 0; k < 3; ++k) s += Q[r * 3 + k] * (k == c ? D[c] : 0.0);
          temp[r * 3 + c] = s;
        }
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
          double s = 0;
          for (int k = 0; k < 3; ++k) s += temp[r * 3 + k] * Q[c * 3 + k];
          Tm[r * 3 + c] = s;
        }
      int comp = 0;
      for (int d = 0; d < 3; ++d)
        for (int d2 = d; d2 < 3; ++d2) T[comp++] = Tm[d * 3 + d2]; /* :348-354 */
    } else {
      T[0] = 1; T[1] = 0; T[2] = 0; T[3] = 1; T[4] = 0; T[5] = 1; /* :357-366 */
    }
  }
}
