/*
 * mad_oracle.c -- CPU restatement (double precision, single thread) of the multigrid
 * anisotropic-diffusion solve of nellogrb/MultigridAnisotropicDiffusion.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (libmadgpu.so) never calls into this file.
 *
 * Parity pin: the reference ships no golden vectors (its tests return EXIT_SUCCESS
 * unconditionally, test/itk2DDiffusionTest_WJ.cxx:151), so this restatement is pinned against
 * the reference's OWN CODE: oracle/_ref/libmadref.so is the unmodified /root/reference/include
 * headers compiled against the stand-in ITK/vnl of oracle/shim (oracle/Makefile, `make ref`).
 *   - tests/test_oracle_vs_ref.py compares every routine and whole GenerateData() runs: operator
 *     rows, smoothers, residual, transfers and direct solve agree bit for bit; whole solves agree in
 *     cycle counts, per-cycle relative residuals and image (<= 1e-13).
 *   - tests/test_cpu_golden.py compares with vectors recorded from that library
 *     (tests/golden/make_golden.py) for the reference's three test programs, so the pin also holds
 *     where oracle/_ref cannot be built (no /root/reference).
 * The one third-party routine, vnl_sparse_lu, is an exact solve; here and in the shim it is a dense LU.
 *
 * Every function cites the reference file:line it follows.  Paths are relative to
 * /root/reference/include.  Arrays are x-fastest (ITK index[0] contiguous).
 *
 * The restatement keeps the reference's data structure on purpose: an explicit
 * 3^dim-entry stencil per voxel in Neighborhood raster order (x fastest), filled by
 * the same "+=" sequence with redirected (mirrored) offsets, traversed by the
 * smoothers through the active-offset list with IsInside guards.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MO_MAXLEV 32

typedef struct {
  int n[3];          /* size per axis (n[2]=1 in 2-D)                              */
  double h[3];       /* spacing                                                    */
  int centering[3];  /* how this level was obtained from the finer one: 0 vertex, 1 cell */
  int64_t nvox;
  double *stencil;   /* nvox * ns doubles, ns = 3^dim, raster order                */
  double *tensor;    /* ncomp planes (SoA), comp order (0,0),(0,1),(0,2),(1,1),(1,2),(2,2) */
} mo_level;

typedef struct {
  int dim, ns, ncomp, nlevels;
  double dt;
  mo_level lv[MO_MAXLEV];
  /* coarsest-grid direct solver: dense LU with partial pivoting */
  int nc;
  double *lu;
  int *piv;
  /* active offset list (Neighborhood raster order) */
  int nactive;
  int act_idx[27];
  int act_off[27][3];
  /* parameters of the filter (mad .../itkMultigridAnisotropicDiffusionImageFilter.hxx:38-49) */
  int smoother;  /* 0 = Gauss-Seidel (lexicographic), 1 = weighted Jacobi */
  double omega;
  int nu;
  int verbose;
  /* statistics */
  int64_t n_smooth, n_resid;
} mo_hier;

static int tcomp(int dim, int a, int b)
{
  /* SymmetricSecondRankTensor upper-triangular row-major storage */
  if (a > b) { int t = a; a = b; b = t; }
  if (dim == 2) return a == 0 ? b : 2;
  return a == 0 ? b : (a == 1 ? 2 + b : 5);
}

/* ---------------------------------------------------------------------------------
 * Level schedule: mad/itkGridsHierarchy.hxx:36-106.
 * Halve every axis (even -> n/2 "cell", odd -> (n-1)/2+1 "vertex") until any axis
 * drops below 6; the level on which that happens is discarded (:57).
 * ------------------------------------------------------------------------------- */
int mo_level_schedule(int dim, const int *n0, int *sizes /*[MO_MAXLEV][3]*/, int *centering /*[MO_MAXLEV][3]*/)
{
  long g[3] = {1, 1, 1};
  for (int d = 0; d < dim; ++d) g[d] = n0[d];
  int coarsest = 0, nlev = 1;
  while (!coarsest) {
    for (int d = 0; d < dim; ++d) {
      g[d] = (g[d] % 2 == 0) ? g[d] / 2 : ((g[d] - 1) / 2) + 1;
      if (g[d] < 6) coarsest = 1;
    }
    ++nlev;
  }
  --nlev;
  if (nlev > MO_MAXLEV) return -1;
  for (int d = 0; d < 3; ++d) { sizes[d] = d < dim ? n0[d] : 1; centering[d] = 0; }
  for (int l = 1; l < nlev; ++l)
    for (int d = 0; d < 3; ++d) {
      int nf = sizes[(l - 1) * 3 + d];
      if (d >= dim) { sizes[l * 3 + d] = 1; centering[l * 3 + d] = 0; continue; }
      if (nf % 2 == 0) { sizes[l * 3 + d] = nf / 2; centering[l * 3 + d] = 1; }
      else { sizes[l * 3 + d] = (nf - 1) / 2 + 1; centering[l * 3 + d] = 0; }
    }
  return nlev;
}

/* 1-D transfer tables: mad/itkInterGridOperators.h:101-127.  Position: 0 left, 1 interior, 2 right.
 * Vertex tables cover offsets -1..1, cell tables offsets -2..2. */
static const double kInterpVertex[3][3] = {{0., 1., .5}, {.5, 1., .5}, {.5, 1., 0.}};
static const double kInterpCell[3][5] = {{0., 0., 1., .75, .25}, {0., .25, .75, .75, .25}, {0., .25, .75, 1., 0.}};
static const double kRestrVertex[3][3] = {{0., 1., 0.}, {.25, .5, .25}, {0., 1., 0.}};
static const double kRestrCell[3][5] = {{0., 0., .5, .375, .125}, {0., .125, .375, .375, .125}, {0., .125, .375, .5, 0.}};

static inline int pos_of(int i, int n) { return i == 0 ? 0 : (n - i == 1 ? 2 : 1); }

/* ---------------------------------------------------------------------------------
 * Full-weighting restriction: mad/itkInterGridOperators.hxx:175-304 (+ GenerateStencil :307-353).
 * Gather; fine index = 2*coarse + offset (:242, :274); per-axis weight from the table chosen by
 * the coarse point's position (left / interior / right, :278-279); out-of-range fine points are
 * skipped (:289).  centering[d]: 0 vertex (radius 1), 1 cell (radius 2).
 * ------------------------------------------------------------------------------- */
void mo_restrict(int dim, const int *nf, const int *centering, const double *fine, const int *nc, double *coarse)
{
  int r[3] = {0, 0, 0};
  for (int d = 0; d < dim; ++d) r[d] = centering[d] ? 2 : 1;
  for (int cz = 0; cz < nc[2]; ++cz)
    for (int cy = 0; cy < nc[1]; ++cy)
      for (int cx = 0; cx < nc[0]; ++cx) {
        int c[3] = {cx, cy, cz};
        int p[3];
        for (int d = 0; d < 3; ++d) p[d] = pos_of(c[d], nc[d]);
        double value = 0.;
        for (int oz = -r[2]; oz <= r[2]; ++oz)
          for (int oy = -r[1]; oy <= r[1]; ++oy)
            for (int ox = -r[0]; ox <= r[0]; ++ox) {
              int o[3] = {ox, oy, oz};
              double w = 1.;
              int inside = 1;
              int fi[3] = {0, 0, 0};
              for (int d = 0; d < dim; ++d) {
                w *= centering[d] ? kRestrCell[p[d]][o[d] + 2] : kRestrVertex[p[d]][o[d] + 1];
                fi[d] = 2 * c[d] + o[d];
                if (fi[d] < 0 || fi[d] >= nf[d]) inside = 0;
              }
              for (int d = dim; d < 3; ++d) fi[d] = 0;
              if (w == 0. || !inside) continue;
              value += w * fine[((int64_t)fi[2] * nf[1] + fi[1]) * nf[0] + fi[0]];
            }
        coarse[((int64_t)cz * nc[1] + cy) * nc[0] + cx] = value;
      }
}

/* ---------------------------------------------------------------------------------
 * Linear interpolation: mad/itkInterGridOperators.hxx:45-172.  Scatter-add of each coarse value
 * into the zero-filled fine image (:78, :118-119, :157-159): fine[2*c + o] += w(o) * coarse[c].
 * ------------------------------------------------------------------------------- */
void mo_interpolate(int dim, const int *nc, const int *centering, const double *coarse, const int *nf, double *fine)
{
  int r[3] = {0, 0, 0};
  for (int d = 0; d < dim; ++d) r[d] = centering[d] ? 2 : 1;
  int64_t nfine = (int64_t)nf[0] * nf[1] * nf[2];
  for (int64_t i = 0; i < nfine; ++i) fine[i] = 0.;
  for (int cz = 0; cz < nc[2]; ++cz)
    for (int cy = 0; cy < nc[1]; ++cy)
      for (int cx = 0; cx < nc[0]; ++cx) {
        int c[3] = {cx, cy, cz};
        int p[3];
        for (int d = 0; d < 3; ++d) p[d] = pos_of(c[d], nc[d]);
        double v = coarse[((int64_t)cz * nc[1] + cy) * nc[0] + cx];
        for (int oz = -r[2]; oz <= r[2]; ++oz)
          for (int oy = -r[1]; oy <= r[1]; ++oy)
            for (int ox = -r[0]; ox <= r[0]; ++ox) {
              int o[3] = {ox, oy, oz};
              double w = 1.;
              int inside = 1;
              int fi[3] = {0, 0, 0};
              for (int d = 0; d < dim; ++d) {
                w *= centering[d] ? kInterpCell[p[d]][o[d] + 2] : kInterpVertex[p[d]][o[d] + 1];
                fi[d] = 2 * c[d] + o[d];
                if (fi[d] < 0 || fi[d] >= nf[d]) inside = 0;
              }
              for (int d = dim; d < 3; ++d) fi[d] = 0;
              if (w == 0. || !inside) continue;
              fine[((int64_t)fi[2] * nf[1] + fi[1]) * nf[0] + fi[0]] += w * v;
            }
      }
}

/* Neighborhood raster index of an offset (x fastest), radius 1. */
static inline int sidx(int dim, const int *o)
{
  return dim == 2 ? (o[1] + 1) * 3 + (o[0] + 1) : ((o[2] + 1) * 3 + (o[1] + 1)) * 3 + (o[0] + 1);
}

/* ---------------------------------------------------------------------------------
 * DCA operator assembly: mad/itkGridsHierarchy.hxx:298-516.
 * Same sequence of "+=" on the same (redirected) offsets as the reference.
 * ------------------------------------------------------------------------------- */
void mo_generate_dca(int dim, const int *n, const double *h, double dt, const double *tensor /*SoA planes*/,
                     double *stencil)
{
  const int ns = dim == 2 ? 9 : 27;
  const int64_t nvox = (int64_t)n[0] * n[1] * n[2];
  const int64_t stride[3] = {1, n[0], (int64_t)n[0] * n[1]};
  for (int z = 0; z < n[2]; ++z)
    for (int y = 0; y < n[1]; ++y)
      for (int x = 0; x < n[0]; ++x) {
        const int index[3] = {x, y, z};
        const int64_t lin = ((int64_t)z * n[1] + y) * n[0] + x;
        double *S = stencil + lin * ns;
        for (int i = 0; i < ns; ++i) S[i] = 0.;                                       /* :344 */
        const int center[3] = {0, 0, 0};
        S[sidx(dim, center)] = 1.;                                                       /* :346 */
#define T(a, b, off) tensor[(int64_t)tcomp(dim, a, b) * nvox + lin + (off)]
        for (int d = 0; d < dim; ++d) {                                                  /* :353 */
          int offP[3] = {0, 0, 0}, offM[3] = {0, 0, 0};
          offP[d] = 1; offM[d] = -1;                                                     /* :356-357 */
          double weight = -dt / (h[d] * h[d]);                                           /* :360 */
          if (index[d] == 0) offM[d] = 1;                                                /* :362 */
          else if (n[d] - index[d] == 1) offP[d] = -1;                                   /* :363 */
          double value = T(d, d, 0) * weight;                                            /* :365 */
          S[sidx(dim, offP)] += value;                                                   /* :367-369 */
          S[sidx(dim, offM)] += value;
          S[sidx(dim, center)] -= 2 * value;
          for (int d2 = 0; d2 < dim; ++d2) {                                             /* :371 */
            weight = -dt / (4 * h[d] * h[d2]);                                           /* :374 */
            int PP[3] = {0, 0, 0}, PM[3] = {0, 0, 0}, MP[3] = {0, 0, 0}, MM[3] = {0, 0, 0};
            PP[d] += 1; PP[d2] += 1;                                                     /* :379-385 */
            PM[d] += 1; PM[d2] -= 1;
            MP[d] -= 1; MP[d2] += 1;
            MM[d] -= 1; MM[d2] -= 1;
            if (index[d] == 0) { MM[d] += 2; MP[d] += 2; }                               /* :388-396 */
            else if (n[d] - index[d] == 1) { PP[d] -= 2; PM[d] -= 2; }                   /* :397-405 */
            if (index[d2] == 0) { MM[d2] += 2; PM[d2] += 2; }                            /* :407-418 */
            else if (n[d2] - index[d2] == 1) { PP[d2] -= 2; MP[d2] -= 2; }               /* :419-430 */
            if (d != d2) {                                                               /* :434-444 */
              value = T(d, d2, 0) * weight;
              S[sidx(dim, PP)] += value;
              S[sidx(dim, PM)] -= value;
              S[sidx(dim, MP)] -= value;
              S[sidx(dim, MM)] += value;
            }
            const int64_t s2 = stride[d2];
            if (index[d2] == 0)                                                          /* :451-457 */
              value = (-3. * T(d, d2, 0) + 4. * T(d, d2, s2) - 1. * T(d, d2, 2 * s2)) * weight;
            else if (n[d2] - index[d2] == 1)                                             /* :458-464 */
              value = (3. * T(d, d2, 0) - 4. * T(d, d2, -s2) + 1. * T(d, d2, -2 * s2)) * weight;
            else                                                                         /* :465-470 */
              value = (T(d, d2, s2) - T(d, d2, -s2)) * weight;
            S[sidx(dim, offP)] += value;                                                 /* :472-473 */
            S[sidx(dim, offM)] -= value;
          }
        }
#undef T
      }
}

/* Active offsets: all of radius 1 in raster order (mad/itkStencilImage.hxx:51-65), minus the 8
 * corners in 3-D (mad/itkGridsHierarchy.hxx:493-513). */
static void build_active(mo_hier *H)
{
  H->nactive = 0;
  int zlo = H->dim == 3 ? -1 : 0, zhi = H->dim == 3 ? 1 : 0;
  for (int oz = zlo; oz <= zhi; ++oz)
    for (int oy = -1; oy <= 1; ++oy)
      for (int ox = -1; ox <= 1; ++ox) {
        if (H->dim == 3 && ox != 0 && oy != 0 && oz != 0) continue;
        int o[3] = {ox, oy, oz};
        int k = H->nactive++;
        H->act_idx[k] = sidx(H->dim, o);
        H->act_off[k][0] = ox; H->act_off[k][1] = oy; H->act_off[k][2] = oz;
      }
}

/* LexOrder(left,right): mad/itkMultigridGaussSeidelSmoother.h:87-100 with right = centre:
 * true iff the offset precedes the centre in (z,y,x) raster order. */
static inline int lex_before_center(int dim, const int *o)
{
  for (int i = dim - 1; i >= 0; --i) {
    if (o[i] < 0) return 1;
    else if (o[i] > 0) return 0;
  }
  return 0;
}

/* ---------------------------------------------------------------------------------
 * Weighted Jacobi: mad/itkMultigridWeightedJacobiSmoother.hxx:33-102.
 * ------------------------------------------------------------------------------- */
void mo_wj_iteration(const mo_hier *H, int l, double omega, const double *in, const double *rhs, double *out)
{
  const mo_level *L = &H->lv[l];
  const int *n = L->n;
  const int ns = H->ns, cidx = ns / 2;
  int64_t lin = 0;
  for (int z = 0; z < n[2]; ++z)
    for (int y = 0; y < n[1]; ++y)
      for (int x = 0; x < n[0]; ++x, ++lin) {
        const double *S = L->stencil + lin * ns;
        double value = rhs[lin];
        for (int k = 0; k < H->nactive; ++k) {
          const int *o = H->act_off[k];
          int xx = x + o[0], yy = y + o[1], zz = z + o[2];
          if (xx < 0 || xx >= n[0] || yy < 0 || yy >= n[1] || zz < 0 || zz >= n[2]) continue;
          if (H->act_idx[k] == cidx) continue;                                           /* :79 */
          value -= S[H->act_idx[k]] * in[((int64_t)zz * n[1] + yy) * n[0] + xx];         /* :82 */
        }
        value *= omega / S[cidx];                                                        /* :88 */
        value += (1 - omega) * in[lin];                                                  /* :89 */
        out[lin] = value;
      }
}

/* ---------------------------------------------------------------------------------
 * Lexicographic Gauss-Seidel: mad/itkMultigridGaussSeidelSmoother.hxx:33-111.
 * Neighbours that precede the centre in raster order are read from the OUTPUT image (:82-86),
 * the others from the input (:88-92); out must not alias in.
 * ------------------------------------------------------------------------------- */
void mo_gs_iteration(const mo_hier *H, int l, const double *in, const double *rhs, double *out)
{
  const mo_level *L = &H->lv[l];
  const int *n = L->n;
  const int ns = H->ns, cidx = ns / 2;
  int64_t lin = 0;
  for (int z = 0; z < n[2]; ++z)
    for (int y = 0; y < n[1]; ++y)
      for (int x = 0; x < n[0]; ++x, ++lin) {
        const double *S = L->stencil + lin * ns;
        double value = rhs[lin];
        for (int k = 0; k < H->nactive; ++k) {
          const int *o = H->act_off[k];
          int xx = x + o[0], yy = y + o[1], zz = z + o[2];
          if (xx < 0 || xx >= n[0] || yy < 0 || yy >= n[1] || zz < 0 || zz >= n[2]) continue;
          if (H->act_idx[k] == cidx) continue;
          const int64_t nb = ((int64_t)zz * n[1] + yy) * n[0] + xx;
          if (lex_before_center(H->dim, o)) value -= S[H->act_idx[k]] * out[nb];
          else value -= S[H->act_idx[k]] * in[nb];
        }
        out[lin] = value / S[cidx];                                                      /* :99 */
      }
}

/* ---------------------------------------------------------------------------------
 * Residual r = f - A u: mad/itkMultigridGaussSeidelSmoother.hxx:114-180
 * (identical in mad/itkMultigridWeightedJacobiSmoother.hxx:105-171).
 * ------------------------------------------------------------------------------- */
void mo_residual(const mo_hier *H, int l, const double *in, const double *rhs, double *res)
{
  const mo_level *L = &H->lv[l];
  const int *n = L->n;
  const int ns = H->ns;
  int64_t lin = 0;
  for (int z = 0; z < n[2]; ++z)
    for (int y = 0; y < n[1]; ++y)
      for (int x = 0; x < n[0]; ++x, ++lin) {
        const double *S = L->stencil + lin * ns;
        double value = rhs[lin];
        for (int k = 0; k < H->nactive; ++k) {
          const int *o = H->act_off[k];
          int xx = x + o[0], yy = y + o[1], zz = z + o[2];
          if (xx < 0 || xx >= n[0] || yy < 0 || yy >= n[1] || zz < 0 || zz >= n[2]) continue;
          value -= S[H->act_idx[k]] * in[((int64_t)zz * n[1] + yy) * n[0] + xx];
        }
        res[lin] = value;
      }
}

/* L2 norm: itkMultigridAnisotropicDiffusionImageFilter.hxx:496-515 (serial accumulation). */
double mo_l2norm(const double *x, int64_t n)
{
  double s = 0;
  for (int64_t i = 0; i < n; ++i) s += x[i] * x[i];
  return sqrt(s);
}

static void smooth_once(mo_hier *H, int l, const double *in, const double *rhs, double *out)
{
  if (H->smoother == 1) mo_wj_iteration(H, l, H->omega, in, rhs, out);
  else mo_gs_iteration(H, l, in, rhs, out);
  H->n_smooth++;
}

/* ---------------------------------------------------------------------------------
 * Coarsest-grid direct solver: mad/itkDirectSolver.hxx:32-88 (assembly with LexPosition,
 * mad/itkDirectSolver.h:89-99; all in-range stencil entries copied) and :91-147 (solve).
 * vnl_sparse_lu (VXL, un-vendored, version unpinned) is an exact sparse LU; restated here as a
 * dense LU with partial pivoting -- any exact solve agrees to rounding.
 * ------------------------------------------------------------------------------- */
static int direct_factor(mo_hier *H)
{
  const mo_level *L = &H->lv[H->nlevels - 1];
  const int *n = L->n;
  const int N = (int)L->nvox;
  H->nc = N;
  H->lu = (double *)calloc((size_t)N * N, sizeof(double));
  H->piv = (int *)malloc(sizeof(int) * N);
  if (!H->lu || !H->piv) return -1;
  int zlo = H->dim == 3 ? -1 : 0, zhi = H->dim == 3 ? 1 : 0;
  int64_t lin = 0;
  for (int z = 0; z < n[2]; ++z)
    for (int y = 0; y < n[1]; ++y)
      for (int x = 0; x < n[0]; ++x, ++lin)
        for (int oz = zlo; oz <= zhi; ++oz)
          for (int oy = -1; oy <= 1; ++oy)
            for (int ox = -1; ox <= 1; ++ox) {
              int xx = x + ox, yy = y + oy, zz = z + oz;
              if (xx < 0 || xx >= n[0] || yy < 0 || yy >= n[1] || zz < 0 || zz >= n[2]) continue;
              int o[3] = {ox, oy, oz};
              int64_t col = ((int64_t)zz * n[1] + yy) * n[0] + xx;
              H->lu[lin * N + col] = L->stencil[lin * H->ns + sidx(H->dim, o)];
            }
  double *A = H->lu;
  for (int k = 0; k < N; ++k) {
    int p = k;
    double best = fabs(A[(int64_t)k * N + k]);
    for (int i = k + 1; i < N; ++i) {
      double v = fabs(A[(int64_t)i * N + k]);
      if (v > best) { best = v; p = i; }
    }
    H->piv[k] = p;
    if (best == 0.) return -2;
    if (p != k)
      for (int j = 0; j < N; ++j) { double t = A[(int64_t)k * N + j]; A[(int64_t)k * N + j] = A[(int64_t)p * N + j]; A[(int64_t)p * N + j] = t; }
    const double inv = 1. / A[(int64_t)k * N + k];
    for (int i = k + 1; i < N; ++i) {
      double m = A[(int64_t)i * N + k];
      if (m == 0.) continue;
      m *= inv;
      A[(int64_t)i * N + k] = m;
      double *ri = A + (int64_t)i * N;
      const double *rk = A + (int64_t)k * N;
      for (int j = k + 1; j < N; ++j) ri[j] -= m * rk[j];
    }
  }
  return 0;
}

void mo_direct_solve(const mo_hier *H, const double *rhs, double *sol)
{
  const int N = H->nc;
  const double *A = H->lu;
  for (int i = 0; i < N; ++i) sol[i] = rhs[i];
  for (int k = 0; k < N; ++k) {
    int p = H->piv[k];
    if (p != k) { double t = sol[k]; sol[k] = sol[p]; sol[p] = t; }
    for (int i = k + 1; i < N; ++i) sol[i] -= A[(int64_t)i * N + k] * sol[k];
  }
  for (int i = N - 1; i >= 0; --i) {
    double s = sol[i];
    for (int j = i + 1; j < N; ++j) s -= A[(int64_t)i * N + j] * sol[j];
    sol[i] = s / A[(int64_t)i * N + i];
  }
}

/* ---------------------------------------------------------------------------------
 * Hierarchy constructor: mad/itkGridsHierarchy.hxx:30-204 (tensor split :112-143, per-level
 * restriction of each component :149-162, DCA per level :110, :188) followed by the direct
 * solver factorisation (itkMultigridAnisotropicDiffusionImageFilter.hxx:131-144).
 * tensor_aos: ITK buffer, dim*(dim+1)/2 doubles per voxel.
 * ------------------------------------------------------------------------------- */
mo_hier *mo_create(int dim, const int *n0, const double *h0, double dt, const double *tensor_aos, int smoother,
                   double omega, int nu, int max_coarse)
{
  mo_hier *H = (mo_hier *)calloc(1, sizeof(mo_hier));
  if (!H) return NULL;
  H->dim = dim; H->ns = dim == 2 ? 9 : 27; H->ncomp = dim == 2 ? 3 : 6; H->dt = dt;
  H->smoother = smoother; H->omega = omega; H->nu = nu;
  int sizes[MO_MAXLEV * 3], cent[MO_MAXLEV * 3];
  H->nlevels = mo_level_schedule(dim, n0, sizes, cent);
  if (H->nlevels < 1) { free(H); return NULL; }
  build_active(H);
  for (int l = 0; l < H->nlevels; ++l) {
    mo_level *L = &H->lv[l];
    L->nvox = 1;
    for (int d = 0; d < 3; ++d) {
      L->n[d] = sizes[l * 3 + d];
      L->centering[d] = cent[l * 3 + d];
      L->h[d] = d < dim ? h0[d] * (double)(1 << l) : 1.;                                /* :80 */
      L->nvox *= L->n[d];
    }
    L->stencil = (double *)malloc(sizeof(double) * L->nvox * H->ns);
    L->tensor = (double *)malloc(sizeof(double) * L->nvox * H->ncomp);
    if (!L->stencil || !L->tensor) return NULL;
    if (l == 0) {
      for (int64_t i = 0; i < L->nvox; ++i)
        for (int c = 0; c < H->ncomp; ++c) L->tensor[(int64_t)c * L->nvox + i] = tensor_aos[i * H->ncomp + c];
    } else {
      const mo_level *F = &H->lv[l - 1];
      for (int c = 0; c < H->ncomp; ++c)
        mo_restrict(dim, F->n, L->centering, F->tensor + (int64_t)c * F->nvox, L->n, L->tensor + (int64_t)c * L->nvox);
    }
    mo_generate_dca(dim, L->n, L->h, dt, L->tensor, L->stencil);
  }
  if (max_coarse > 0 && H->lv[H->nlevels - 1].nvox > max_coarse) { H->nc = 0; return H; }
  if (direct_factor(H) != 0) return NULL;
  return H;
}

void mo_destroy(mo_hier *H)
{
  if (!H) return;
  for (int l = 0; l < H->nlevels; ++l) { free(H->lv[l].stencil); free(H->lv[l].tensor); }
  free(H->lu); free(H->piv); free(H);
}

int mo_nlevels(const mo_hier *H) { return H->nlevels; }
void mo_level_info(const mo_hier *H, int l, int *n, double *h, int *centering)
{
  for (int d = 0; d < 3; ++d) { n[d] = H->lv[l].n[d]; h[d] = H->lv[l].h[d]; centering[d] = H->lv[l].centering[d]; }
}
const double *mo_level_stencil(const mo_hier *H, int l) { return H->lv[l].stencil; }
const double *mo_level_tensor(const mo_hier *H, int l) { return H->lv[l].tensor; }
void mo_set_smoother(mo_hier *H, int smoother, double omega, int nu) { H->smoother = smoother; H->omega = omega; H->nu = nu; }
void mo_smooth(mo_hier *H, int l, const double *in, const double *rhs, double *out) { smooth_once(H, l, in, rhs, out); }

/* ---------------------------------------------------------------------------------
 * V-cycle: itkMultigridAnisotropicDiffusionImageFilter.hxx:341-493.
 * faithful != 0 reproduces the reference's redundant residual + norm after EVERY sweep
 * (:384-411, :437-439, :460-487) so that CPU timings are the reference's; the values returned
 * are identical either way (only the last residual of the descending leg is used, :413).
 * ------------------------------------------------------------------------------- */
static void vcycle(mo_hier *H, int l, const double *u_in, const double *f, double *u_out, int faithful)
{
  const mo_level *L = &H->lv[l];
  const int64_t N = L->nvox;
  double *res = (double *)malloc(sizeof(double) * N);
  double rhsNorm = faithful ? mo_l2norm(f, N) : 1.;                                     /* :352 */
  (void)rhsNorm;
  if (l == H->nlevels - 1) {                                                             /* :356-371 */
    mo_direct_solve(H, f, u_out);
    if (faithful) { mo_residual(H, l, u_out, f, res); H->n_resid++; (void)mo_l2norm(res, N); }
    free(res);
    return;
  }
  double *a = (double *)malloc(sizeof(double) * N);
  double *b = (double *)malloc(sizeof(double) * N);
  memcpy(a, u_in, sizeof(double) * N);                                                   /* :375-379 */
  for (int it = 0; it < H->nu; ++it) {                                                   /* :384-411 */
    smooth_once(H, l, a, f, b);
    { double *t = a; a = b; b = t; }
    if (faithful || it == H->nu - 1) { mo_residual(H, l, a, f, res); H->n_resid++; }
    if (faithful) (void)mo_l2norm(res, N);
  }
  if (H->nu == 0) { mo_residual(H, l, a, f, res); H->n_resid++; }
  const mo_level *C = &H->lv[l + 1];
  double *rc = (double *)malloc(sizeof(double) * C->nvox);
  double *ec0 = (double *)calloc(C->nvox, sizeof(double));                               /* :415-416 */
  double *ec = (double *)malloc(sizeof(double) * C->nvox);
  mo_restrict(H->dim, L->n, C->centering, res, C->n, rc);                                /* :413 */
  vcycle(H, l + 1, ec0, rc, ec, faithful);                                               /* :418-420 */
  mo_interpolate(H->dim, C->n, C->centering, ec, L->n, b);                               /* :422 */
  for (int64_t i = 0; i < N; ++i) a[i] += b[i];                                          /* :424-435 */
  if (faithful) { mo_residual(H, l, a, f, res); H->n_resid++; (void)mo_l2norm(res, N); } /* :437-439 */
  for (int it = 0; it < H->nu; ++it) {                                                   /* :460-487 */
    smooth_once(H, l, a, f, b);
    { double *t = a; a = b; b = t; }
    if (faithful) { mo_residual(H, l, a, f, res); H->n_resid++; (void)mo_l2norm(res, N); }
  }
  memcpy(u_out, a, sizeof(double) * N);
  free(a); free(b); free(res); free(rc); free(ec0); free(ec);
}

void mo_vcycle(mo_hier *H, int l, const double *u_in, const double *f, double *u_out, int faithful)
{
  vcycle(H, l, u_in, f, u_out, faithful);
}

/* Full multigrid: itkMultigridAnisotropicDiffusionImageFilter.hxx:300-338. */
static void fmg(mo_hier *H, int l, const double *f, double *u_out, int faithful)
{
  const mo_level *L = &H->lv[l];
  const int64_t N = L->nvox;
  double *tmp = (double *)malloc(sizeof(double) * N);
  if (l == H->nlevels - 1) {
    memset(u_out, 0, sizeof(double) * N);                                                /* :311-312 */
  } else {
    const mo_level *C = &H->lv[l + 1];
    double *fc = (double *)malloc(sizeof(double) * C->nvox);
    double *uc = (double *)malloc(sizeof(double) * C->nvox);
    mo_restrict(H->dim, L->n, C->centering, f, C->n, fc);                                /* :324 */
    fmg(H, l + 1, fc, uc, faithful);                                                     /* :326 */
    mo_interpolate(H->dim, C->n, C->centering, uc, L->n, u_out);                         /* :330 */
    free(fc); free(uc);
  }
  for (int it = 0; it < H->nu; ++it) {                                                   /* :314, :332 */
    vcycle(H, l, u_out, f, tmp, faithful);
    memcpy(u_out, tmp, sizeof(double) * N);
  }
  free(tmp);
}

void mo_fmg(mo_hier *H, const double *f, double *u_out, int faithful) { fmg(H, 0, f, u_out, faithful); }

/* ---------------------------------------------------------------------------------
 * GenerateData time-step loop: itkMultigridAnisotropicDiffusionImageFilter.hxx:158-263.
 * cycle: 0 VCYCLE, 1 FMG, 2 SMOOTHER (enum CycleType, itkMultigridAnisotropicDiffusionImageFilter.h:123).
 * image: in = rhs of the first step, out = solution of the last step (double; output-pixel
 * cast is done by the caller).  cycles_per_step[n], relres_hist[n*max_cycles + k] are filled if non-NULL.
 * ------------------------------------------------------------------------------- */
int mo_solve(mo_hier *H, int cycle, double tolerance, int max_cycles, int number_of_steps, double *image,
             int *cycles_per_step, double *relres_hist, int faithful)
{
  const int64_t N = H->lv[0].nvox;
  double *rhs = (double *)malloc(sizeof(double) * N);
  double *u = (double *)malloc(sizeof(double) * N);
  double *tmp = (double *)malloc(sizeof(double) * N);
  double *res = (double *)malloc(sizeof(double) * N);
  if (!rhs || !u || !tmp || !res) return -1;
  memcpy(rhs, image, sizeof(double) * N);
  for (int n = 0; n < number_of_steps; ++n) {
    if (cycle == 1) fmg(H, 0, rhs, u, faithful);                                         /* :174 */
    else memcpy(u, rhs, sizeof(double) * N);                                             /* :182-199 */
    double relres;
    const double rhsNorm = mo_l2norm(rhs, N);                                            /* :204 */
    int it = 0;
    do {                                                                                 /* :207-246 */
      if (cycle == 2) smooth_once(H, 0, u, rhs, tmp);                                    /* :213 */
      else vcycle(H, 0, u, rhs, tmp, faithful);                                          /* :235 */
      { double *t = u; u = tmp; tmp = t; }
      mo_residual(H, 0, u, rhs, res); H->n_resid++;                                      /* :215, :237 */
      relres = mo_l2norm(res, N) / rhsNorm;                                              /* :217, :239 */
      if (relres_hist) relres_hist[(int64_t)n * max_cycles + it] = relres;
      if (H->verbose) printf("step %d cycle %d relres %.6e\n", n + 1, it + 1, relres);
      ++it;
    } while (relres > tolerance && it < max_cycles);
    if (cycles_per_step) cycles_per_step[n] = it;
    memcpy(rhs, u, sizeof(double) * N);                                                  /* :248-261 */
  }
  memcpy(image, u, sizeof(double) * N);
  free(rhs); free(u); free(tmp); free(res);
  return 0;
}

void mo_set_verbose(mo_hier *H, int v) { H->verbose = v; }
void mo_counters(const mo_hier *H, int64_t *n_smooth, int64_t *n_resid) { *n_smooth = H->n_smooth; *n_resid = H->n_resid; }
