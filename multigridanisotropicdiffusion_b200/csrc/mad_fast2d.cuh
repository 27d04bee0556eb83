// mad_fast2d.cuh -- streaming kernels for 2-D levels (the reference's itk2DDiffusionTest_{WJ,GS} path), sm_100a.
//
// Same idea as mad_fast.cuh one dimension down: a thread owns four consecutive pixels of a row (128-bit loads / stores),
// a warp 128 columns, and the WARP marches along y over a chunk of rows with the three live rows of u and of the
// y-differentiated tensor components (xy, yy) in registers; x-neighbours come from warp shuffles (+ one scalar load at the
// two warp ends).  A row step loads every field exactly once: 24 B per pixel and sweep (u, f, u', three tensor planes).
// Warps never talk to each other, so there is no block-level synchronisation at all.
//
// The 9-point operator row is evaluated on the fly in the closed form of row_coeffs<2> (mad_kernels.cuh), incl. the Neumann
// node mirror and the one-sided tensor differences of mad/itkGridsHierarchy.hxx:451-470.
//
// Gauss-Seidel (mad/itkMultigridGaussSeidelSmoother.hxx:33-111) in ONE pass: rows in y order (the reference's outer loop),
// inside a row the even columns, then the odd columns -- no two pixels of one of these sets are neighbours in the 9-point
// stencil, and every pixel sees the new values of all pixels before it in that order, so inside a tile of 128 columns x yc rows
// this IS a sequential Gauss-Seidel sweep; values outside the tile are those of the previous sweep (`u`; the result goes to
// `out`).  Same fixed point as the reference's lexicographic sweep; parity is stated on the converged image.
#pragma once
#include "mad_fast.cuh"

namespace mad {
namespace fast {

enum { M2_WJ = 0, M2_RES = 1, M2_GS = 2 };

__device__ __forceinline__ Pos make_pos2(const Geom& g)
{
  Pos p;
  p.lane = threadIdx.x;
  p.xt = blockIdx.x * TX + p.lane * 4;
  p.y = 0;
  p.xl = p.xt < g.nx ? p.xt : 0;
  p.jl = g.nx - 1 - p.xt;
  p.edge = p.lane == 0 || p.lane == 31;
  p.dh = (p.lane == 0 ? max(p.xt - 1, 0) : min(p.xt + 4, g.nx - 1)) - p.xl;
  p.xb = p.xt == 0 || (p.jl >= 0 && p.jl < 4);
  p.ylo = p.yhi = false;
  p.oym = p.oyp = 0;
  return p;
}

template <typename T>
__device__ __forceinline__ V6<T> zero6()
{
  V6<T> w;
#pragma unroll
  for (int i = 0; i < 6; ++i) w.v[i] = T(0);
  return w;
}

// MODE M2_WJ : out = weighted-Jacobi update of u (mad/itkMultigridWeightedJacobiSmoother.hxx:88-89)
// MODE M2_RES: out = f - A u (when out != null) and per-block partial sums of its square (when partials != null)
// MODE M2_GS : out = one Gauss-Seidel sweep of u in the ordering documented above
// T = arithmetic type; UT / FT = storage types of u and f (double for the level-0 outer residual).
// uzero: u is identically zero (first sweep of a V-cycle leg) and is not read.
// grid = (ceil(nx/128), ceil(ceil(ny/yc)/WY)), block = (32, WY).
template <int MODE, typename T, typename UT, typename FT, int WY>
__global__ void __launch_bounds__(32 * WY) k2_sweep(Geom g, Tensor D, const UT* __restrict__ u, const FT* __restrict__ f, float* __restrict__ out,
                                                    double* __restrict__ partials, float omega, int yc, int uzero)
{
  const Pos p = make_pos2(g);
  const int y0 = ((int)blockIdx.y * WY + (int)threadIdx.y) * yc, y1 = min(y0 + yc, g.ny);
  const GeomConst<T> k(g);
  const float* __restrict__ Dxx = D.p[XX2];
  const float* __restrict__ Dxy = D.p[XY2];
  const float* __restrict__ Dyy = D.p[YY2];
  double sq = 0.0;
  if (y0 < g.ny) {
    auto off_of = [&](int y) { return y * g.pitch + p.xl; };
    auto row_m = [&](int y) { return y == 0 ? 1 : y - 1; };                  // node mirror (ny >= 2)
    auto row_p = [&](int y) { return y == g.ny - 1 ? g.ny - 2 : y + 1; };
    auto clamp_y = [&](int y) { return min(max(y, 0), g.ny - 1); };
    auto load_u = [&](int y) {
      V6<T> w = finish6<T, UT>(issue6(u, off_of(y), p), p);
      if (p.xb) mirror_x(w, p.xt, p.jl);
      return w;
    };
    V6<T> um = zero6<T>(), uc = zero6<T>(), up = zero6<T>();
    if (!uzero) { um = load_u(row_m(y0)); uc = load_u(y0); }
    V4<float> xym = finish4<float>(issue4(Dxy, off_of(clamp_y(y0 - 1))));
    V4<float> yym = finish4<float>(issue4(Dyy, off_of(clamp_y(y0 - 1))));
    V6<float> xyc = finish6<float, float>(issue6(Dxy, off_of(y0), p), p);
    V4<float> yyc = finish4<float>(issue4(Dyy, off_of(y0)));
    for (int y = y0; y < y1; ++y) {
      const int oc = off_of(y), on = off_of(clamp_y(y + 1));
      // every load of the step is issued before anything is consumed
      Raw6<UT> ru;
      if (!uzero) ru = issue6(u, off_of(row_p(y)), p);
      const Raw6<float> rxy = issue6(Dxy, on, p), rxx = issue6(Dxx, oc, p);
      const Raw4<float> ryy = issue4(Dyy, on);
      const Raw4<FT> rf = issue4(f, oc);
      if (!uzero) {
        up = finish6<T, UT>(ru, p);
        if (p.xb) mirror_x(up, p.xt, p.jl);
      }
      // last row: its mirrored y+1 neighbour IS row ny-2, which this sweep has already relaxed (when it lies in the tile)
      if (MODE == M2_GS && y == g.ny - 1) up = um;
      const V6<float> xyn = finish6<float, float>(rxy, p), xxc = finish6<float, float>(rxx, p);
      const V4<float> yyn = finish4<float>(ryy);
      const V4<T> fv = finish4<T>(rf);
      // tensor differences: central (un-normalised D+ - D-), one-sided on the first / last node
      T dx_xx[4], dx_xy[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { dx_xx[j] = T(xxc.v[j + 2]) - T(xxc.v[j]); dx_xy[j] = T(xyc.v[j + 2]) - T(xyc.v[j]); }
      if (p.xb) { xdiff_boundary<T>(xxc, p, Dxx, oc, dx_xx); xdiff_boundary<T>(xyc, p, Dxy, oc, dx_xy); }
      V4<T> dy_xy, dy_yy;
      if (y == 0 || y == g.ny - 1) {  // warp-uniform
        const int y2 = y == 0 ? min(2, g.ny - 1) : max(g.ny - 3, 0);
        const V4<float> xy2 = finish4<float>(issue4(Dxy, off_of(y2))), yy2 = finish4<float>(issue4(Dyy, off_of(y2)));
        const bool lo = y == 0;
        dy_xy = onesided4<T>(mid4(xyc), lo ? mid4(xyn) : xym, xy2, lo ? T(-1) : T(1));
        dy_yy = onesided4<T>(yyc, lo ? yyn : yym, yy2, lo ? T(-1) : T(1));
      } else {
        dy_xy = sub4<T>(mid4(xyn), xym);
        dy_yy = sub4<T>(yyn, yym);
      }
      T diag[4], cxp[4], cxm[4], cyp[4], cym[4], cxy[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const T ax = k.wx * T(xxc.v[j + 1]), ay = k.wy * T(yyc.v[j]);
        diag[j] = T(1) + T(2) * (ax + ay);
        const T bx = -(k.bxx * dx_xx[j] + k.bxy * dy_xy.v[j]);
        const T by = -(k.bxy * dx_xy[j] + k.byy * dy_yy.v[j]);
        cxp[j] = -ax + bx; cxm[j] = -ax - bx; cyp[j] = -ay + by; cym[j] = -ay - by;
        cxy[j] = -k.cxy * T(xyc.v[j + 1]);
      }
      auto offd = [&](int j) {
        T s = cxp[j] * uc.v[j + 2] + cxm[j] * uc.v[j] + cyp[j] * up.v[j + 1] + cym[j] * um.v[j + 1];
        s += cxy[j] * ((up.v[j + 2] - um.v[j + 2]) - (up.v[j] - um.v[j]));
        return s;
      };
      float res[4];
      if (MODE == M2_GS) {
        // even columns (slots 0, 2), then odd columns (1, 3) with the new even values -- the right neighbour's slot 0
        // arrives by shuffle; lanes 0 / 31 keep the previous sweep's value of the pixel outside the tile
        const T n0 = fast_div(fv.v[0] - offd(0), diag[0]), n2 = fast_div(fv.v[2] - offd(2), diag[2]);
        uc.v[1] = n0; uc.v[3] = n2;
        {
          const T r = __shfl_down_sync(FULL, n0, 1);
          if (p.lane < 31) uc.v[5] = r;
          if (p.xb) mirror_x(uc, p.xt, p.jl);
        }
        const T n1 = fast_div(fv.v[1] - offd(1), diag[1]), n3 = fast_div(fv.v[3] - offd(3), diag[3]);
        uc.v[2] = n1; uc.v[4] = n3;
        {
          const T l = __shfl_up_sync(FULL, n3, 1);
          if (p.lane > 0) uc.v[0] = l;
          if (p.xb) mirror_x(uc, p.xt, p.jl);
        }
        res[0] = (float)n0; res[1] = (float)n1; res[2] = (float)n2; res[3] = (float)n3;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const T off = offd(j);
          if (MODE == M2_WJ) {
            res[j] = (float)((fv.v[j] - off) * fast_div(T(omega), diag[j]) + (T(1) - T(omega)) * uc.v[j + 1]);
          } else {
            const T v = fv.v[j] - diag[j] * uc.v[j + 1] - off;
            res[j] = (float)v;
            if (p.xt + j < g.nx) sq += (double)v * (double)v;
          }
        }
      }
      if (out && p.xt < g.nx) store4<float>(out, y * g.pitch + p.xt, p.xt, g.nx, res);
      um = uc; uc = up;
      xym = mid4(xyc); xyc = xyn;
      yym = yyc; yyc = yyn;
    }
  }
  if (MODE == M2_RES && partials) {
    const double t = block_sum(sq);
    if (threadIdx.x == 0 && threadIdx.y == 0) partials[(size_t)blockIdx.x + (size_t)gridDim.x * blockIdx.y] = t;
  }
}

// coarse = R fine, 2-D full weighting (mad/itkInterGridOperators.hxx:175-304, tables .h:115-127): a thread reads the fine
// pixels 4t-1..4t+4 of the (up to) four contributing fine rows and produces the coarse pixels 2t, 2t+1 (one 8-byte store).
// grid = (ceil(nxf/128), ceil(nyc/WY)), block = (32, WY).
template <int WY>
__global__ void __launch_bounds__(32 * WY) k2_restrict(Geom gf, Geom gc, Transfer t, const float* __restrict__ fine, float* __restrict__ coarse)
{
  const Pos p = make_pos2(gf);
  const int yc = blockIdx.y * WY + threadIdx.y;
  if (yc >= gc.ny) return;  // whole warp
  float wy[4], wxa[4], wxb[4];
  restrict_taps(yc, gc.ny, t.cent[1], wy);
  const int xc0 = p.xt >> 1;
  restrict_taps(min(xc0, gc.nx - 1), gc.nx, t.cent[0], wxa);
  restrict_taps(min(xc0 + 1, gc.nx - 1), gc.nx, t.cent[0], wxb);
  Raw6<float> raw[4];
#pragma unroll
  for (int ky = 0; ky < 4; ++ky) raw[ky] = issue6(fine, min(max(2 * yc + ky - 1, 0), gf.ny - 1) * gf.pitch + p.xl, p);  // clamped rows carry weight 0
  float a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int ky = 0; ky < 4; ++ky) {
    const V6<float> v = finish6<float, float>(raw[ky], p);
    float m[6];  // pixels beyond the fine row only ever meet zero weights, but may hold anything: mask them
#pragma unroll
    for (int i = 0; i < 6; ++i) { const int fx = p.xt - 1 + i; m[i] = (fx >= 0 && fx < gf.nx) ? v.v[i] : 0.f; }
    a0 += wy[ky] * (wxa[0] * m[0] + wxa[1] * m[1] + wxa[2] * m[2] + wxa[3] * m[3]);
    a1 += wy[ky] * (wxb[0] * m[2] + wxb[1] * m[3] + wxb[2] * m[4] + wxb[3] * m[5]);
  }
  const int o = yc * gc.pitch + xc0;
  if (xc0 + 1 < gc.nx) *reinterpret_cast<float2*>(coarse + o) = make_float2(a0, a1);
  else if (xc0 < gc.nx) coarse[o] = a0;
}

// fine (+)= P coarse, 2-D bilinear interpolation in gather form (mad/itkInterGridOperators.hxx:45-172, tables .h:101-113) fused
// with the correction add (…Filter.hxx:424-435): the four fine pixels 4t..4t+3 of a thread interpolate from the coarse pixels
// 2t-1..2t+2 of two coarse rows (one 8-byte pair + two scalars per row), one 16-byte load / store of the fine row.
// grid = (ceil(nxf/128), ceil(nyf/WY)), block = (32, WY).
template <bool ADD, int WY>
__global__ void __launch_bounds__(32 * WY) k2_prolong(Geom gc, Geom gf, Transfer t, const float* __restrict__ coarse, float* __restrict__ fine)
{
  const int xt = blockIdx.x * TX + threadIdx.x * 4;
  const int y = blockIdx.y * WY + threadIdx.y;
  if (xt >= gf.nx || y >= gf.ny) return;
  int y0, y1;
  float wy0, wy1;
  prolong_taps(y, gf.ny, gc.ny, t.cent[1], y0, y1, wy0, wy1);
  // x taps of the four fine pixels as weights on the coarse pixels cb..cb+3, cb = 2t-1
  const int cb = (xt >> 1) - 1;
  float W[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int i0, i1;
    float w0, w1;
    prolong_taps(xt + j, gf.nx, gc.nx, t.cent[0], i0, i1, w0, w1);
    if (xt + j >= gf.nx) { w0 = 0.f; w1 = 0.f; i0 = i1 = cb + 1; }
#pragma unroll
    for (int q = 0; q < 4; ++q) W[j][q] = (i0 - cb == q ? w0 : 0.f) + (i1 - cb == q ? w1 : 0.f);
  }
  const int xa = max(cb, 0), xd = min(cb + 3, gc.nx - 1);  // clamped (their weights are zero when clamped)
  const int xb = cb + 1, xc = min(cb + 2, gc.nx - 1);     // cb+1 = 2t is always a valid, 8-byte aligned pixel
  const bool pair = xb + 1 < gc.nx;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const float wr = r ? wy1 : wy0;
    const float* row = coarse + (r ? y1 : y0) * gc.pitch;
    float c[4];
    c[0] = __ldg(row + xa);
    if (pair) { const float2 q = __ldg(reinterpret_cast<const float2*>(row + xb)); c[1] = q.x; c[2] = q.y; }
    else { c[1] = __ldg(row + xb); c[2] = __ldg(row + xc); }
    c[3] = __ldg(row + xd);
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] += wr * (W[j][0] * c[0] + W[j][1] * c[1] + W[j][2] * c[2] + W[j][3] * c[3]);
  }
  const int o = y * gf.pitch + xt;
  if (ADD) {
    const float4 q = *reinterpret_cast<const float4*>(fine + o);
    acc[0] += q.x; acc[1] += q.y; acc[2] += q.z; acc[3] += q.w;
  }
  store4<float>(fine, o, xt, gf.nx, acc);
}

}  // namespace fast
}  // namespace mad
