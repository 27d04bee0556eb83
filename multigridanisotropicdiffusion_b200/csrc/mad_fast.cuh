// mad_fast.cuh -- the streaming 3-D kernels of libmadgpu (sm_100a): one thread = 4 consecutive x
// voxels (128-bit loads/stores), one warp = 128 voxels of one image row, one CTA = WY consecutive
// rows, marching along z over a chunk of planes with the three live planes of u and of the
// z-differentiated tensor components (xz, yz, zz) held in REGISTERS.  x-neighbours come from warp
// shuffles (plus one scalar load at the two warp ends), y-neighbour rows are re-loaded through L1
// (the neighbouring warp of the same CTA loads the same lines at the same time), so every field is
// fetched from HBM once per sweep: 36 B/voxel (u, f, u', six tensor planes).
//
// The operator row is evaluated on the fly in the closed form of row_coeffs()/apply_offdiag()
// (mad_kernels.cuh; reference: mad/itkGridsHierarchy.hxx:298-516) with node-mirrored reads at the
// Neumann boundary and the one-sided tensor differences of mad/itkGridsHierarchy.hxx:451-470.
//
//   k_fast_sweep<MODE_WJ>   mad/itkMultigridWeightedJacobiSmoother.hxx:33-102
//   k_fast_sweep<MODE_RES>  mad/itkMultigridGaussSeidelSmoother.hxx:114-180 (+ L2Norm partial sums,
//                           itkMultigridAnisotropicDiffusionImageFilter.hxx:496-515)
//   k_fast_gs               mad/itkMultigridGaussSeidelSmoother.hxx:33-111 in the ordering
//                           "planes in z order; inside a plane even rows (even x, then odd x), then
//                           odd rows", exact inside a CTA tile, previous-sweep values outside it.
#pragma once
#include "mad_kernels.cuh"

namespace mad {
namespace fast {

constexpr unsigned FULL = 0xffffffffu;
constexpr int TX = 128;  // voxels per warp row

template <typename T>
struct V4 {
  T v[4];
};
template <typename T>
struct V6 {
  T v[6];  // v[0] = x-1, v[1..4] = the thread's four voxels, v[5] = x+4
};

// Per-thread position inside the volume.
struct Pos {
  int xt, lane, y, jl, hx;
  bool act, edge, ylo, yhi;
  long long oym, oyp;  // element offsets of the mirrored y-1 / y+1 rows relative to row y
};

__device__ __forceinline__ Pos make_pos(const Geom& g)
{
  Pos p;
  p.lane = threadIdx.x;
  p.xt = blockIdx.x * TX + p.lane * 4;
  p.y = blockIdx.y * blockDim.y + threadIdx.y;
  p.act = p.xt < g.nx;
  p.jl = g.nx - 1 - p.xt;
  p.edge = p.lane == 0 || p.lane == 31;
  p.hx = p.lane == 0 ? max(p.xt - 1, 0) : min(p.xt + 4, g.nx - 1);  // x of the halo voxel the warp-end lanes fetch
  p.ylo = p.y == 0;
  p.yhi = p.y == g.ny - 1;
  p.oym = p.ylo ? g.pitch : -(long long)g.pitch;
  p.oyp = p.yhi ? -(long long)g.pitch : g.pitch;
  return p;
}

// Loads are split in two phases so that a plane step first ISSUES every load it needs (nothing in
// between depends on a loaded value, the warp keeps ~11 KB in flight) and only then consumes them:
// issue4/issue6 return the raw registers, finish4/finish6 convert and exchange the x-neighbours.
template <typename ST>
struct Raw4;
template <>
struct Raw4<float> {
  float4 q;
};
template <>
struct Raw4<double> {
  double2 a, b;
};
template <typename ST>
struct Raw6 {
  Raw4<ST> c;
  ST h;  // halo voxel, meaningful in lanes 0 (x-1) and 31 (x+4)
};

__device__ __forceinline__ Raw4<float> issue4(const float* __restrict__ row, const Pos& p)
{
  Raw4<float> r;
  r.q = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.act) r.q = __ldg(reinterpret_cast<const float4*>(row + p.xt));
  return r;
}
__device__ __forceinline__ Raw4<double> issue4(const double* __restrict__ row, const Pos& p)
{
  Raw4<double> r;
  r.a = r.b = make_double2(0.0, 0.0);
  if (p.act) {
    r.a = __ldg(reinterpret_cast<const double2*>(row + p.xt));
    r.b = __ldg(reinterpret_cast<const double2*>(row + p.xt + 2));
  }
  return r;
}
template <typename ST>
__device__ __forceinline__ Raw6<ST> issue6(const ST* __restrict__ row, const Pos& p)
{
  Raw6<ST> r;
  r.c = issue4(row, p);
  r.h = ST(0);
  if (p.edge) r.h = __ldg(row + p.hx);
  return r;
}

template <typename T>
__device__ __forceinline__ V4<T> finish4(const Raw4<float>& r)
{
  V4<T> o;
  o.v[0] = T(r.q.x); o.v[1] = T(r.q.y); o.v[2] = T(r.q.z); o.v[3] = T(r.q.w);
  return o;
}
template <typename T>
__device__ __forceinline__ V4<T> finish4(const Raw4<double>& r)
{
  V4<T> o;
  o.v[0] = T(r.a.x); o.v[1] = T(r.a.y); o.v[2] = T(r.b.x); o.v[3] = T(r.b.y);
  return o;
}
// Four voxels plus the two x-neighbours: shuffles inside the warp, the pre-fetched halo voxel at the warp ends.
template <typename T, typename ST>
__device__ __forceinline__ V6<T> finish6(const Raw6<ST>& r, const Pos& p)
{
  const V4<T> c = finish4<T>(r.c);
  T l = __shfl_up_sync(FULL, c.v[3], 1);
  T rr = __shfl_down_sync(FULL, c.v[0], 1);
  if (p.lane == 0) l = T(r.h);
  if (p.lane == 31) rr = T(r.h);
  V6<T> w;
  w.v[0] = l; w.v[1] = c.v[0]; w.v[2] = c.v[1]; w.v[3] = c.v[2]; w.v[4] = c.v[3]; w.v[5] = rr;
  return w;
}

// one-shot forms (start-up planes, boundary rows)
template <typename T, typename ST>
__device__ __forceinline__ V4<T> load4(const ST* __restrict__ row, const Pos& p)
{
  return finish4<T>(issue4(row, p));
}
template <typename T, typename ST>
__device__ __forceinline__ V6<T> load6(const ST* __restrict__ row, const Pos& p)
{
  return finish6<T, ST>(issue6(row, p), p);
}

// Node mirror along x for fields the stencil is applied to: u(-1) = u(1), u(nx) = u(nx-2).
// jl = nx-1-xt is the slot of the last voxel of the row when it lies in this thread.
template <typename T>
__device__ __forceinline__ void mirror_x(V6<T>& w, int xt, int jl)
{
  if (xt == 0) w.v[0] = w.v[2];
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (jl == j) w.v[j + 2] = w.v[j];
}

// x-difference of a tensor component at slot j (mad/itkGridsHierarchy.hxx:451-470)
template <typename T>
__device__ __forceinline__ T xdiff(const V6<float>& w, int j, int xt, int jl, const float* __restrict__ row)
{
  T d = T(w.v[j + 2]) - T(w.v[j]);
  if (xt == 0 && j == 0) d = T(-3) * T(w.v[1]) + T(4) * T(w.v[2]) - T(w.v[3]);
  if (jl == j) {
    const T m2 = j >= 1 ? T(w.v[j - 1]) : T(__ldg(row + xt - 2));
    d = T(3) * T(w.v[j + 1]) - T(4) * T(w.v[j]) + m2;
  }
  return d;
}

// The tensor is kept in registers as stored (fp32); differences are formed in the arithmetic type T.
template <typename T>
__device__ __forceinline__ V4<T> sub4(const V4<float>& a, const V4<float>& b)
{
  V4<T> r;
#pragma unroll
  for (int j = 0; j < 4; ++j) r.v[j] = T(a.v[j]) - T(b.v[j]);
  return r;
}

// one-sided second-order difference: sgn * (3 c - 4 n1 + n2)
template <typename T>
__device__ __forceinline__ V4<T> onesided4(const V4<float>& c, const V4<float>& n1, const V4<float>& n2, T sgn)
{
  V4<T> r;
#pragma unroll
  for (int j = 0; j < 4; ++j) r.v[j] = sgn * (T(3) * T(c.v[j]) - T(4) * T(n1.v[j]) + T(n2.v[j]));
  return r;
}

// element-wise select (a `c ? a : b` on the structs would make ptxas index them through local memory)
__device__ __forceinline__ V4<float> sel4(bool c, const V4<float>& a, const V4<float>& b)
{
  V4<float> r;
#pragma unroll
  for (int j = 0; j < 4; ++j) r.v[j] = c ? a.v[j] : b.v[j];
  return r;
}

template <typename T>
__device__ __forceinline__ V4<T> mid4(const V6<T>& w)
{
  V4<T> r;
  r.v[0] = w.v[1]; r.v[1] = w.v[2]; r.v[2] = w.v[3]; r.v[3] = w.v[4];
  return r;
}

template <typename OT>
__device__ __forceinline__ void store4(OT* __restrict__ row, int xt, int nx, const float v[4])
{
  if (xt + 3 < nx) {
    if constexpr (sizeof(OT) == 4) *reinterpret_cast<float4*>(row + xt) = make_float4(v[0], v[1], v[2], v[3]);
    else {
      *reinterpret_cast<double2*>(row + xt) = make_double2(v[0], v[1]);
      *reinterpret_cast<double2*>(row + xt + 2) = make_double2(v[2], v[3]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (xt + j < nx) row[xt + j] = OT(v[j]);
  }
}

// Tensor rows of one plane step: everything the row of A needs besides u.
template <typename T>
struct Coef {
  T diag[4], xp[4], xm[4], yp[4], ym[4], zp[4], zm[4], exy[4], exz[4], eyz[4];
};

// z-marched tensor state (components differentiated along z), kept as stored (fp32)
struct DzState {
  V4<float> xz_m, yz_m, zz_m;  // plane z-1
  V6<float> xz_c;              // plane z (x-halo for the x-difference)
  V4<float> yz_c, zz_c;
  V6<float> xz_p;              // plane z+1
  V4<float> yz_p, zz_p;
};

struct DzRaw {
  Raw6<float> xz;
  Raw4<float> yz, zz;
};
__device__ __forceinline__ DzRaw issue_dz(const Tensor& D, long long rowoff, const Pos& p)
{
  DzRaw r;
  r.xz = issue6(D.p[XZ3] + rowoff, p);
  r.yz = issue4(D.p[YZ3] + rowoff, p);
  r.zz = issue4(D.p[ZZ3] + rowoff, p);
  return r;
}

// tensor rows that are only needed on the plane being updated
struct DcRaw {
  Raw6<float> xx, xy;
  Raw4<float> xy_m, xy_p, yy, yy_m, yy_p, yz_ym, yz_yp;
};
__device__ __forceinline__ DcRaw issue_dc(const Tensor& D, long long rowoff, const Pos& p)
{
  DcRaw r;
  r.xx = issue6(D.p[XX3] + rowoff, p);
  r.xy = issue6(D.p[XY3] + rowoff, p);
  r.xy_m = issue4(D.p[XY3] + rowoff + p.oym, p);
  r.xy_p = issue4(D.p[XY3] + rowoff + p.oyp, p);
  r.yy = issue4(D.p[YY3] + rowoff, p);
  r.yy_m = issue4(D.p[YY3] + rowoff + p.oym, p);
  r.yy_p = issue4(D.p[YY3] + rowoff + p.oyp, p);
  r.yz_ym = issue4(D.p[YZ3] + rowoff + p.oym, p);
  r.yz_yp = issue4(D.p[YZ3] + rowoff + p.oyp, p);
  return r;
}

// Coefficients of the four rows of A at plane z (rowoff = element offset of row (y, z)).
template <typename T>
__device__ __forceinline__ void coefficients(const Geom& g, const Tensor& D, const Pos& p, int z, long long rowoff, const DzState& S,
                                             const DcRaw& R, Coef<T>& c)
{
  const GeomConst<T> k(g);
  typedef V4<float> F4;
  typedef V6<float> F6;
  const F6 xx = finish6<float, float>(R.xx, p), xy = finish6<float, float>(R.xy, p);
  const F4 xy_m = finish4<float>(R.xy_m), xy_p = finish4<float>(R.xy_p);
  const F4 yy = finish4<float>(R.yy), yy_m = finish4<float>(R.yy_m), yy_p = finish4<float>(R.yy_p);
  const F4 yz_ym = finish4<float>(R.yz_ym), yz_yp = finish4<float>(R.yz_yp);

  // y-differences (central; one-sided on the first / last row -- warp-uniform branches)
  V4<T> dy_xy = sub4<T>(xy_p, xy_m), dy_yy = sub4<T>(yy_p, yy_m), dy_yz = sub4<T>(yz_yp, yz_ym);
  if (p.ylo || p.yhi) {
    const long long o2 = p.ylo ? 2ll * g.pitch : -2ll * g.pitch;
    const T sgn = p.ylo ? T(-1) : T(1);
    dy_xy = onesided4<T>(mid4(xy), sel4(p.ylo, xy_p, xy_m), load4<float, float>(D.p[XY3] + rowoff + o2, p), sgn);
    dy_yy = onesided4<T>(yy, sel4(p.ylo, yy_p, yy_m), load4<float, float>(D.p[YY3] + rowoff + o2, p), sgn);
    dy_yz = onesided4<T>(S.yz_c, sel4(p.ylo, yz_yp, yz_ym), load4<float, float>(D.p[YZ3] + rowoff + o2, p), sgn);
  }
  // z-differences
  V4<T> dz_xz = sub4<T>(mid4(S.xz_p), S.xz_m), dz_yz = sub4<T>(S.yz_p, S.yz_m), dz_zz = sub4<T>(S.zz_p, S.zz_m);
  const bool zlo = z == 0 && g.zlo_phys, zhi = z == g.nz - 1 && g.zhi_phys;
  if (zlo || zhi) {
    const long long o2 = zlo ? 2 * g.plane : -2 * g.plane;
    const T sgn = zlo ? T(-1) : T(1);
    dz_xz = onesided4<T>(mid4(S.xz_c), sel4(zlo, mid4(S.xz_p), S.xz_m), load4<float, float>(D.p[XZ3] + rowoff + o2, p), sgn);
    dz_yz = onesided4<T>(S.yz_c, sel4(zlo, S.yz_p, S.yz_m), load4<float, float>(D.p[YZ3] + rowoff + o2, p), sgn);
    dz_zz = onesided4<T>(S.zz_c, sel4(zlo, S.zz_p, S.zz_m), load4<float, float>(D.p[ZZ3] + rowoff + o2, p), sgn);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const T ax = k.wx * T(xx.v[j + 1]), ay = k.wy * T(yy.v[j]), az = k.wz * T(S.zz_c.v[j]);
    c.diag[j] = T(1) + T(2) * (ax + ay + az);
    const T dx_xx = xdiff<T>(xx, j, p.xt, p.jl, D.p[XX3] + rowoff);
    const T dx_xy = xdiff<T>(xy, j, p.xt, p.jl, D.p[XY3] + rowoff);
    const T dx_xz = xdiff<T>(S.xz_c, j, p.xt, p.jl, D.p[XZ3] + rowoff);
    const T bx = -(k.bxx * dx_xx + k.bxy * dy_xy.v[j] + k.bxz * dz_xz.v[j]);
    const T by = -(k.bxy * dx_xy + k.byy * dy_yy.v[j] + k.byz * dz_yz.v[j]);
    const T bz = -(k.bxz * dx_xz + k.byz * dy_yz.v[j] + k.bzz * dz_zz.v[j]);
    c.xp[j] = -ax + bx; c.xm[j] = -ax - bx;
    c.yp[j] = -ay + by; c.ym[j] = -ay - by;
    c.zp[j] = -az + bz; c.zm[j] = -az - bz;
    c.exy[j] = -k.cxy * T(xy.v[j + 1]);
    c.exz[j] = -k.cxz * T(S.xz_c.v[j + 1]);
    c.eyz[j] = -k.cyz * T(S.yz_c.v[j]);
  }
}

// three rows (y-1, y, y+1, mirrored) of one plane of u
template <typename T>
struct UPlane {
  V6<T> r[3];
};
template <typename UT>
struct URaw {
  Raw6<UT> r[3];
};
template <typename UT>
__device__ __forceinline__ URaw<UT> issue_u(const UT* __restrict__ u, long long rowoff, const Pos& p)
{
  URaw<UT> R;
  R.r[0] = issue6(u + rowoff + p.oym, p);
  R.r[1] = issue6(u + rowoff, p);
  R.r[2] = issue6(u + rowoff + p.oyp, p);
  return R;
}
template <typename T, typename UT>
__device__ __forceinline__ void finish_u(const URaw<UT>& R, const Pos& p, UPlane<T>& P)
{
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    P.r[i] = finish6<T, UT>(R.r[i], p);
    mirror_x(P.r[i], p.xt, p.jl);
  }
}

// sum over the off-diagonal entries at slot j (apply_offdiag of mad_kernels.cuh on registers)
template <typename T>
__device__ __forceinline__ T offdiag(const Coef<T>& c, const UPlane<T>& m, const UPlane<T>& q, const UPlane<T>& n, int j)
{
  T s = c.xp[j] * q.r[1].v[j + 2] + c.xm[j] * q.r[1].v[j] + c.yp[j] * q.r[2].v[j + 1] + c.ym[j] * q.r[0].v[j + 1];
  s += c.exy[j] * ((q.r[2].v[j + 2] - q.r[0].v[j + 2]) - (q.r[2].v[j] - q.r[0].v[j]));
  s += c.zp[j] * n.r[1].v[j + 1] + c.zm[j] * m.r[1].v[j + 1];
  s += c.exz[j] * ((n.r[1].v[j + 2] - m.r[1].v[j + 2]) - (n.r[1].v[j] - m.r[1].v[j]));
  s += c.eyz[j] * ((n.r[2].v[j + 1] - m.r[2].v[j + 1]) - (n.r[0].v[j + 1] - m.r[0].v[j + 1]));
  return s;
}

__device__ __forceinline__ long long zmirror_lo(const Geom& g, int z) { return (z == 0 && g.zlo_phys) ? 1 : z - 1; }
__device__ __forceinline__ long long zmirror_hi(const Geom& g, int z) { return (z == g.nz - 1 && g.zhi_phys) ? g.nz - 2 : z + 1; }

enum { MODE_WJ = 0, MODE_RES = 1 };

// One pass over the volume.  MODE_WJ: out = weighted-Jacobi update of u.  MODE_RES: out = f - A u
// (out may be null) and per-CTA partial sums of its squares (partials may be null).
// grid = (ceil(nx/128), ceil(ny/WY), ceil(nz/zc)), block = (32, WY); MINB = CTAs per SM the register
// allocation is capped for.
template <int MODE, typename T, typename UT, typename FT, typename OT, int WY, int MINB>
__global__ void __launch_bounds__(32 * WY, MINB) k_fast_sweep(Geom g, Tensor D, const UT* __restrict__ u, const FT* __restrict__ f,
                                                         OT* __restrict__ out, double* __restrict__ partials, float omega, int zc)
{
  const Pos p = make_pos(g);
  const bool valid = p.y < g.ny;
  double sq = 0.0;
  if (valid) {
    const int z0 = blockIdx.z * zc, z1 = min(z0 + zc, g.nz);
    const long long rowy = (long long)p.y * g.pitch;
    UPlane<T> um, uc, up;
    DzState S;
    {
      const long long om = zmirror_lo(g, z0) * g.plane + rowy, oc = (long long)z0 * g.plane + rowy;
      const URaw<UT> r0 = issue_u(u, om, p), r1 = issue_u(u, oc, p);
      const DzRaw d0 = issue_dz(D, om, p), d1 = issue_dz(D, oc, p);
      finish_u<T, UT>(r0, p, um);
      finish_u<T, UT>(r1, p, uc);
      S.xz_m = finish4<float>(d0.xz.c); S.yz_m = finish4<float>(d0.yz); S.zz_m = finish4<float>(d0.zz);
      S.xz_c = finish6<float, float>(d1.xz, p); S.yz_c = finish4<float>(d1.yz); S.zz_c = finish4<float>(d1.zz);
    }
    for (int z = z0; z < z1; ++z) {
      const long long oc = (long long)z * g.plane + rowy, on = zmirror_hi(g, z) * g.plane + rowy;
      // ---- issue every load of this plane step ----
      const URaw<UT> ru = issue_u(u, on, p);
      const DzRaw rd = issue_dz(D, on, p);
      const DcRaw rc = issue_dc(D, oc, p);
      const Raw4<FT> rf = issue4(f + oc, p);
      // ---- consume ----
      finish_u<T, UT>(ru, p, up);
      S.xz_p = finish6<float, float>(rd.xz, p); S.yz_p = finish4<float>(rd.yz); S.zz_p = finish4<float>(rd.zz);
      const V4<T> fv = finish4<T>(rf);
      Coef<T> c;
      coefficients<T>(g, D, p, z, oc, S, rc, c);
      float res[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const T s = offdiag<T>(c, um, uc, up, j);
        const T uj = uc.r[1].v[j + 1];
        if (MODE == MODE_WJ) {
          // mad/itkMultigridWeightedJacobiSmoother.hxx:88-89
          res[j] = float((fv.v[j] - s) * (T(omega) / c.diag[j]) + (T(1) - T(omega)) * uj);
        } else {
          const T r = fv.v[j] - c.diag[j] * uj - s;
          res[j] = float(r);
          if (p.xt + j < g.nx) sq += (double)r * (double)r;
        }
      }
      if (out && p.act) store4<OT>(out + oc, p.xt, g.nx, res);
      um = uc; uc = up;
      S.xz_m = mid4(S.xz_c); S.yz_m = S.yz_c; S.zz_m = S.zz_c;
      S.xz_c = S.xz_p; S.yz_c = S.yz_p; S.zz_c = S.zz_p;
    }
  }
  if (MODE == MODE_RES && partials) {
    const double t = block_sum(sq);
    if (threadIdx.x == 0 && threadIdx.y == 0)
      partials[(size_t)blockIdx.x + (size_t)gridDim.x * (blockIdx.y + (size_t)gridDim.y * blockIdx.z)] = t;
  }
}

}  // namespace fast
}  // namespace mad
