// mad_fast.cuh -- the streaming 3-D kernels of libmadgpu (sm_100a): one thread = 4 consecutive x
// voxels (128-bit loads/stores), one warp = 128 voxels of one image row, one CTA = WY consecutive
// rows, marching along z over a chunk of planes with the three live planes of u and of the
// z-differentiated tensor components (xz, yz, zz) held in REGISTERS.  x-neighbours come from warp
// shuffles (plus one scalar load at the two warp ends), y-neighbour rows are re-loaded through L1
// (the neighbouring warp of the same CTA loads the same lines at the same time), so every field is
// fetched from HBM once per sweep: 36 B/voxel (u, f, u', six tensor planes) -- ncu: 36.05 B/voxel.
//
// Latency is hidden by software pipelining, not by occupancy (the marching state costs ~170
// registers): a plane step first consumes the raw registers loaded during the previous step,
// immediately ISSUES every load of the next step (~11 KB per warp in flight, nothing in between
// depends on a loaded value) and only then does the arithmetic.
//
// The operator row is evaluated on the fly in the closed form of row_coeffs()/apply_offdiag()
// (mad_kernels.cuh; reference: mad/itkGridsHierarchy.hxx:298-516) with node-mirrored reads at the
// Neumann boundary and the one-sided tensor differences of mad/itkGridsHierarchy.hxx:451-470.
//
//   k_fast_sweep<MODE_WJ>   mad/itkMultigridWeightedJacobiSmoother.hxx:33-102
//   k_fast_sweep<MODE_RES>  mad/itkMultigridGaussSeidelSmoother.hxx:114-180 (+ L2Norm partial sums,
//                           itkMultigridAnisotropicDiffusionImageFilter.hxx:496-515)
#pragma once
#include <cuda_fp16.h>

#include "mad_kernels.cuh"

namespace mad {
namespace fast {

constexpr unsigned FULL = 0xffffffffu;
constexpr int TX = 128;  // voxels per warp row

// L2 prefetch without a register destination (PTX); a no-op in the CPU test build (tests/mad_host/)
#ifndef MAD_HOST_EMULATION
__device__ __forceinline__ void mad_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#else
inline void mad_prefetch_l2(const void*) {}
#endif

template <typename T>
struct V4 {
  T v[4];
};
template <typename T>
struct V6 {
  T v[6];  // v[0] = x-1, v[1..4] = the thread's four voxels, v[5] = x+4
};

// Per-thread position inside the volume.  All element offsets are 32-bit (the host only selects
// these kernels for levels with fewer than 2^31 elements per field).
struct Pos {
  int xt;    // x of the thread's first voxel
  int xl;    // x the thread loads from: xt, or 0 in lanes right of the image (their results are discarded)
  int lane, y, jl;
  int dh;    // halo voxel fetched by the warp-end lanes, relative to xl: x-1 in lane 0, x+4 in lane 31
  bool edge; // lane 0 or 31
  bool xb;   // this thread holds x = 0 or x = nx-1
  bool ylo, yhi;
  int oym, oyp;  // element offsets of the mirrored y-1 / y+1 rows relative to row y
};

__device__ __forceinline__ Pos make_pos(const Geom& g)
{
  Pos p;
  p.lane = threadIdx.x;
  p.xt = blockIdx.x * TX + p.lane * 4;
  p.y = blockIdx.y * blockDim.y + threadIdx.y;
  p.xl = p.xt < g.nx ? p.xt : 0;
  p.jl = g.nx - 1 - p.xt;
  p.edge = p.lane == 0 || p.lane == 31;
  p.dh = (p.lane == 0 ? max(p.xt - 1, 0) : min(p.xt + 4, g.nx - 1)) - p.xl;
  p.xb = p.xt == 0 || (p.jl >= 0 && p.jl < 4);
  p.ylo = p.y == 0;
  p.yhi = p.y == g.ny - 1;
  p.oym = p.ylo ? g.pitch : -g.pitch;
  p.oyp = p.yhi ? -g.pitch : g.pitch;
  return p;
}

// ---- two-phase loads: issue4/issue6 return the raw registers, finish4/finish6 convert and exchange
// ---- the x-neighbours.  `o` is the element offset of the thread's first voxel.
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

__device__ __forceinline__ Raw4<float> issue4(const float* __restrict__ base, int o)
{
  Raw4<float> r;
  r.q = __ldg(reinterpret_cast<const float4*>(base + o));
  return r;
}
__device__ __forceinline__ Raw4<double> issue4(const double* __restrict__ base, int o)
{
  Raw4<double> r;
  r.a = __ldg(reinterpret_cast<const double2*>(base + o));
  r.b = __ldg(reinterpret_cast<const double2*>(base + o) + 1);
  return r;
}
template <typename ST>
__device__ __forceinline__ Raw6<ST> issue6(const ST* __restrict__ base, int o, const Pos& p)
{
  Raw6<ST> r;
  r.c = issue4(base, o);
  r.h = ST(0);
  if (p.edge) r.h = __ldg(base + o + p.dh);
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

// One-sided x-difference of a tensor component for the slot that lies on the x boundary
// (mad/itkGridsHierarchy.hxx:451-470); only called by threads with p.xb.
template <typename T>
__device__ __forceinline__ void xdiff_boundary(const V6<float>& w, const Pos& p, const float* __restrict__ base, int o, T d[4])
{
  if (p.xt == 0) d[0] = T(-3) * T(w.v[1]) + T(4) * T(w.v[2]) - T(w.v[3]);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (p.jl == j) {
      const T m2 = j >= 1 ? T(w.v[j - 1]) : T(__ldg(base + o - 2));
      d[j] = T(3) * T(w.v[j + 1]) - T(4) * T(w.v[j]) + m2;
    }
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
__device__ __forceinline__ void store4(OT* __restrict__ base, int o, int xt, int nx, const float v[4])
{
  if (xt + 3 < nx) {
    if constexpr (sizeof(OT) == 4) *reinterpret_cast<float4*>(base + o) = make_float4(v[0], v[1], v[2], v[3]);
    else {
      *reinterpret_cast<double2*>(base + o) = make_double2(v[0], v[1]);
      *(reinterpret_cast<double2*>(base + o) + 1) = make_double2(v[2], v[3]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (xt + j < nx) base[o + j] = OT(v[j]);
  }
}

// L2 prefetch of the rows a warp will need `dist` planes ahead: the marching state leaves room for only
// ~12 warps per SM, far too few to cover DRAM latency with loads that have a register destination, so the
// HBM fetch is started early with prefetch.global.L2 (no destination) and the real loads hit L2.
// One instruction per plane step covers the warp's 128-voxel row of all eight fp32 fields (u, f, six tensor
// planes): lane = field * 4 + 128-byte line.  fp64 fields (level-0 outer residual) need a second one.
struct Prefetch {
  const char* a;     // lanes 0..31: field (lane >> 2), line (lane & 3)
  const char* b;     // second half of the 1 KB rows of fp64 fields (null for fp32 fields)
  long long stride;  // bytes per plane of this lane's field
};
template <typename UT, typename FT>
__device__ __forceinline__ Prefetch make_prefetch(const Geom& g, const Tensor& D, const UT* u, const FT* f, int lane, int y)
{
  const int fid = lane >> 2, ln = lane & 3;
  const size_t es = fid == 0 ? sizeof(UT) : fid == 1 ? sizeof(FT) : 4;
  const char* base = fid == 0 ? reinterpret_cast<const char*>(u) : fid == 1 ? reinterpret_cast<const char*>(f)
                                                                           : reinterpret_cast<const char*>(D.p[fid - 2]);
  Prefetch P;
  P.a = base + ((size_t)y * g.pitch + (size_t)blockIdx.x * TX) * es + ln * 128;
  P.b = es == 8 ? P.a + 512 : nullptr;
  P.stride = g.plane * (long long)es;
  return P;
}
__device__ __forceinline__ void prefetch_plane(const Prefetch& P, int z)
{
  const char* q = P.a + P.stride * z;
  mad_prefetch_l2((q));
  if (P.b) mad_prefetch_l2((P.b + P.stride * z));
}

// first / last plane of a slab: the same row also goes to the neighbour's ghost plane (peer store over NVLink)
template <typename OT>
__device__ __forceinline__ void store_ghosts(const Geom& g, int z, int rowoff, int xt, const float v[4])
{
  if (z == 0 && g.glo) store4<OT>(reinterpret_cast<OT*>(g.glo), rowoff, xt, g.nx, v);
  if (z == g.nz - 1 && g.ghi) store4<OT>(reinterpret_cast<OT*>(g.ghi), rowoff, xt, g.nx, v);
}

__device__ __forceinline__ float fast_div(float a, float b) { return __fdividef(a, b); }
__device__ __forceinline__ double fast_div(double a, double b) { return a / b; }

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
__device__ __forceinline__ DzRaw issue_dz(const Tensor& D, int o, const Pos& p)
{
  DzRaw r;
  r.xz = issue6(D.p[XZ3], o, p);
  r.yz = issue4(D.p[YZ3], o);
  r.zz = issue4(D.p[ZZ3], o);
  return r;
}

// tensor rows that are only needed on the plane being updated
struct DcRaw {
  Raw6<float> xx, xy;
  Raw4<float> xy_m, xy_p, yy, yy_m, yy_p, yz_ym, yz_yp;
};
__device__ __forceinline__ DcRaw issue_dc(const Tensor& D, int o, const Pos& p)
{
  DcRaw r;
  const int om = o + p.oym, op = o + p.oyp;
  r.xx = issue6(D.p[XX3], o, p);
  r.xy = issue6(D.p[XY3], o, p);
  r.xy_m = issue4(D.p[XY3], om);
  r.xy_p = issue4(D.p[XY3], op);
  r.yy = issue4(D.p[YY3], o);
  r.yy_m = issue4(D.p[YY3], om);
  r.yy_p = issue4(D.p[YY3], op);
  r.yz_ym = issue4(D.p[YZ3], om);
  r.yz_yp = issue4(D.p[YZ3], op);
  return r;
}
struct DcFin {
  V6<float> xx, xy;
  V4<float> xy_m, xy_p, yy, yy_m, yy_p, yz_ym, yz_yp;
};
__device__ __forceinline__ DcFin finish_dc(const DcRaw& R, const Pos& p)
{
  DcFin F;
  F.xx = finish6<float, float>(R.xx, p); F.xy = finish6<float, float>(R.xy, p);
  F.xy_m = finish4<float>(R.xy_m); F.xy_p = finish4<float>(R.xy_p);
  F.yy = finish4<float>(R.yy); F.yy_m = finish4<float>(R.yy_m); F.yy_p = finish4<float>(R.yy_p);
  F.yz_ym = finish4<float>(R.yz_ym); F.yz_yp = finish4<float>(R.yz_yp);
  return F;
}

// Coefficients of the four rows of A at plane z (o = element offset of the thread's first voxel).
template <typename T>
__device__ __forceinline__ void coefficients(const Geom& g, const Tensor& D, const Pos& p, int z, int o, const DzState& S, const DcFin& F,
                                             Coef<T>& c)
{
  const GeomConst<T> k(g);
  // y-differences (central; one-sided on the first / last row -- warp-uniform branches)
  V4<T> dy_xy = sub4<T>(F.xy_p, F.xy_m), dy_yy = sub4<T>(F.yy_p, F.yy_m), dy_yz = sub4<T>(F.yz_yp, F.yz_ym);
  if (p.ylo || p.yhi) {
    const int o2 = o + (p.ylo ? 2 * g.pitch : -2 * g.pitch);
    const T sgn = p.ylo ? T(-1) : T(1);
    dy_xy = onesided4<T>(mid4(F.xy), sel4(p.ylo, F.xy_p, F.xy_m), finish4<float>(issue4(D.p[XY3], o2)), sgn);
    dy_yy = onesided4<T>(F.yy, sel4(p.ylo, F.yy_p, F.yy_m), finish4<float>(issue4(D.p[YY3], o2)), sgn);
    dy_yz = onesided4<T>(S.yz_c, sel4(p.ylo, F.yz_yp, F.yz_ym), finish4<float>(issue4(D.p[YZ3], o2)), sgn);
  }
  // z-differences
  V4<T> dz_xz = sub4<T>(mid4(S.xz_p), S.xz_m), dz_yz = sub4<T>(S.yz_p, S.yz_m), dz_zz = sub4<T>(S.zz_p, S.zz_m);
  const bool zlo = z == 0 && g.zlo_phys, zhi = z == g.nz - 1 && g.zhi_phys;
  if (zlo || zhi) {
    const int o2 = o + (int)(zlo ? 2 * g.plane : -2 * g.plane);
    const T sgn = zlo ? T(-1) : T(1);
    dz_xz = onesided4<T>(mid4(S.xz_c), sel4(zlo, mid4(S.xz_p), S.xz_m), finish4<float>(issue4(D.p[XZ3], o2)), sgn);
    dz_yz = onesided4<T>(S.yz_c, sel4(zlo, S.yz_p, S.yz_m), finish4<float>(issue4(D.p[YZ3], o2)), sgn);
    dz_zz = onesided4<T>(S.zz_c, sel4(zlo, S.zz_p, S.zz_m), finish4<float>(issue4(D.p[ZZ3], o2)), sgn);
  }
  // x-differences (central; the slot on the x boundary is redone one-sided by the few threads that hold it)
  T dx_xx[4], dx_xy[4], dx_xz[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    dx_xx[j] = T(F.xx.v[j + 2]) - T(F.xx.v[j]);
    dx_xy[j] = T(F.xy.v[j + 2]) - T(F.xy.v[j]);
    dx_xz[j] = T(S.xz_c.v[j + 2]) - T(S.xz_c.v[j]);
  }
  if (p.xb) {
    xdiff_boundary<T>(F.xx, p, D.p[XX3], o, dx_xx);
    xdiff_boundary<T>(F.xy, p, D.p[XY3], o, dx_xy);
    xdiff_boundary<T>(S.xz_c, p, D.p[XZ3], o, dx_xz);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const T ax = k.wx * T(F.xx.v[j + 1]), ay = k.wy * T(F.yy.v[j]), az = k.wz * T(S.zz_c.v[j]);
    c.diag[j] = T(1) + T(2) * (ax + ay + az);
    const T bx = -(k.bxx * dx_xx[j] + k.bxy * dy_xy.v[j] + k.bxz * dz_xz.v[j]);
    const T by = -(k.bxy * dx_xy[j] + k.byy * dy_yy.v[j] + k.byz * dz_yz.v[j]);
    const T bz = -(k.bxz * dx_xz[j] + k.byz * dy_yz.v[j] + k.bzz * dz_zz.v[j]);
    c.xp[j] = -ax + bx; c.xm[j] = -ax - bx;
    c.yp[j] = -ay + by; c.ym[j] = -ay - by;
    c.zp[j] = -az + bz; c.zm[j] = -az - bz;
    c.exy[j] = -k.cxy * T(F.xy.v[j + 1]);
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
__device__ __forceinline__ URaw<UT> issue_u(const UT* __restrict__ u, int o, const Pos& p)
{
  URaw<UT> R;
  R.r[0] = issue6(u, o + p.oym, p);
  R.r[1] = issue6(u, o, p);
  R.r[2] = issue6(u, o + p.oyp, p);
  return R;
}
template <typename T, typename UT>
__device__ __forceinline__ void finish_u(const URaw<UT>& R, const Pos& p, UPlane<T>& P)
{
#pragma unroll
  for (int i = 0; i < 3; ++i) P.r[i] = finish6<T, UT>(R.r[i], p);
  if (p.xb) {
#pragma unroll
    for (int i = 0; i < 3; ++i) mirror_x(P.r[i], p.xt, p.jl);
  }
}

template <typename T>
__device__ __forceinline__ void zero_uplane(UPlane<T>& P)
{
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 6; ++k) P.r[i].v[k] = T(0);
}

// sum over the off-diagonal entries at slot j (apply_offdiag of mad_kernels.cuh on registers)
// CT: the type the coefficients were evaluated in (T, or float under a double accumulation: MODE_RES_C32)
template <typename T, typename CT>
__device__ __forceinline__ T offdiag(const Coef<CT>& c, const UPlane<T>& m, const UPlane<T>& q, const UPlane<T>& n, int j)
{
  T s = T(c.xp[j]) * q.r[1].v[j + 2] + T(c.xm[j]) * q.r[1].v[j] + T(c.yp[j]) * q.r[2].v[j + 1] + T(c.ym[j]) * q.r[0].v[j + 1];
  s += T(c.exy[j]) * ((q.r[2].v[j + 2] - q.r[0].v[j + 2]) - (q.r[2].v[j] - q.r[0].v[j]));
  s += T(c.zp[j]) * n.r[1].v[j + 1] + T(c.zm[j]) * m.r[1].v[j + 1];
  s += T(c.exz[j]) * ((n.r[1].v[j + 2] - m.r[1].v[j + 2]) - (n.r[1].v[j] - m.r[1].v[j]));
  s += T(c.eyz[j]) * ((n.r[2].v[j + 1] - m.r[2].v[j + 1]) - (n.r[0].v[j + 1] - m.r[0].v[j + 1]));
  return s;
}

__device__ __forceinline__ int zmirror_lo(const Geom& g, int z) { return (z == 0 && g.zlo_phys) ? 1 : z - 1; }
__device__ __forceinline__ int zmirror_hi(const Geom& g, int z) { return (z == g.nz - 1 && g.zhi_phys) ? g.nz - 2 : z + 1; }

// every load of one plane step: u and the z-marched tensor rows of plane z+1, the plane-z-only tensor rows, f
template <typename UT, typename FT>
struct StepRaw {
  URaw<UT> u;
  DzRaw dz;
  DcRaw dc;
  Raw4<FT> f;
};
template <typename UT, typename FT>
__device__ __forceinline__ StepRaw<UT, FT> issue_step(const Geom& g, const Tensor& D, const UT* __restrict__ u, const FT* __restrict__ f,
                                                      const Pos& p, int rowo, int z)
{
  StepRaw<UT, FT> R;
  const int oc = z * (int)g.plane + rowo, on = zmirror_hi(g, z) * (int)g.plane + rowo;
  R.u = issue_u(u, on, p);
  R.dz = issue_dz(D, on, p);
  R.dc = issue_dc(D, oc, p);
  R.f = issue4(f, oc);
  return R;
}

// MODE_RES_C32: MODE_RES with the operator row evaluated in fp32 (exactly the row the fp32 sweeps use) and applied in T -- for the
// fp64 stop-test residual, whose fp64 row evaluation (conversions + FP64 pipe) is what bounds it; opt-in, MADGPU_RES64_COEF32=1
enum { MODE_WJ = 0, MODE_RES = 1, MODE_COEF = 2, MODE_RES_C32 = 3 };

// Packed operator rows for the Gauss-Seidel smoother: per group of four x-voxels ten fp16 values per voxel, 80 bytes =
// five 16-byte words; word i holds entries 2i and 2i+1, each for the four voxels.  Entries: 0 1/diag, then the
// off-diagonal coefficients DIVIDED by diag: 1 x+, 2 x-, 3 y+, 4 y-, 5 z+, 6 z-, 7 xy edges, 8 xz edges, 9 yz edges.
constexpr int COEF_WORDS = 5;
// Entry 0 is stored as (1/diag) * 2^14: diag = 1 + 2 dt sum_d D_dd / h_d^2 >= 1 grows with dt / h^2 (small spacings, large time
// steps), and a plain fp16 1/diag would lose bits below 6.1e-5 (diag > 1.6e4) and flush to zero beyond 3.4e7.  Scaled by a power
// of two (exact) the entry stays a normal fp16 number for diag < 2^28; the packing kernel reports rows beyond COEF_DIAG_MAX so that
// the host keeps the exact-row sweep for such a level instead of relaxing with a denormal diagonal.
constexpr float COEF_INV_SCALE = 16384.f, COEF_INV_UNSCALE = 1.f / 16384.f, COEF_DIAG_MAX = 1.0e8f;
__device__ __forceinline__ size_t coef_quad(const Geom& g, int x4, int y, int z) { return ((size_t)z * g.ny + y) * (size_t)(g.pitch >> 2) + x4; }

// One pass over the volume.  MODE_WJ: out = weighted-Jacobi update of u.  MODE_RES: out = f - A u
// (out may be null) and per-CTA partial sums of its squares (partials may be null).
// grid = (ceil(nx/128), ceil(ny/WY), ceil(nz/zc)), block = (32, WY); MINB = CTAs per SM the register
// allocation is capped for; PF = issue the loads of step z+1 before the arithmetic of step z (needs ~70 more
// registers) instead of at the top of step z+1.
template <int MODE, typename T, typename UT, typename FT, typename OT, int WY, int MINB, bool PF>
__global__ void __launch_bounds__(32 * WY, MINB) k_fast_sweep(Geom g, Tensor D, const UT* __restrict__ u, const FT* __restrict__ f,
                                                         OT* __restrict__ out, double* __restrict__ partials, float omega, int zc, int pfd, int uzero)
{
  // uzero: the iterate is identically zero (first sweep of a V-cycle leg): `u` is not read
  constexpr bool RES = MODE == MODE_RES || MODE == MODE_RES_C32;
  typedef typename std::conditional<MODE == MODE_RES_C32, float, T>::type CT;  // type the row is evaluated in
  const Pos p = make_pos(g);
  const bool valid = p.y < g.ny;
  double sq = 0.0;
  if (valid) {
    const int z0 = blockIdx.z * zc, z1 = min(z0 + zc, g.nz);
    const int rowo = p.y * g.pitch + p.xl;
    const Prefetch PFL = make_prefetch(g, D, u, f, p.lane, p.y);
    const int zpf_end = min(z1 + 1, g.nz);  // planes 0..nz-1 exist for every field
    const T om = T(omega), om1 = T(1) - T(omega);
    UPlane<T> um, uc, up;
    DzState S;
    {
      const int o_m = zmirror_lo(g, z0) * (int)g.plane + rowo, o_c = z0 * (int)g.plane + rowo;
      const DzRaw d0 = issue_dz(D, o_m, p), d1 = issue_dz(D, o_c, p);
      if (uzero) { zero_uplane(um); zero_uplane(uc); }
      else {
        const URaw<UT> r0 = issue_u(u, o_m, p), r1 = issue_u(u, o_c, p);
        finish_u<T, UT>(r0, p, um);
        finish_u<T, UT>(r1, p, uc);
      }
      S.xz_m = finish4<float>(d0.xz.c); S.yz_m = finish4<float>(d0.yz); S.zz_m = finish4<float>(d0.zz);
      S.xz_c = finish6<float, float>(d1.xz, p); S.yz_c = finish4<float>(d1.yz); S.zz_c = finish4<float>(d1.zz);
    }
    StepRaw<UT, FT> R;
    if (PF) R = issue_step<UT, FT>(g, D, u, f, p, rowo, z0);
#pragma unroll 2
    for (int z = z0; z < z1; ++z) {
      const int oc = z * (int)g.plane + rowo;
      if (pfd > 0 && z + pfd < zpf_end) prefetch_plane(PFL, z + pfd);
      if (!PF) R = issue_step<UT, FT>(g, D, u, f, p, rowo, z);
      // ---- consume the loads issued one step ago ----
      finish_u<T, UT>(R.u, p, up);
      if (uzero) zero_uplane(up);
      S.xz_p = finish6<float, float>(R.dz.xz, p); S.yz_p = finish4<float>(R.dz.yz); S.zz_p = finish4<float>(R.dz.zz);
      const DcFin F = finish_dc(R.dc, p);
      const V4<T> fv = finish4<T>(R.f);
      // ---- issue every load of the next plane step before doing any arithmetic ----
      if (PF && z + 1 < z1) R = issue_step<UT, FT>(g, D, u, f, p, rowo, z + 1);
      Coef<CT> c;
      coefficients<CT>(g, D, p, z, oc, S, F, c);
      float res[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 4 && MODE != MODE_COEF; ++j) {
        const T s = offdiag<T>(c, um, uc, up, j);
        const T uj = uc.r[1].v[j + 1];
        if (MODE == MODE_WJ) {
          // mad/itkMultigridWeightedJacobiSmoother.hxx:88-89
          res[j] = float((fv.v[j] - s) * fast_div(om, T(c.diag[j])) + om1 * uj);
        } else {
          const T r = fv.v[j] - T(c.diag[j]) * uj - s;
          res[j] = float(r);
          if (p.xt + j < g.nx) sq += (double)r * (double)r;
        }
      }
      if (MODE != MODE_COEF && out && p.xt < g.nx) {
        store4<OT>(out, oc, p.xt, g.nx, res);
        store_ghosts<OT>(g, z, rowo, p.xt, res);
      }
      if (MODE == MODE_COEF && p.xt < g.nx) {
        __half2 h[10][2];
#pragma unroll
        for (int j = 0; j < 4; j += 2) {
          const float i0 = 1.f / float(c.diag[j]), i1 = 1.f / float(c.diag[j + 1]);
          h[0][j >> 1] = __floats2half2_rn(i0 * COEF_INV_SCALE, i1 * COEF_INV_SCALE);
          if (partials && (!(float(c.diag[j]) < COEF_DIAG_MAX) || !(float(c.diag[j + 1]) < COEF_DIAG_MAX))) partials[0] = 1.0;  // benign race: every writer stores 1
          h[1][j >> 1] = __floats2half2_rn(float(c.xp[j]) * i0, float(c.xp[j + 1]) * i1);
          h[2][j >> 1] = __floats2half2_rn(float(c.xm[j]) * i0, float(c.xm[j + 1]) * i1);
          h[3][j >> 1] = __floats2half2_rn(float(c.yp[j]) * i0, float(c.yp[j + 1]) * i1);
          h[4][j >> 1] = __floats2half2_rn(float(c.ym[j]) * i0, float(c.ym[j + 1]) * i1);
          h[5][j >> 1] = __floats2half2_rn(float(c.zp[j]) * i0, float(c.zp[j + 1]) * i1);
          h[6][j >> 1] = __floats2half2_rn(float(c.zm[j]) * i0, float(c.zm[j + 1]) * i1);
          h[7][j >> 1] = __floats2half2_rn(float(c.exy[j]) * i0, float(c.exy[j + 1]) * i1);
          h[8][j >> 1] = __floats2half2_rn(float(c.exz[j]) * i0, float(c.exz[j + 1]) * i1);
          h[9][j >> 1] = __floats2half2_rn(float(c.eyz[j]) * i0, float(c.eyz[j + 1]) * i1);
        }
        uint4* dst = reinterpret_cast<uint4*>(out) + coef_quad(g, p.xt >> 2, p.y, z) * COEF_WORDS;
#pragma unroll
        for (int i = 0; i < COEF_WORDS; ++i) {
          uint4 q;
          q.x = *reinterpret_cast<unsigned*>(&h[2 * i][0]); q.y = *reinterpret_cast<unsigned*>(&h[2 * i][1]);
          q.z = *reinterpret_cast<unsigned*>(&h[2 * i + 1][0]); q.w = *reinterpret_cast<unsigned*>(&h[2 * i + 1][1]);
          dst[i] = q;
        }
      }
      um = uc; uc = up;
      S.xz_m = mid4(S.xz_c); S.yz_m = S.yz_c; S.zz_m = S.zz_c;
      S.xz_c = S.xz_p; S.yz_c = S.yz_p; S.zz_c = S.zz_p;
    }
  }
  if (RES && partials) {
    const double t = block_sum(sq);
    if (threadIdx.x == 0 && threadIdx.y == 0)
      partials[(size_t)blockIdx.x + (size_t)gridDim.x * (blockIdx.y + (size_t)gridDim.y * blockIdx.z)] = t;
  }
}


// ------------------------------------------------------------------------------------------
// Level-0 fp64 defect / stop-test residual with the y-neighbour rows in SHARED MEMORY (k_fast_res64).
// MEASURED AND NOT THE DEFAULT (MADGPU_RES64_SMEM=3 selects it): 2.16 ms at 512^3 on B200 against 1.66 ms for the all-register form
// it was meant to replace -- 12 instead of 8 warps per SM, but a CTA-wide barrier per plane step and 0.6 short-scoreboard stalls per
// issue on the ring (profiles/r02h_res64_smem_512_full.txt).
//
// k_fast_sweep<MODE_RES_C32, double> keeps three planes x three rows x six fp64 values of u per thread: 255 registers, 8 warps
// per SM, and ncu shows it waiting on latency (issue slots 35 % busy) at 0.55 of the copy bandwidth.  Here a thread keeps only
// its OWN row of the three live planes (18 doubles); the rows y-1 / y+1 come from a ring of four planes of the CTA's tile of
// rows (WY rows + one halo row below and above) in shared memory, which every warp fills with the row it loads anyway (the
// first and the last warp also load the halo row).  The plane two steps ahead is loaded while the current one is evaluated;
// one __syncthreads per plane step.  Arithmetic as MODE_RES_C32: operator row evaluated in fp32 from the tensor planes (the
// row the fp32 sweeps relax), applied in fp64 to the fp64 iterate; out = fp32 residual, partials = per-CTA sums of squares.
// Row layout in shared memory: A[32] = the lanes' voxels 0,1 (double2), B[32] = voxels 2,3, then the two x-halo voxels of the
// warp ends -- 16-byte accesses at a stride of 16 bytes, no bank conflicts; x-neighbours by shuffle as everywhere else.
// grid = (ceil(nx/128), ceil(ny/WY), ceil(nz/zc)), block = (32, WY).
// ------------------------------------------------------------------------------------------
constexpr int R64_ROW = 130;  // doubles per tile row: 64 (A) + 64 (B) + halo left + halo right
template <int WY>
struct Res64Smem {
  double v[4][WY + 2][R64_ROW];
};
__device__ __forceinline__ void r64_store_row(double* __restrict__ row, const Raw6<double>& r, int lane)
{
  reinterpret_cast<double2*>(row)[lane] = r.c.a;
  reinterpret_cast<double2*>(row + 64)[lane] = r.c.b;
  if (lane == 0) row[128] = r.h;
  if (lane == 31) row[129] = r.h;
}
__device__ __forceinline__ Raw6<double> r64_load_row6(const double* __restrict__ row, int lane)
{
  Raw6<double> r;
  r.c.a = reinterpret_cast<const double2*>(row)[lane];
  r.c.b = reinterpret_cast<const double2*>(row + 64)[lane];
  r.h = lane == 0 ? row[128] : row[129];  // only the two warp-end lanes use it
  return r;
}
__device__ __forceinline__ V4<double> r64_load_row4(const double* __restrict__ row, int lane)
{
  const double2 a = reinterpret_cast<const double2*>(row)[lane], b = reinterpret_cast<const double2*>(row + 64)[lane];
  V4<double> w;
  w.v[0] = a.x; w.v[1] = a.y; w.v[2] = b.x; w.v[3] = b.y;
  return w;
}

template <int WY, int MINB>
__global__ void __launch_bounds__(32 * WY, MINB) k_fast_res64(Geom g, Tensor D, const double* __restrict__ u, const double* __restrict__ f,
                                                                float* __restrict__ out, double* __restrict__ partials, int zc, int pfd)
{
  __shared__ Res64Smem<WY> sm;
  const Pos p = make_pos(g);
  const int w = threadIdx.y, lane = p.lane;
  const int y0t = blockIdx.y * WY;
  const bool valid = p.y < g.ny;
  const int z0 = blockIdx.z * zc, z1 = min(z0 + zc, g.nz);
  // tile rows (slot 0 = image row y0t - 1) that hold this warp's mirrored y-1 / y+1 rows: row -1 is row 1, row ny is row ny - 2
  const int rm = p.y == 0 ? 2 : w, rp = p.y == g.ny - 1 ? w : w + 2;
  // the halo rows of the tile: loaded by the first / last warp when they exist
  const bool halo_lo = w == 0 && y0t > 0, halo_hi = w == WY - 1 && y0t + WY < g.ny;
  const int rowo = (valid ? p.y : 0) * g.pitch + p.xl;
  const int rowo_lo = (y0t - 1) * g.pitch + p.xl, rowo_hi = (y0t + WY) * g.pitch + p.xl;
  // logical plane L (z0 - 1 .. z1) -> ring slot; its physical plane is the node mirror at the two physical ends
  auto slot_of = [&](int L) { return (L - (z0 - 1)) & 3; };
  auto phys = [&](int L) { return L < z0 ? zmirror_lo(g, z0) : (L >= g.nz ? zmirror_hi(g, g.nz - 1) : L); };
  struct PlaneRaw { Raw6<double> own, lo, hi; };
  auto issue_plane = [&](int L) {
    PlaneRaw r;
    const int b = phys(L) * (int)g.plane;
    r.own.c.a = r.own.c.b = make_double2(0.0, 0.0);
    r.own.h = 0.0;
    if (valid) r.own = issue6(u, b + rowo, p);
    r.lo = r.own; r.hi = r.own;
    if (halo_lo) r.lo = issue6(u, b + rowo_lo, p);
    if (halo_hi) r.hi = issue6(u, b + rowo_hi, p);
    return r;
  };
  auto publish_plane = [&](int L, const PlaneRaw& r) {  // raw rows -> the ring; returns nothing, the own row is finished by the caller
    const int s = slot_of(L);
    if (valid) r64_store_row(sm.v[s][w + 1], r.own, lane);
    if (halo_lo) r64_store_row(sm.v[s][0], r.lo, lane);
    if (halo_hi) r64_store_row(sm.v[s][WY + 1], r.hi, lane);
  };
  auto own_row = [&](const PlaneRaw& r) {
    V6<double> v = finish6<double, double>(r.own, p);
    if (p.xb) mirror_x(v, p.xt, p.jl);
    return v;
  };
  auto nb_row6 = [&](int L, int r) {
    V6<double> v = finish6<double, double>(r64_load_row6(sm.v[slot_of(L)][r], lane), p);
    if (p.xb) mirror_x(v, p.xt, p.jl);
    return v;
  };
  double sq = 0.0;
  const Prefetch PFL = make_prefetch(g, D, u, f, lane, valid ? p.y : 0);
  const int zpf_end = min(z1 + 1, g.nz);
  V6<double> um, uc, up;
  DzState S;
  {
    const PlaneRaw r0 = issue_plane(z0 - 1), r1 = issue_plane(z0), r2 = issue_plane(z0 + 1);
    const int o_m = zmirror_lo(g, z0) * (int)g.plane + rowo, o_c = z0 * (int)g.plane + rowo;
    publish_plane(z0 - 1, r0); publish_plane(z0, r1); publish_plane(z0 + 1, r2);
    um = own_row(r0); uc = own_row(r1); up = own_row(r2);
    if (valid) {
      const DzRaw d0 = issue_dz(D, o_m, p), d1 = issue_dz(D, o_c, p);
      S.xz_m = finish4<float>(d0.xz.c); S.yz_m = finish4<float>(d0.yz); S.zz_m = finish4<float>(d0.zz);
      S.xz_c = finish6<float, float>(d1.xz, p); S.yz_c = finish4<float>(d1.yz); S.zz_c = finish4<float>(d1.zz);
    }
  }
  __syncthreads();
  for (int z = z0; z < z1; ++z) {
    const int oc = z * (int)g.plane + rowo, on = zmirror_hi(g, z) * (int)g.plane + rowo;
    if (pfd > 0 && z + pfd < zpf_end && valid) prefetch_plane(PFL, z + pfd);
    // ---- every load of the step: u of the plane after next, the tensor rows of this step, f ----
    const bool more = z + 1 < z1;
    PlaneRaw rn;
    if (more) rn = issue_plane(z + 2);
    if (valid) {
      const DzRaw dz = issue_dz(D, on, p);
      const DcRaw dc = issue_dc(D, oc, p);
      const Raw4<double> rf = issue4(f, oc);
      S.xz_p = finish6<float, float>(dz.xz, p); S.yz_p = finish4<float>(dz.yz); S.zz_p = finish4<float>(dz.zz);
      const DcFin F = finish_dc(dc, p);
      const V4<double> fv = finish4<double>(rf);
      Coef<float> c;
      coefficients<float>(g, D, p, z, oc, S, F, c);
      double acc[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {  // own rows of the three planes
        double s = (double)c.xp[j] * uc.v[j + 2] + (double)c.xm[j] * uc.v[j] + (double)c.zp[j] * up.v[j + 1] + (double)c.zm[j] * um.v[j + 1];
        s += (double)c.exz[j] * ((up.v[j + 2] - um.v[j + 2]) - (up.v[j] - um.v[j]));
        acc[j] = fv.v[j] - (double)c.diag[j] * uc.v[j + 1] - s;
      }
      {  // rows y+1, y-1 of this plane
        const V6<double> q2 = nb_row6(z, rp), q0 = nb_row6(z, rm);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          acc[j] -= (double)c.yp[j] * q2.v[j + 1] + (double)c.ym[j] * q0.v[j + 1] + (double)c.exy[j] * ((q2.v[j + 2] - q0.v[j + 2]) - (q2.v[j] - q0.v[j]));
      }
      {  // rows y+1, y-1 of the planes above and below
        const V4<double> n2 = r64_load_row4(sm.v[slot_of(z + 1)][rp], lane), m2 = r64_load_row4(sm.v[slot_of(z - 1)][rp], lane);
        const V4<double> n0 = r64_load_row4(sm.v[slot_of(z + 1)][rm], lane), m0 = r64_load_row4(sm.v[slot_of(z - 1)][rm], lane);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] -= (double)c.eyz[j] * ((n2.v[j] - m2.v[j]) - (n0.v[j] - m0.v[j]));
      }
      float res[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        res[j] = (float)acc[j];
        if (p.xt + j < g.nx) sq += acc[j] * acc[j];
      }
      if (out && p.xt < g.nx) {
        store4<float>(out, oc, p.xt, g.nx, res);
        store_ghosts<float>(g, z, rowo, p.xt, res);
      }
    }
    // ---- the plane after next joins the ring (its slot was last read one step ago) ----
    um = uc; uc = up;
    if (more) {
      publish_plane(z + 2, rn);
      up = own_row(rn);
    }
    if (valid) {
      S.xz_m = mid4(S.xz_c); S.yz_m = S.yz_c; S.zz_m = S.zz_c;
      S.xz_c = S.xz_p; S.yz_c = S.yz_p; S.zz_c = S.zz_p;
    }
    __syncthreads();
  }
  if (partials) {
    const double t = block_sum(sq);
    if (threadIdx.x == 0 && threadIdx.y == 0)
      partials[(size_t)blockIdx.x + (size_t)gridDim.x * (blockIdx.y + (size_t)gridDim.y * blockIdx.z)] = t;
  }
}

// ------------------------------------------------------------------------------------------
// Gauss-Seidel sweep (mad/itkMultigridGaussSeidelSmoother.hxx:33-111) in ONE pass over the volume.
// Ordering: planes in z order (as the reference's outer loop); inside a plane the even rows first
// (even x, then odd x), then the odd rows -- the four in-plane colours of the 9-point plane stencil,
// so no two voxels updated together are coupled and every update sees the newest value of every
// neighbour that precedes it: a true Gauss-Seidel ordering INSIDE a CTA tile (128 x WY x zc voxels).
// Values outside the tile are read from the previous sweep (`u`; the result goes to `out`), i.e. tile
// faces are relaxed Jacobi-style.  All orderings share the fixed point A u = f, which is what the
// parity bound for Gauss-Seidel is stated on; tools/gs_order_experiment.py shows the V-cycle
// convergence factor of this ordering next to the reference's lexicographic one.
//
// New values travel: own row -> registers; x-neighbours -> warp shuffles; rows y+-1 of the plane
// being updated and of the plane below -> shared memory (two row buffers, alternating with z).
// ------------------------------------------------------------------------------------------
#ifndef MAD_HOST_EMULATION
__device__ __forceinline__ void bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
#endif  // the CPU test build (tests/mad_host/) supplies named barriers of its own

template <int WY, int MINB, bool PF, bool SPLIT>
__global__ void __launch_bounds__(32 * WY, MINB) k_fast_gs(Geom g, Tensor D, const float* __restrict__ u, const float* __restrict__ f,
                                                             float* __restrict__ out, int zc, int pfd, int uzero)
{
  static_assert(WY % 2 == 0, "row colours alternate with the warp index");
  __shared__ float4 sh[2][WY][32];
  const Pos p = make_pos(g);
  const int w = threadIdx.y;
  const bool valid = p.y < g.ny;
  const int wm = p.ylo ? w + 1 : w - 1, wp = p.yhi ? w - 1 : w + 1;  // tile rows holding the (mirrored) y-1 / y+1 rows
  const bool has_m = wm >= 0 && wm < WY, has_p = wp >= 0 && wp < WY && blockIdx.y * WY + wp < g.ny;
  const int z0 = blockIdx.z * zc, z1 = min(z0 + zc, g.nz);
  const int rowo = (valid ? p.y : 0) * g.pitch + p.xl;
  const Prefetch PFL = make_prefetch(g, D, u, f, p.lane, valid ? p.y : 0);
  const int zpf_end = min(z1 + 1, g.nz);
  UPlane<float> um, uc, up;
  DzState S;
  {
    const int o_m = zmirror_lo(g, z0) * (int)g.plane + rowo, o_c = z0 * (int)g.plane + rowo;
    const DzRaw d0 = issue_dz(D, o_m, p), d1 = issue_dz(D, o_c, p);
    if (uzero) { zero_uplane(um); zero_uplane(uc); }
    else {
      const URaw<float> r0 = issue_u(u, o_m, p), r1 = issue_u(u, o_c, p);
      finish_u<float, float>(r0, p, um);
      finish_u<float, float>(r1, p, uc);
    }
    S.xz_m = finish4<float>(d0.xz.c); S.yz_m = finish4<float>(d0.yz); S.zz_m = finish4<float>(d0.zz);
    S.xz_c = finish6<float, float>(d1.xz, p); S.yz_c = finish4<float>(d1.yz); S.zz_c = finish4<float>(d1.zz);
  }
  StepRaw<float, float> R;
  if (PF) R = issue_step<float, float>(g, D, u, f, p, rowo, z0);
#pragma unroll 2
  for (int z = z0; z < z1; ++z) {
    const int oc = z * (int)g.plane + rowo;
    const int cb = z & 1, pb = cb ^ 1;
    if (pfd > 0 && z + pfd < zpf_end) prefetch_plane(PFL, z + pfd);
    if (!PF) R = issue_step<float, float>(g, D, u, f, p, rowo, z);
    finish_u<float, float>(R.u, p, up);
    if (uzero) zero_uplane(up);
    S.xz_p = finish6<float, float>(R.dz.xz, p); S.yz_p = finish4<float>(R.dz.yz); S.zz_p = finish4<float>(R.dz.zz);
    const DcFin F = finish_dc(R.dc, p);
    const V4<float> fv = finish4<float>(R.f);
    // software pipelining: the loads of the next plane fly while this plane's two phases run
    if (PF && z + 1 < z1) R = issue_step<float, float>(g, D, u, f, p, rowo, z + 1);
    Coef<float> c;
    coefficients<float>(g, D, p, z, oc, S, F, c);
    // rows y+-1 of the plane below (updated one step ago) and, for odd rows, of this plane (updated in this step's
    // first phase) come from the tile's row buffers
    auto rows_below = [&]() {
      if (has_m) { const float4 q = sh[pb][wm][p.lane]; um.r[0].v[1] = q.x; um.r[0].v[2] = q.y; um.r[0].v[3] = q.z; um.r[0].v[4] = q.w; }
      if (has_p) { const float4 q = sh[pb][wp][p.lane]; um.r[2].v[1] = q.x; um.r[2].v[2] = q.y; um.r[2].v[3] = q.z; um.r[2].v[4] = q.w; }
    };
    auto even_rows_of_this_plane = [&]() {
      const float* sm = reinterpret_cast<const float*>(&sh[cb][has_m ? wm : 0][0]);
      const float* sp = reinterpret_cast<const float*>(&sh[cb][has_p ? wp : 0][0]);
      if (has_m) {
        const float4 q = reinterpret_cast<const float4*>(sm)[p.lane];
        uc.r[0].v[1] = q.x; uc.r[0].v[2] = q.y; uc.r[0].v[3] = q.z; uc.r[0].v[4] = q.w;
        if (p.lane > 0) uc.r[0].v[0] = sm[p.lane * 4 - 1];
        if (p.lane < 31) uc.r[0].v[5] = sm[p.lane * 4 + 4];
      }
      if (has_p) {
        const float4 q = reinterpret_cast<const float4*>(sp)[p.lane];
        uc.r[2].v[1] = q.x; uc.r[2].v[2] = q.y; uc.r[2].v[3] = q.z; uc.r[2].v[4] = q.w;
        if (p.lane > 0) uc.r[2].v[0] = sp[p.lane * 4 - 1];
        if (p.lane < 31) uc.r[2].v[5] = sp[p.lane * 4 + 4];
      }
      if (p.xb) { mirror_x(uc.r[0], p.xt, p.jl); mirror_x(uc.r[2], p.xt, p.jl); }
    };
    auto update_row = [&]() {
      // even x (slots 0, 2): x-neighbours still hold the values of the previous sweep
      const float n0 = fast_div(fv.v[0] - offdiag<float>(c, um, uc, up, 0), c.diag[0]);  // mad/itkMultigridGaussSeidelSmoother.hxx:99
      const float n2 = fast_div(fv.v[2] - offdiag<float>(c, um, uc, up, 2), c.diag[2]);
      uc.r[1].v[1] = n0; uc.r[1].v[3] = n2;
      {
        const float r = __shfl_down_sync(FULL, n0, 1);
        if (p.lane < 31) uc.r[1].v[5] = r;
        if (p.xb) mirror_x(uc.r[1], p.xt, p.jl);
      }
      // odd x (slots 1, 3)
      const float n1 = fast_div(fv.v[1] - offdiag<float>(c, um, uc, up, 1), c.diag[1]);
      const float n3 = fast_div(fv.v[3] - offdiag<float>(c, um, uc, up, 3), c.diag[3]);
      uc.r[1].v[2] = n1; uc.r[1].v[4] = n3;
      {
        const float l = __shfl_up_sync(FULL, n3, 1);
        if (p.lane > 0) uc.r[1].v[0] = l;
        if (p.xb) mirror_x(uc.r[1], p.xt, p.jl);
      }
      sh[cb][w][p.lane] = make_float4(n0, n1, n2, n3);
      const float res[4] = {n0, n1, n2, n3};
      if (p.xt < g.nx) { store4<float>(out, oc, p.xt, g.nx, res); store_ghosts<float>(g, z, rowo, p.xt, res); }
    };
    const bool last_plane = z == g.nz - 1 && g.zhi_phys;  // the mirrored plane z+1 IS plane z-1, which has already been updated
    if (SPLIT) {
      // Producer/consumer barriers instead of CTA-wide ones: barrier 2 = "even rows of this plane published" (even-row
      // warps arrive, odd-row warps wait), barrier 1 = "odd rows published" (the other way round).  A warp that has
      // published runs ahead into the loads and the coefficient arithmetic of the next plane, which depend on no other
      // warp, while the other parity relaxes its rows.
      constexpr int NT = 32 * WY;
      const bool even = (w & 1) == 0;
      if (even) {
        if (z > z0) { bar_sync(1, NT); rows_below(); }
      } else {
        // (the even rows of the plane below are already in registers: read from the row buffer one step ago)
        bar_sync(2, NT);
        if (valid) even_rows_of_this_plane();
      }
      if (last_plane) up = um;
      if (valid) update_row();
      __threadfence_block();
      if (even) bar_arrive(2, NT);
      else bar_arrive(1, NT);
    } else {
      if (z > z0) rows_below();
      if (last_plane) up = um;
#pragma unroll 1
      for (int phase = 0; phase < 2; ++phase) {
        if (valid && (w & 1) == phase) {
          if (phase == 1) even_rows_of_this_plane();
          update_row();
        }
        __syncthreads();
      }
    }
    um = uc; uc = up;
    S.xz_m = mid4(S.xz_c); S.yz_m = S.yz_c; S.zz_m = S.zz_c;
    S.xz_c = S.xz_p; S.yz_c = S.yz_p; S.zz_c = S.zz_p;
  }
}

// The same fused Gauss-Seidel sweep fed by PRE-EVALUATED operator rows (k_fast_sweep<MODE_COEF>): ten fp16 values per
// voxel (20 B) instead of the six fp32 tensor planes (24 B) and ~2/3 of the instructions (no tensor differences, no tensor
// halo traffic).  fp16 rows make this a Gauss-Seidel sweep on an operator rounded to 11 bits -- a smoother just as good, and
// harmless to the result: every cycle is a correction to the fp64 defect f - A u evaluated with the exact rows, so the
// fixed point is the reference's (Gauss-Seidel parity is stated on the converged image).  Weighted Jacobi, whose parity is
// stated per V-cycle, keeps the exact on-the-fly rows.
struct CoefRaw {
  uint4 w[COEF_WORDS];
};
__device__ __forceinline__ CoefRaw issue_coef(const uint4* __restrict__ coef, const Geom& g, const Pos& p, int y, int z)
{
  CoefRaw r;
  const uint4* src = coef + coef_quad(g, p.xl >> 2, y, z) * COEF_WORDS;
#pragma unroll
  for (int i = 0; i < COEF_WORDS; ++i) r.w[i] = __ldg(src + i);
  return r;
}
// entry k (0..9) of voxel j (0..3)
__device__ __forceinline__ float coef_at(const CoefRaw& r, int k, int j)
{
  const uint4& q = r.w[k >> 1];
  const unsigned u = (k & 1) ? (j < 2 ? q.z : q.w) : (j < 2 ? q.x : q.y);
  const __half2 h = *reinterpret_cast<const __half2*>(&u);
  return (j & 1) ? __high2float(h) : __low2float(h);
}
// 1/diag of voxel j (entry 0 without its storage scale)
__device__ __forceinline__ float coef_inv(const CoefRaw& r, int j) { return coef_at(r, 0, j) * COEF_INV_UNSCALE; }
// normalised off-diagonal sum at slot j
__device__ __forceinline__ float offdiag16(const CoefRaw& c, const UPlane<float>& m, const UPlane<float>& q, const UPlane<float>& n, int j)
{
  float s = coef_at(c, 1, j) * q.r[1].v[j + 2] + coef_at(c, 2, j) * q.r[1].v[j] + coef_at(c, 3, j) * q.r[2].v[j + 1] + coef_at(c, 4, j) * q.r[0].v[j + 1];
  s += coef_at(c, 7, j) * ((q.r[2].v[j + 2] - q.r[0].v[j + 2]) - (q.r[2].v[j] - q.r[0].v[j]));
  s += coef_at(c, 5, j) * n.r[1].v[j + 1] + coef_at(c, 6, j) * m.r[1].v[j + 1];
  s += coef_at(c, 8, j) * ((n.r[1].v[j + 2] - m.r[1].v[j + 2]) - (n.r[1].v[j] - m.r[1].v[j]));
  s += coef_at(c, 9, j) * ((n.r[2].v[j + 1] - m.r[2].v[j + 1]) - (n.r[0].v[j + 1] - m.r[0].v[j + 1]));
  return s;
}

// L2 prefetch of one plane step of this kernel: u (4 lines), f (4 lines), packed rows (20 lines) of the warp's row
__device__ __forceinline__ const char* coef_prefetch_base(const Geom& g, const float* u, const float* f, const uint4* coef, int lane, int y)
{
  const size_t x0 = (size_t)blockIdx.x * TX;
  if (lane < 4) return reinterpret_cast<const char*>(u + (size_t)y * g.pitch + x0) + lane * 128;
  if (lane < 8) return reinterpret_cast<const char*>(f + (size_t)y * g.pitch + x0) + (lane - 4) * 128;
  if (lane < 28) return reinterpret_cast<const char*>(coef + ((size_t)y * (g.pitch >> 2) + (x0 >> 2)) * COEF_WORDS) + (lane - 8) * 128;
  return nullptr;
}

template <int WY, int MINB>
__global__ void __launch_bounds__(32 * WY, MINB) k_coef_gs(Geom g, const uint4* __restrict__ coef, const float* __restrict__ u,
                                                             const float* __restrict__ f, float* __restrict__ out, int zc, int pfd, int uzero)
{
  static_assert(WY % 2 == 0, "row colours alternate with the warp index");
  __shared__ float4 sh[2][WY][32];
  const Pos p = make_pos(g);
  const int w = threadIdx.y;
  const bool valid = p.y < g.ny;
  const int y = valid ? p.y : 0;
  const int wm = p.ylo ? w + 1 : w - 1, wp = p.yhi ? w - 1 : w + 1;
  const bool has_m = wm >= 0 && wm < WY, has_p = wp >= 0 && wp < WY && blockIdx.y * WY + wp < g.ny;
  const int z0 = blockIdx.z * zc, z1 = min(z0 + zc, g.nz);
  const int rowo = y * g.pitch + p.xl;
  const char* pf = coef_prefetch_base(g, u, f, coef, p.lane, y);
  const long long pf_stride = p.lane < 8 ? g.plane * 4ll : (long long)g.ny * (g.pitch >> 2) * COEF_WORDS * 16ll;
  const int zpf_end = min(z1 + 1, g.nz);
  UPlane<float> um, uc, up;
  if (uzero) { zero_uplane(um); zero_uplane(uc); }
  else {
    const URaw<float> r0 = issue_u(u, zmirror_lo(g, z0) * (int)g.plane + rowo, p), r1 = issue_u(u, z0 * (int)g.plane + rowo, p);
    finish_u<float, float>(r0, p, um);
    finish_u<float, float>(r1, p, uc);
  }
  for (int z = z0; z < z1; ++z) {
    const int oc = z * (int)g.plane + rowo;
    const int cb = z & 1, pb = cb ^ 1;
    if (pfd > 0 && z + pfd < zpf_end && pf && !(uzero && p.lane < 4)) mad_prefetch_l2((pf + pf_stride * (z + pfd)));
    const CoefRaw c = issue_coef(coef, g, p, y, z);
    const Raw4<float> rf = issue4(f, oc);
    if (uzero) zero_uplane(up);
    else {
      const URaw<float> ru = issue_u(u, zmirror_hi(g, z) * (int)g.plane + rowo, p);
      finish_u<float, float>(ru, p, up);
    }
    const V4<float> fv = finish4<float>(rf);
    if (z > z0) {
      if (has_m) { const float4 q = sh[pb][wm][p.lane]; um.r[0].v[1] = q.x; um.r[0].v[2] = q.y; um.r[0].v[3] = q.z; um.r[0].v[4] = q.w; }
      if (has_p) { const float4 q = sh[pb][wp][p.lane]; um.r[2].v[1] = q.x; um.r[2].v[2] = q.y; um.r[2].v[3] = q.z; um.r[2].v[4] = q.w; }
    }
    if (z == g.nz - 1 && g.zhi_phys) up = um;
#pragma unroll 1
    for (int phase = 0; phase < 2; ++phase) {
      if (valid && (w & 1) == phase) {
        if (phase == 1) {
          const float* sm = reinterpret_cast<const float*>(&sh[cb][has_m ? wm : 0][0]);
          const float* sp = reinterpret_cast<const float*>(&sh[cb][has_p ? wp : 0][0]);
          if (has_m) {
            const float4 q = reinterpret_cast<const float4*>(sm)[p.lane];
            uc.r[0].v[1] = q.x; uc.r[0].v[2] = q.y; uc.r[0].v[3] = q.z; uc.r[0].v[4] = q.w;
            if (p.lane > 0) uc.r[0].v[0] = sm[p.lane * 4 - 1];
            if (p.lane < 31) uc.r[0].v[5] = sm[p.lane * 4 + 4];
          }
          if (has_p) {
            const float4 q = reinterpret_cast<const float4*>(sp)[p.lane];
            uc.r[2].v[1] = q.x; uc.r[2].v[2] = q.y; uc.r[2].v[3] = q.z; uc.r[2].v[4] = q.w;
            if (p.lane > 0) uc.r[2].v[0] = sp[p.lane * 4 - 1];
            if (p.lane < 31) uc.r[2].v[5] = sp[p.lane * 4 + 4];
          }
          if (p.xb) { mirror_x(uc.r[0], p.xt, p.jl); mirror_x(uc.r[2], p.xt, p.jl); }
        }
        const float n0 = fv.v[0] * coef_inv(c, 0) - offdiag16(c, um, uc, up, 0);
        const float n2 = fv.v[2] * coef_inv(c, 2) - offdiag16(c, um, uc, up, 2);
        uc.r[1].v[1] = n0; uc.r[1].v[3] = n2;
        {
          const float r = __shfl_down_sync(FULL, n0, 1);
          if (p.lane < 31) uc.r[1].v[5] = r;
          if (p.xb) mirror_x(uc.r[1], p.xt, p.jl);
        }
        const float n1 = fv.v[1] * coef_inv(c, 1) - offdiag16(c, um, uc, up, 1);
        const float n3 = fv.v[3] * coef_inv(c, 3) - offdiag16(c, um, uc, up, 3);
        uc.r[1].v[2] = n1; uc.r[1].v[4] = n3;
        {
          const float l = __shfl_up_sync(FULL, n3, 1);
          if (p.lane > 0) uc.r[1].v[0] = l;
          if (p.xb) mirror_x(uc.r[1], p.xt, p.jl);
        }
        sh[cb][w][p.lane] = make_float4(n0, n1, n2, n3);
        const float res[4] = {n0, n1, n2, n3};
        if (p.xt < g.nx) { store4<float>(out, oc, p.xt, g.nx, res); store_ghosts<float>(g, z, rowo, p.xt, res); }
      }
      __syncthreads();
    }
    um = uc; uc = up;
  }
}

// Row-pair variant of k_coef_gs: one warp owns an even row and the odd row above it and relaxes them one after the other, so
// no warp idles while the other row parity is being relaxed (in k_coef_gs half of a CTA's warps wait at the phase barrier).
// Same ordering and tile semantics (tile = 128 x 2*WP x zc); needs an even ny.  Per plane step a thread loads the four rows
// y0-1 .. y0+2 of plane z+1 (two row loads per relaxed row instead of three), the packed rows and f of its two rows.
struct URows4 {
  V6<float> r[4];  // rows y0-1, y0 (even), y0+1 (odd), y0+2
};
__device__ __forceinline__ float offdiag16_rows(const CoefRaw& c, const URows4& m, const URows4& q, const URows4& n, int row, int j)
{
  // `row` = 1 (even row) or 2 (odd row): its y-neighbours are the slots row-1 and row+1
  float s = coef_at(c, 1, j) * q.r[row].v[j + 2] + coef_at(c, 2, j) * q.r[row].v[j] + coef_at(c, 3, j) * q.r[row + 1].v[j + 1] +
            coef_at(c, 4, j) * q.r[row - 1].v[j + 1];
  s += coef_at(c, 7, j) * ((q.r[row + 1].v[j + 2] - q.r[row - 1].v[j + 2]) - (q.r[row + 1].v[j] - q.r[row - 1].v[j]));
  s += coef_at(c, 5, j) * n.r[row].v[j + 1] + coef_at(c, 6, j) * m.r[row].v[j + 1];
  s += coef_at(c, 8, j) * ((n.r[row].v[j + 2] - m.r[row].v[j + 2]) - (n.r[row].v[j] - m.r[row].v[j]));
  s += coef_at(c, 9, j) * ((n.r[row + 1].v[j + 1] - m.r[row + 1].v[j + 1]) - (n.r[row - 1].v[j + 1] - m.r[row - 1].v[j + 1]));
  return s;
}

// PRIVATE: the tile is the warp's own row pair (128 x 2 x zc): the rows above and below always carry the previous sweep's values, so
// the warps of a CTA never exchange anything and the two CTA-wide barriers per plane step are gone.  Measured on B200 at 512^3
// (profiles/r02q_*): 0.68 instead of 0.78 ms per sweep -- 6.4 TB/s on the 32 B per voxel it moves, the rate of the barrier-free
// residual kernel; the four-times-denser tile faces cost one V-cycle in 23 to relres 1e-10 (6,6,6,6 instead of 6,6,6,5), the whole
// filter call still gets faster (0.297 against 0.305 s).  The default since then (MADGPU_GS_PRIVATE=2: alternating grid, see yshift).
template <int WP, int MINB, bool PRIVATE = false>
__global__ void __launch_bounds__(32 * WP, MINB) k_coef_gs2(Geom g, const uint4* __restrict__ coef, const float* __restrict__ u,
                                                              const float* __restrict__ f, float* __restrict__ out, int zc, int pfd, int uzero,
                                                              int yshift = 0)
{
  // yshift (PRIVATE only, 0 or 1): the row pairs start at odd rows -- (-1, 0), (1, 2), ..., (ny-1, ny), the two end pairs holding one
  // image row each -- so that a host which alternates it between sweeps moves the tile faces (MADGPU_GS_PRIVATE=2)
  __shared__ float4 sh[PRIVATE ? 1 : 2][PRIVATE ? 1 : 2 * WP][32];
  const int lane = threadIdx.x, w = threadIdx.y;
  Pos p;
  p.lane = lane;
  p.xt = blockIdx.x * TX + lane * 4;
  p.xl = p.xt < g.nx ? p.xt : 0;
  p.jl = g.nx - 1 - p.xt;
  p.edge = lane == 0 || lane == 31;
  p.dh = (lane == 0 ? max(p.xt - 1, 0) : min(p.xt + 4, g.nx - 1)) - p.xl;
  p.xb = p.xt == 0 || (p.jl >= 0 && p.jl < 4);
  const int ytile = blockIdx.y * 2 * WP;
  const int y0 = ytile + 2 * w - (PRIVATE ? yshift : 0);  // first row of this warp's pair; ny is even
  // which of the pair's two rows exist (both, except in the two end pairs of a shifted grid)
  const bool relaxB = y0 >= 0 && y0 < g.ny, relaxC = y0 + 1 >= 0 && y0 + 1 < g.ny;
  const bool valid = relaxB || relaxC;
  // image rows behind the four slots: node mirror at the y ends (row -1 is row 1, row ny is row ny - 2)
  auto mir = [&](int y) { return !valid ? 0 : (y < 0 ? -y : (y >= g.ny ? 2 * (g.ny - 1) - y : y)); };
  const int rA = mir(y0 - 1), rB = mir(y0), rC = mir(y0 + 1), rD = mir(y0 + 2);
  const int tA = rA - ytile, tD = rD - ytile;  // tile rows that publish the new values of rows A and D (shared tiles)
  // PRIVATE: only the node mirrors at the two y ends point back into the warp's own pair
  const bool has_A = PRIVATE ? (relaxB && relaxC && y0 == 0) : (tA >= 0 && tA < 2 * WP);
  const bool has_D = PRIVATE ? (relaxB && relaxC && y0 + 2 == g.ny) : (tD >= 0 && tD < 2 * WP);
  const bool c_first = PRIVATE && (y0 & 1);  // the even row of the pair is relaxed first (in a shifted grid that is row C)
  const int z0 = blockIdx.z * zc, z1 = min(z0 + zc, g.nz);
  const int xo = p.xl;
  const int oA = rA * g.pitch + xo, oB = rB * g.pitch + xo, oC = rC * g.pitch + xo, oD = rD * g.pitch + xo;
  URows4 um, uc, up;
  auto zero_rows = [](URows4& R) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int k = 0; k < 6; ++k) R.r[i].v[k] = 0.f;
  };
  auto load_rows = [&](int zplane, URows4& R) {
    const int b = zplane * (int)g.plane;
    const Raw6<float> a0 = issue6(u, b + oA, p), a1 = issue6(u, b + oB, p), a2 = issue6(u, b + oC, p), a3 = issue6(u, b + oD, p);
    R.r[0] = finish6<float, float>(a0, p); R.r[1] = finish6<float, float>(a1, p);
    R.r[2] = finish6<float, float>(a2, p); R.r[3] = finish6<float, float>(a3, p);
    if (p.xb) {
#pragma unroll
      for (int i = 0; i < 4; ++i) mirror_x(R.r[i], p.xt, p.jl);
    }
  };
  if (uzero) { zero_rows(um); zero_rows(uc); }
  else { load_rows(zmirror_lo(g, z0), um); load_rows(z0, uc); }
  for (int z = z0; z < z1; ++z) {
    const int cb = z & 1, pb = cb ^ 1;
    const int zb = z * (int)g.plane;
    // ---- loads of this plane step ----
    const CoefRaw cB = issue_coef(coef, g, p, rB, z), cC = issue_coef(coef, g, p, rC, z);
    const Raw4<float> rfB = issue4(f, zb + oB), rfC = issue4(f, zb + oC);
    if (uzero) zero_rows(up);
    else load_rows(zmirror_hi(g, z), up);
    const V4<float> fB = finish4<float>(rfB), fC = finish4<float>(rfC);
    // plane z-1 was relaxed one step ago: the neighbouring warps' rows come from the tile's row buffer
    if (z > z0) {
      if constexpr (PRIVATE) {  // the mirrored rows are the warp's own: their new values of plane z-1 are in registers
        if (has_A) um.r[0] = um.r[2];
        if (has_D) um.r[3] = um.r[1];
      } else {
        if (has_A) { const float4 q = sh[pb][tA][lane]; um.r[0].v[1] = q.x; um.r[0].v[2] = q.y; um.r[0].v[3] = q.z; um.r[0].v[4] = q.w; }
        if (has_D) { const float4 q = sh[pb][tD][lane]; um.r[3].v[1] = q.x; um.r[3].v[2] = q.y; um.r[3].v[3] = q.z; um.r[3].v[4] = q.w; }
      }
    }
    if (z == g.nz - 1 && g.zhi_phys) up = um;  // mirrored plane z+1 == plane z-1, already relaxed
    // relax one of the warp's two rows: even x, then odd x (x-neighbours through shuffles)
    auto relax = [&](int row, const CoefRaw& c, const V4<float>& fv, int orow) {
      const float n0 = fv.v[0] * coef_inv(c, 0) - offdiag16_rows(c, um, uc, up, row, 0);
      const float n2 = fv.v[2] * coef_inv(c, 2) - offdiag16_rows(c, um, uc, up, row, 2);
      uc.r[row].v[1] = n0; uc.r[row].v[3] = n2;
      {
        const float r = __shfl_down_sync(FULL, n0, 1);
        if (lane < 31) uc.r[row].v[5] = r;
        if (p.xb) mirror_x(uc.r[row], p.xt, p.jl);
      }
      const float n1 = fv.v[1] * coef_inv(c, 1) - offdiag16_rows(c, um, uc, up, row, 1);
      const float n3 = fv.v[3] * coef_inv(c, 3) - offdiag16_rows(c, um, uc, up, row, 3);
      uc.r[row].v[2] = n1; uc.r[row].v[4] = n3;
      {
        const float l = __shfl_up_sync(FULL, n3, 1);
        if (lane > 0) uc.r[row].v[0] = l;
        if (p.xb) mirror_x(uc.r[row], p.xt, p.jl);
      }
      if constexpr (!PRIVATE) sh[cb][2 * w + row - 1][lane] = make_float4(n0, n1, n2, n3);
      const float res[4] = {n0, n1, n2, n3};
      if (p.xt < g.nx) { store4<float>(out, zb + orow, p.xt, g.nx, res); store_ghosts<float>(g, z, orow, p.xt, res); }
    };
    // phase 1: even rows (rows y0-1 and y0+1 of this plane still hold the previous sweep)
    if (c_first) {  // shifted warp-private pair: row C is the even one
      if (relaxC) relax(2, cC, fC, oC);
      if (relaxB) relax(1, cB, fB, oB);
    } else if (relaxB) relax(1, cB, fB, oB);
    if constexpr (!PRIVATE) __syncthreads();
    // phase 2: odd rows; row y0+2 (an even row) was relaxed in phase 1 by the next warp
    if (relaxC && !c_first) {
      if constexpr (PRIVATE) {
        if (has_D) uc.r[3] = uc.r[1];  // top row pair: the mirrored row y0+2 is row y0, just relaxed
      } else if (has_D) {
        const float* sd = reinterpret_cast<const float*>(&sh[cb][tD][0]);
        const float4 q = reinterpret_cast<const float4*>(sd)[lane];
        uc.r[3].v[1] = q.x; uc.r[3].v[2] = q.y; uc.r[3].v[3] = q.z; uc.r[3].v[4] = q.w;
        if (lane > 0) uc.r[3].v[0] = sd[lane * 4 - 1];
        if (lane < 31) uc.r[3].v[5] = sd[lane * 4 + 4];
        if (p.xb) mirror_x(uc.r[3], p.xt, p.jl);
      }
      relax(2, cC, fC, oC);
    }
    if constexpr (!PRIVATE) __syncthreads();
    um = uc; uc = up;
  }
}

// ------------------------------------------------------------------------------------------
// Temporal blocking: S Gauss-Seidel sweeps of a V-cycle leg in ONE pass over the volume (k_coef_gs_tb).
//
// A sweep of k_coef_gs2 moves exactly the bytes it has to (32 B per voxel: u, f, packed rows, u') at ~0.87 of the copy
// bandwidth, so the nu sweeps of a leg can only get faster by not moving u, f and the rows nu times.  Here the CTA keeps a
// ring of 2S+2 planes of its tile of u in SHARED MEMORY (tile = 128 x 2*WP rows, plus a frozen one-voxel halo), fed one plane
// per step by cp.async, and runs the S sweeps as a software pipeline along z: at step t sweep s relaxes plane t - 2s.  The lag
// of two planes makes the planes written in a step (t, t-2, ...) disjoint from the planes read from other sweeps (t+-1, t-2+-1,
// ...), so all sweeps share the two phases of a step (even rows, odd rows -- the in-plane colouring of k_coef_gs2) and the CTA
// synchronises twice per step however many sweeps are fused.  Inside the tile the result is exactly S sequential sweeps in the
// documented order; values OUTSIDE the tile (the halo rows / columns, the planes below and above the z chunk, the ghost planes
// of a z-slab) stay those of the array the pass started from for all S sweeps ("frozen halo").  That is a weaker smoother at
// the tile faces than S separate sweeps, which see the neighbours' previous sweep; the host therefore SHIFTS the tile grid by
// half a tile in y and z between consecutive passes, so the faces of one leg lie in the interior of the next leg's tiles
// (tools/gs_order_experiment.py and tests: same cycle counts as separate sweeps on the reference's volume).  All orderings share
// the fixed point A u = f; Gauss-Seidel parity is stated on the converged image.
// Packed rows and f of the S planes being relaxed are re-read per sweep (L1 / L2 hits: a plane is reused two and four steps
// after its first touch), u is read from HBM once and written once per pass: ~32 B per voxel for S sweeps.
// grid = (ceil(nx/128), ceil((ny+oy)/(2 WP)), ceil((nz+oz)/zc)), block = (32, WP), dynamic shared memory = tb_smem_bytes(S, WP).
// ------------------------------------------------------------------------------------------
constexpr int TB_ROWF = 136;  // floats per tile row: [3] = x0-1, [4..131] = the 128 voxels, [132] = x0+128
__host__ __device__ constexpr int tb_planes(int S) { return 2 * S + 2; }
__host__ __device__ constexpr size_t tb_smem_bytes(int S, int WP) { return (size_t)tb_planes(S) * (2 * WP + 2) * TB_ROWF * sizeof(float); }
// STAGE (single sweeps only): the packed rows and f of the planes t, t+1, t+2 are staged in shared memory as well, by cp.async two
// planes ahead of their use -- loads in flight without a register destination, which is what the register-bound k_coef_gs2 lacks.
// Measured on B200 at 512^3 (MADGPU_GS_TB_SINGLE=2|3): 1.37 ms per sweep against 0.77 -- every operand now crosses shared memory
// twice and the sweep is bound by its load / store unit.  Kept as a tested A/B hook.
constexpr int TB_STAGE_DEPTH = 3;
__host__ __device__ constexpr size_t tb_stage_bytes(int WP)  // + one more plane of the u ring
{
  return (size_t)TB_STAGE_DEPTH * 2 * WP * (32 * COEF_WORDS * 16 + 128 * sizeof(float)) + (size_t)(2 * WP + 2) * TB_ROWF * sizeof(float);
}

#ifndef MAD_HOST_EMULATION
__device__ __forceinline__ void cp_async16(float* dst_smem, const float* src)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async16u(uint4* dst_smem, const uint4* src)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
#else  // the CPU test build copies at issue time
inline void cp_async16(float* dst, const float* src) { for (int i = 0; i < 4; ++i) dst[i] = src[i]; }
inline void cp_async4(float* dst, const float* src) { dst[0] = src[0]; }
inline void cp_async_wait_all() {}
inline void cp_async_commit() {}
template <int N>
inline void cp_async_wait_group() {}
inline void cp_async16u(uint4* dst, const uint4* src) { *dst = *src; }
#endif

// six consecutive values x-1 .. x+4 of a tile row (the thread's four voxels and their x-neighbours)
__device__ __forceinline__ V6<float> tb_row6(const float* __restrict__ row, int lane)
{
  // one 16-byte read per lane (conflict-free); the x-neighbours come from the adjacent lanes by shuffle -- scalar reads at a
  // stride of four floats would be 4-way bank conflicts -- and from the frozen halo slots at the two warp ends
  const float* q = row + 4 + 4 * lane;
  const float4 c = *reinterpret_cast<const float4*>(q);
  float l = __shfl_up_sync(FULL, c.w, 1), r = __shfl_down_sync(FULL, c.x, 1);
  if (lane == 0) l = q[-1];
  if (lane == 31) r = q[4];
  V6<float> w;
  w.v[0] = l; w.v[1] = c.x; w.v[2] = c.y; w.v[3] = c.z; w.v[4] = c.w; w.v[5] = r;
  return w;
}
__device__ __forceinline__ V4<float> tb_row4(const float* __restrict__ row, int lane)
{
  const float4 c = *reinterpret_cast<const float4*>(row + 4 + 4 * lane);
  V4<float> w;
  w.v[0] = c.x; w.v[1] = c.y; w.v[2] = c.z; w.v[3] = c.w;
  return w;
}

// neighbourhood of a row being relaxed: plane below (m), its own plane (c), plane above (p); rows y-1 (A), y (B), y+1 (C)
struct TbNb {
  V6<float> cA, cB, cC, mB, pB;
  V4<float> mA, mC, pA, pC;
};
__device__ __forceinline__ float offdiag16_tb(const CoefRaw& c, const TbNb& n, int j)
{
  float s = coef_at(c, 1, j) * n.cB.v[j + 2] + coef_at(c, 2, j) * n.cB.v[j] + coef_at(c, 3, j) * n.cC.v[j + 1] + coef_at(c, 4, j) * n.cA.v[j + 1];
  s += coef_at(c, 7, j) * ((n.cC.v[j + 2] - n.cA.v[j + 2]) - (n.cC.v[j] - n.cA.v[j]));
  s += coef_at(c, 5, j) * n.pB.v[j + 1] + coef_at(c, 6, j) * n.mB.v[j + 1];
  s += coef_at(c, 8, j) * ((n.pB.v[j + 2] - n.mB.v[j + 2]) - (n.pB.v[j] - n.mB.v[j]));
  s += coef_at(c, 9, j) * ((n.pC.v[j] - n.mC.v[j]) - (n.pA.v[j] - n.mA.v[j]));
  return s;
}

template <int S, int WP, int MINB, bool STAGE = false>
__global__ void __launch_bounds__(32 * WP, MINB) k_coef_gs_tb(Geom g, const uint4* __restrict__ coef, const float* __restrict__ u,
                                                                const float* __restrict__ f, float* __restrict__ out, int zc, int oy, int oz,
                                                                int pfd, int uzero)
{
  static_assert(!STAGE || S == 1, "staged operands: single sweeps only");
  constexpr int TY = 2 * WP, NP = tb_planes(S) + (STAGE ? 1 : 0), ROWS = TY + 2;  // STAGE: u runs one plane further ahead
  MAD_DYNAMIC_SHARED(float, smem);  // [NP][ROWS][TB_ROWF], then (STAGE) [DEPTH][TY][32][COEF_WORDS] uint4 and [DEPTH][TY][128] float
  uint4* const scoef = reinterpret_cast<uint4*>(smem + (size_t)NP * ROWS * TB_ROWF);
  float* const sf = reinterpret_cast<float*>(scoef + (size_t)TB_STAGE_DEPTH * TY * 32 * COEF_WORDS);
  const int lane = threadIdx.x, w = threadIdx.y;
  Pos p;
  p.lane = lane;
  p.xt = blockIdx.x * TX + lane * 4;
  p.xl = p.xt < g.nx ? p.xt : 0;
  p.jl = g.nx - 1 - p.xt;
  p.xb = p.xt == 0 || (p.jl >= 0 && p.jl < 4);
  const int x0 = blockIdx.x * TX;
  const int y0 = (int)blockIdx.y * TY - oy;              // first image row of the tile (may be negative for a shifted grid)
  const int ya = y0 + 2 * w;                             // the warp's even row; ny is even, so ya + 1 exists whenever ya does
  const bool valid = ya >= 0 && ya < g.ny;
  const int ra = 2 * w + 1, rb = ra + 1;                 // tile rows of ya, ya + 1 (tile row 0 = image row y0 - 1)
  const int rm_a = ya == 0 ? ra + 1 : ra - 1;            // node mirror at the y ends: row -1 is row 1, row ny is row ny - 2
  const int rp_b = ya + 1 == g.ny - 1 ? rb - 1 : rb + 1;
  const int zlo = (int)blockIdx.z * zc - oz;
  const int z0 = max(zlo, 0), z1 = min(zlo + zc, g.nz);
  if (z0 >= z1) return;                                  // whole CTA
  const int zmin_load = max(z0 - 1, g.zlo_phys ? 0 : -1), zmax_load = min(z1, g.zhi_phys ? g.nz - 1 : g.nz);
  auto slot_of = [&](int pz) { return (pz - (z0 - 1)) % NP; };
  auto tile_row = [&](int slot, int r) { return smem + ((size_t)slot * ROWS + r) * TB_ROWF; };
  // one plane of the tile (its rows inside the image, with the x halo) -> shared memory, asynchronously
  auto load_row = [&](int pz, int slot, int r, int yy) {
    if (yy < 0 || yy >= g.ny) return;
    const float* src = u + (long long)pz * g.plane + (long long)yy * g.pitch;
    float* dst = tile_row(slot, r);
    if (p.xt < g.nx) cp_async16(dst + 4 + 4 * lane, src + p.xt);
    if (lane == 0 && x0 > 0) cp_async4(dst + 3, src + x0 - 1);
    if (lane == 31 && x0 + TX < g.nx) cp_async4(dst + 4 + TX, src + x0 + TX);
  };
  auto load_plane = [&](int pz) {
    const int slot = slot_of(pz);
    load_row(pz, slot, ra, ya);
    load_row(pz, slot, rb, ya + 1);
    if (w == 0) load_row(pz, slot, 0, y0 - 1);
    if (w == WP - 1) load_row(pz, slot, TY + 1, y0 + TY);
  };
  // STAGE: packed rows and f of the warp's two rows of plane pz -> slot pz % DEPTH (every thread copies what it will read itself)
  auto stage_slot = [&](int pz) { return (pz - z0) % TB_STAGE_DEPTH; };
  auto stage_plane = [&](int pz) {
    if (!STAGE || !valid || pz >= z1 || p.xt >= g.nx) return;
    const int sl = stage_slot(pz);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint4* src = coef + coef_quad(g, p.xt >> 2, ya + h, pz) * COEF_WORDS;
      uint4* dst = scoef + (((size_t)sl * TY + 2 * w + h) * 32 + lane) * COEF_WORDS;
#pragma unroll
      for (int i = 0; i < COEF_WORDS; ++i) cp_async16u(dst + i, src + i);
      cp_async16(sf + ((size_t)sl * TY + 2 * w + h) * 128 + 4 * lane, f + (long long)pz * g.plane + (long long)(ya + h) * g.pitch + p.xt);
    }
  };
  if (uzero) {  // the iterate is identically zero (first sweeps of a leg): nothing is loaded, the ring starts as zeros
    for (int i = (w * 32 + lane) * 4; i < NP * ROWS * TB_ROWF; i += 32 * WP * 4)
      *reinterpret_cast<float4*>(smem + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    for (int pz = zmin_load; pz <= min(z0 + 1, zmax_load); ++pz) load_plane(pz);
  }
  if (STAGE) {
    // one commit group per plane step: the group of step t holds the operands of plane t + 2 and u of plane t + 3; at the end of a
    // step everything but its own group has landed -- the operands of plane t + 1 and u up to plane t + 2, what step t + 1 reads
    stage_plane(z0);
    cp_async_commit();
    if (!uzero && z0 + 2 <= zmax_load) load_plane(z0 + 2);
    stage_plane(z0 + 1);
    cp_async_commit();
    cp_async_wait_group<1>();
  } else {
    cp_async_wait_all();
  }
  __syncthreads();
  // L2 prefetch of the first-touch stream (packed rows and f of the plane sweep 0 reaches in pfd steps)
  const char* pfa = coef_prefetch_base(g, u, f, coef, lane, valid ? ya : 0);
  const long long pf_stride = lane < 8 ? g.plane * 4ll : (long long)g.ny * (g.pitch >> 2) * COEF_WORDS * 16ll;
  const long long pf_row = lane < 8 ? g.pitch * 4ll : (long long)(g.pitch >> 2) * COEF_WORDS * 16ll;

  // relax the four voxels of the thread in tile row r (image row y) of plane pc with sweep s
  auto relax = [&](int s, int pc, int r, int y, int rm, int rp) {
    const int zm = (pc == 0 && g.zlo_phys) ? pc + 1 : pc - 1, zp = (pc == g.nz - 1 && g.zhi_phys) ? pc - 1 : pc + 1;
    const int o = pc * (int)g.plane + y * g.pitch + p.xl;
    CoefRaw c;
    V4<float> fv;
    if (STAGE) {
      const int sl = stage_slot(pc), tr = r - 1;  // tile row 1 = the tile's first image row
      const uint4* q = scoef + (((size_t)sl * TY + tr) * 32 + lane) * COEF_WORDS;
#pragma unroll
      for (int i = 0; i < COEF_WORDS; ++i) c.w[i] = q[i];
      const float4 t4 = *reinterpret_cast<const float4*>(sf + ((size_t)sl * TY + tr) * 128 + 4 * lane);
      fv.v[0] = t4.x; fv.v[1] = t4.y; fv.v[2] = t4.z; fv.v[3] = t4.w;
    } else {
      c = issue_coef(coef, g, p, y, pc);
      fv = finish4<float>(issue4(f, o));
    }
    const int sm = slot_of(zm), sc = slot_of(pc), sp = slot_of(zp);
    TbNb n;
    n.cA = tb_row6(tile_row(sc, rm), lane); n.cB = tb_row6(tile_row(sc, r), lane); n.cC = tb_row6(tile_row(sc, rp), lane);
    n.mB = tb_row6(tile_row(sm, r), lane); n.pB = tb_row6(tile_row(sp, r), lane);
    n.mA = tb_row4(tile_row(sm, rm), lane); n.mC = tb_row4(tile_row(sm, rp), lane);
    n.pA = tb_row4(tile_row(sp, rm), lane); n.pC = tb_row4(tile_row(sp, rp), lane);
    if (p.xb) {
      mirror_x(n.cA, p.xt, p.jl); mirror_x(n.cB, p.xt, p.jl); mirror_x(n.cC, p.xt, p.jl);
      mirror_x(n.mB, p.xt, p.jl); mirror_x(n.pB, p.xt, p.jl);
    }
    // even x (slots 0, 2), then odd x (1, 3); the right neighbour's new slot 0 arrives by shuffle (lane 31 keeps the frozen halo)
    const float n0 = fv.v[0] * coef_inv(c, 0) - offdiag16_tb(c, n, 0);
    const float n2 = fv.v[2] * coef_inv(c, 2) - offdiag16_tb(c, n, 2);
    n.cB.v[1] = n0; n.cB.v[3] = n2;
    {
      const float rr = __shfl_down_sync(FULL, n0, 1);
      if (lane < 31) n.cB.v[5] = rr;
      if (p.xb) mirror_x(n.cB, p.xt, p.jl);
    }
    const float n1 = fv.v[1] * coef_inv(c, 1) - offdiag16_tb(c, n, 1);
    const float n3 = fv.v[3] * coef_inv(c, 3) - offdiag16_tb(c, n, 3);
    __syncwarp();  // every lane has read its neighbours' old values of this row
    *reinterpret_cast<float4*>(tile_row(sc, r) + 4 + 4 * lane) = make_float4(n0, n1, n2, n3);
    if (s == S - 1 && p.xt < g.nx) {
      const float res[4] = {n0, n1, n2, n3};
      store4<float>(out, o, p.xt, g.nx, res);
      store_ghosts<float>(g, pc, y * g.pitch + p.xl, p.xt, res);
    }
  };

  const int t_end = z1 - 1 + 2 * (S - 1);
  for (int t = z0; t <= t_end; ++t) {
    if (!uzero && t + 2 + (STAGE ? 1 : 0) <= zmax_load) load_plane(t + 2 + (STAGE ? 1 : 0));
    if (STAGE) { stage_plane(t + 2); cp_async_commit(); }
    if (!STAGE && pfd > 0 && valid && pfa && t + pfd < z1) {
      const char* q = pfa + pf_stride * (t + pfd);
      if (lane >= 4) { mad_prefetch_l2(q); mad_prefetch_l2(q + pf_row); }
    }
    // phase 1: the even rows of every sweep's plane (their y-neighbours are odd rows: previous values)
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int pc = t - 2 * s;
      if (valid && pc >= z0 && pc < z1) relax(s, pc, ra, ya, rm_a, rb);
    }
    __syncthreads();
    // phase 2: the odd rows (their y-neighbours were relaxed in phase 1, by this warp and the next one)
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int pc = t - 2 * s;
      if (valid && pc >= z0 && pc < z1) relax(s, pc, rb, ya + 1, ra, rp_b);
    }
    if (STAGE) cp_async_wait_group<1>();  // everything but this step's group: the operands and the u plane of step t + 1 are in
    else cp_async_wait_all();
    __syncthreads();
  }
}

// r = f - A u from the packed rows (the residual that feeds the restriction INSIDE a Gauss-Seidel V-cycle: it is taken
// with the same fp16-rounded operator the sweeps relax, so the inner cycle is a consistent multigrid cycle for that
// operator; the outer defect and the stop test keep the exact rows).  Same marching structure, no phases.
template <int WY, int MINB>
__global__ void __launch_bounds__(32 * WY, MINB) k_coef_residual(Geom g, const uint4* __restrict__ coef, const float* __restrict__ u,
                                                                   const float* __restrict__ f, float* __restrict__ out, int zc, int pfd)
{
  const Pos p = make_pos(g);
  if (p.y >= g.ny) return;
  const int z0 = blockIdx.z * zc, z1 = min(z0 + zc, g.nz);
  const int rowo = p.y * g.pitch + p.xl;
  const char* pf = coef_prefetch_base(g, u, f, coef, p.lane, p.y);
  const long long pf_stride = p.lane < 8 ? g.plane * 4ll : (long long)g.ny * (g.pitch >> 2) * COEF_WORDS * 16ll;
  const int zpf_end = min(z1 + 1, g.nz);
  UPlane<float> um, uc, up;
  {
    const URaw<float> r0 = issue_u(u, zmirror_lo(g, z0) * (int)g.plane + rowo, p), r1 = issue_u(u, z0 * (int)g.plane + rowo, p);
    finish_u<float, float>(r0, p, um);
    finish_u<float, float>(r1, p, uc);
  }
  for (int z = z0; z < z1; ++z) {
    const int oc = z * (int)g.plane + rowo;
    if (pfd > 0 && z + pfd < zpf_end && pf) mad_prefetch_l2((pf + pf_stride * (z + pfd)));
    const URaw<float> ru = issue_u(u, zmirror_hi(g, z) * (int)g.plane + rowo, p);
    const CoefRaw c = issue_coef(coef, g, p, p.y, z);
    const Raw4<float> rf = issue4(f, oc);
    finish_u<float, float>(ru, p, up);
    const V4<float> fv = finish4<float>(rf);
    float res[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float inv = coef_inv(c, j);
      // rows are stored divided by diag: A u = (u + sum_k c_k u_k) / inv
      res[j] = __fdividef(fv.v[j] * inv - uc.r[1].v[j + 1] - offdiag16(c, um, uc, up, j), inv);
    }
    if (p.xt < g.nx) { store4<float>(out, oc, p.xt, g.nx, res); store_ghosts<float>(g, z, rowo, p.xt, res); }
    um = uc; uc = up;
  }
}

// ------------------------------------------------------------------------------------------
// Inter-grid transfers, 4 fine voxels per thread along x.
// ------------------------------------------------------------------------------------------

// fine += P coarse (ADD) / fine = P coarse: gather form of Interpolation (mad/itkInterGridOperators.hxx:45-172,
// tables .h:101-113) fused with the correction add of the V-cycle (…Filter.hxx:424-435).
// The four fine voxels 4t..4t+3 of a thread interpolate from the coarse voxels 2t-1..2t+2 of up to four
// coarse rows: 3 loads per coarse row (one 8-byte pair + two scalars), one 16-byte load/store of the fine row.
constexpr int PROLONG_ZB = 8;  // fine planes per CTA (a CTA per plane would be ~10^5 very short CTAs on a 512^3 level)

// interpolated values of the thread's four fine voxels of row (y, z)
__device__ __forceinline__ void prolong_row(const Geom& gc, const Geom& gf, const Transfer& t, const float* __restrict__ coarse, int xt, int y, int z,
                                            float acc[4])
{
  int y0, y1, z0, z1;
  float wy0, wy1, wz0, wz1;
  prolong_taps(y, gf.ny, gc.ny, t.cent[1], y0, y1, wy0, wy1);
  prolong_taps(z, gf.nz, gc.nz, t.cent[2], z0, z1, wz0, wz1, gf.zlo_phys != 0, gf.zhi_phys != 0);
  if (t.cent[0] == 1 && xt > 0 && xt + 4 < gf.nx) {
    // cell-centred axis away from the two ends (every thread but two per row on power-of-two volumes): fixed 3/4, 1/4 taps,
    // fine 4t..4t+3 from coarse 2t-1..2t+2
    const int cb = (xt >> 1) - 1;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int yy = (r & 1) ? y1 : y0, zz = (r & 2) ? z1 : z0;
      const float wr = ((r & 1) ? wy1 : wy0) * ((r & 2) ? wz1 : wz0);
      const float* row = coarse + (long long)zz * gc.plane + (long long)yy * gc.pitch + cb;
      const float c0 = __ldg(row), c3 = __ldg(row + 3);
      const float2 c12 = __ldg(reinterpret_cast<const float2*>(row + 1));
      a0 += wr * (.75f * c12.x + .25f * c0);
      a1 += wr * (.75f * c12.x + .25f * c12.y);
      a2 += wr * (.75f * c12.y + .25f * c12.x);
      a3 += wr * (.75f * c12.y + .25f * c3);
    }
    acc[0] = a0; acc[1] = a1; acc[2] = a2; acc[3] = a3;
    return;
  }
  // x taps of the four fine voxels as weights on the coarse voxels cb..cb+3, cb = 2t-1
  const int cb = (xt >> 1) - 1;
  float W[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int i0, i1;
    float w0, w1;
    prolong_taps(xt + j, gf.nx, gc.nx, t.cent[0], i0, i1, w0, w1);
    if (xt + j >= gf.nx) { w0 = 0.f; w1 = 0.f; i0 = i1 = cb + 1; }
#pragma unroll
    for (int k = 0; k < 4; ++k) W[j][k] = (i0 - cb == k ? w0 : 0.f) + (i1 - cb == k ? w1 : 0.f);
  }
  const int xa = max(cb, 0), xd = min(cb + 3, gc.nx - 1);            // clamped (their weights are zero when clamped)
  const int xb = cb + 1, xc = min(cb + 2, gc.nx - 1);               // cb+1 = 2t is always a valid, 8-byte aligned voxel
  const bool pair = xb + 1 < gc.nx;
  acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int yy = (r & 1) ? y1 : y0, zz = (r & 2) ? z1 : z0;
    const float wr = ((r & 1) ? wy1 : wy0) * ((r & 2) ? wz1 : wz0);
    // no early-out on wr == 0: the tap rows are always valid, and branch-free code lets all twelve loads issue together
    const float* row = coarse + (long long)zz * gc.plane + (long long)yy * gc.pitch;
    float c[4];
    c[0] = __ldg(row + xa);
    if (pair) { const float2 q = __ldg(reinterpret_cast<const float2*>(row + xb)); c[1] = q.x; c[2] = q.y; }
    else { c[1] = __ldg(row + xb); c[2] = __ldg(row + xc); }
    c[3] = __ldg(row + xd);
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] += wr * (W[j][0] * c[0] + W[j][1] * c[1] + W[j][2] * c[2] + W[j][3] * c[3]);
  }
}

// grid = (ceil(nxf/128), ceil(nyf/WY), ceil(nzf/PROLONG_ZB)), block = (32, WY)
template <bool ADD, int WY>
__global__ void __launch_bounds__(32 * WY) k_fast_prolong(Geom gc, Geom gf, Transfer t, const float* __restrict__ coarse, float* __restrict__ fine)
{
  const int xt = blockIdx.x * TX + threadIdx.x * 4;
  const int y = blockIdx.y * WY + threadIdx.y;
  if (xt >= gf.nx || y >= gf.ny) return;
  const int z0 = blockIdx.z * PROLONG_ZB, z1 = min(z0 + PROLONG_ZB, gf.nz);
  for (int z = z0; z < z1; ++z) {
    const int o = z * (int)gf.plane + y * gf.pitch + xt;
    float acc[4];
    prolong_row(gc, gf, t, coarse, xt, y, z, acc);
    if (ADD) {
      const float4 q = *reinterpret_cast<const float4*>(fine + o);
      acc[0] += q.x; acc[1] += q.y; acc[2] += q.z; acc[3] += q.w;
    }
    store4<float>(fine, o, xt, gf.nx, acc);
    store_ghosts<float>(gf, z, y * gf.pitch + xt, xt, acc);
  }
}

// The same operator for transfers that are cell-centred along all three axes (every level of a power-of-two volume), blocked
// 4 x 2 x 2: a thread produces the fine voxels 4t..4t+3 of the rows 2j, 2j+1 of the planes 2k, 2k+1 from the coarse voxels
// 2t-1..2t+2 of the rows j-1..j+1 of the planes k-1..k+1.  The CTA marches along z with the x-interpolated rows of three coarse
// planes in registers, so a step loads ONE coarse plane (three 8-byte loads per thread, x-neighbours through shuffles) and emits
// two fine planes: ~10 instructions per fine voxel instead of ~95 in k_fast_prolong, whose per-voxel tap arithmetic bounded it.
// The end rules of the cell-centred tables (fine[0] = c[0], fine[n-1] = c[nc-1], mad/itkInterGridOperators.h:101-113) are the
// interior weights on an index clamped to the grid (3/4 c0 + 1/4 c0); the inner faces of a z-slab read the ghost planes.
// grid = (ceil(nxf/128), ceil(nyc/WY), ceil(nzc/zcc)), block = (32, WY); needs nxf, nyf, nzf even.
template <bool ADD, int WY>
__global__ void __launch_bounds__(32 * WY) k_fast_prolong_cell(Geom gc, Geom gf, const float* __restrict__ coarse, float* __restrict__ fine, int zcc)
{
  const int lane = threadIdx.x;
  const int xt = blockIdx.x * TX + lane * 4;        // first fine voxel of the thread
  const int j = blockIdx.y * WY + threadIdx.y;      // coarse row
  if (j >= gc.ny) return;                           // whole warp
  const int xc = min(xt >> 1, (gc.nx - 1) & ~1);    // coarse pair (xc, xc+1) loaded by this lane: even (8-byte aligned), clamped into the row
  const int jm = max(j - 1, 0), jp = min(j + 1, gc.ny - 1);
  const int k0 = blockIdx.z * zcc, k1 = min(k0 + zcc, gc.nz);
  const bool act = xt < gf.nx;
  const int xl = max((xt >> 1) - 1, 0), xr = min((xt >> 1) + 2, gc.nx - 1);  // the two x-neighbours at the warp ends (clamped = end rule)
  // x-interpolated coarse row: the four fine x values of the thread
  auto xrow = [&](int k, int jj, float o[4]) {
    const float* row = coarse + (long long)k * gc.plane + (long long)jj * gc.pitch;
    float2 c = __ldg(reinterpret_cast<const float2*>(row + xc));  // (rows are padded to 32 elements: xc + 1 is always readable)
    if (xc + 1 >= gc.nx) c.y = c.x;                                // odd coarse row length: the last pair is (c[nc-1], end rule)
    float l = __shfl_up_sync(FULL, c.y, 1), r = __shfl_down_sync(FULL, c.x, 1);
    if (lane == 0 || xt == 0) l = __ldg(row + xl);
    if (lane == 31 || (xt >> 1) + 2 >= gc.nx) r = __ldg(row + xr);
    o[0] = .75f * c.x + .25f * l; o[1] = .75f * c.x + .25f * c.y; o[2] = .75f * c.y + .25f * c.x; o[3] = .75f * c.y + .25f * r;
  };
  // y-interpolated pair of fine rows (2j, 2j+1) of one coarse plane: [0] = row 2j, [1] = row 2j+1
  struct P2 { float v[2][4]; };
  auto yplane = [&](int k, P2& o) {
    float a[4], b[4], c[4];
    xrow(k, jm, a); xrow(k, j, b); xrow(k, jp, c);
#pragma unroll
    for (int i = 0; i < 4; ++i) { o.v[0][i] = .75f * b[i] + .25f * a[i]; o.v[1][i] = .75f * b[i] + .25f * c[i]; }
  };
  const int klo = gf.zlo_phys ? 0 : -1, khi = gf.zhi_phys ? gc.nz - 1 : gc.nz;  // ghost planes of a z-slab are valid
  P2 pm, pc, pp;
  yplane(max(k0 - 1, klo), pm);
  yplane(k0, pc);
  for (int k = k0; k < k1; ++k) {
    yplane(min(k + 1, khi), pp);
#pragma unroll
    for (int h = 0; h < 2; ++h) {      // fine plane 2k + h
      const P2& q = h ? pp : pm;
      const int z = 2 * k + h;
#pragma unroll
      for (int r = 0; r < 2; ++r) {    // fine row 2j + r
        const int y = 2 * j + r;
        const int rowoff = y * gf.pitch + xt;
        const int o = z * (int)gf.plane + rowoff;
        float acc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = .75f * pc.v[r][i] + .25f * q.v[r][i];
        if (act) {
          if (ADD) {
            const float4 t = *reinterpret_cast<const float4*>(fine + o);
            acc[0] += t.x; acc[1] += t.y; acc[2] += t.z; acc[3] += t.w;
          }
          store4<float>(fine, o, xt, gf.nx, acc);
          store_ghosts<float>(gf, z, rowoff, xt, acc);
        }
      }
    }
    pm = pc; pc = pp;
  }
}

// coarse = R fine: full weighting (mad/itkInterGridOperators.hxx:175-304, tables .h:115-127).  A thread reads
// the fine voxels 4t-1..4t+4 of each contributing fine row (one 16-byte load + two shuffles) and produces the
// coarse voxels 2t, 2t+1 (one 8-byte store); a warp covers 128 fine = 64 coarse voxels of one coarse row.
// grid = (ceil(nxf/128), ceil(nyc/WY), nzc), block = (32, WY).
template <int WY>
__global__ void __launch_bounds__(32 * WY) k_fast_restrict(Geom gf, Geom gc, Transfer t, const float* __restrict__ fine, float* __restrict__ coarse)
{
  Pos p;
  p.lane = threadIdx.x;
  p.xt = blockIdx.x * TX + p.lane * 4;
  p.xl = p.xt < gf.nx ? p.xt : 0;
  p.edge = p.lane == 0 || p.lane == 31;
  p.dh = (p.lane == 0 ? max(p.xt - 1, 0) : min(p.xt + 4, gf.nx - 1)) - p.xl;
  const int yc = blockIdx.y * WY + threadIdx.y;
  const int zc = blockIdx.z;
  if (yc >= gc.ny) return;  // whole warp
  float wy[4], wz[4], wxa[4], wxb[4];
  restrict_taps(yc, gc.ny, t.cent[1], wy);
  restrict_taps(zc, gc.nz, t.cent[2], wz, gc.zlo_phys != 0, gc.zhi_phys != 0);
  const int xc0 = p.xt >> 1;  // coarse voxels xc0, xc0+1
  restrict_taps(min(xc0, gc.nx - 1), gc.nx, t.cent[0], wxa);
  restrict_taps(min(xc0 + 1, gc.nx - 1), gc.nx, t.cent[0], wxb);
  if (t.cent[1] == 1 && t.cent[2] == 1 && yc > 0 && yc < gc.ny - 1 && (zc > 0 || !gf.zlo_phys) && (zc < gc.nz - 1 || !gf.zhi_phys)) {
    // away from the y/z ends of a cell-centred transfer (warp-uniform test): fixed (1/8, 3/8, 3/8, 1/8) taps along y and z, no
    // clamping; along x the per-thread taps (zero on voxels outside the row, whose stand-in values are finite)
    const float wt[4] = {.125f, .375f, .375f, .125f};
    Raw6<float> rw[4][4];
#pragma unroll
    for (int kz = 0; kz < 4; ++kz)
#pragma unroll
      for (int ky = 0; ky < 4; ++ky) rw[kz][ky] = issue6(fine, (2 * zc + kz - 1) * (int)gf.plane + (2 * yc + ky - 1) * gf.pitch + p.xl, p);
    float b0 = 0.f, b1 = 0.f;
#pragma unroll
    for (int kz = 0; kz < 4; ++kz)
#pragma unroll
      for (int ky = 0; ky < 4; ++ky) {
        const V6<float> v = finish6<float, float>(rw[kz][ky], p);
        const float wr = wt[ky] * wt[kz];
        b0 += wr * (wxa[0] * v.v[0] + wxa[1] * v.v[1] + wxa[2] * v.v[2] + wxa[3] * v.v[3]);
        b1 += wr * (wxb[0] * v.v[2] + wxb[1] * v.v[3] + wxb[2] * v.v[4] + wxb[3] * v.v[5]);
      }
    const long long oo = (long long)zc * gc.plane + (long long)yc * gc.pitch + xc0;
    if (xc0 + 1 < gc.nx) *reinterpret_cast<float2*>(coarse + oo) = make_float2(b0, b1);
    else if (xc0 < gc.nx) coarse[oo] = b0;
    return;
  }
  float a0 = 0.f, a1 = 0.f;
  // All sixteen tap rows are loaded unconditionally (rows outside the image are clamped and carry weight 0):
  // branch-free, so the loads are issued back to back.
  Raw6<float> raw[4][4];
#pragma unroll
  for (int kz = 0; kz < 4; ++kz) {
    const int fz = min(max(2 * zc + kz - 1, gf.zlo_phys ? 0 : -1), gf.zhi_phys ? gf.nz - 1 : gf.nz);  // ghost planes of a z-slab are valid
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
      const int fy = min(max(2 * yc + ky - 1, 0), gf.ny - 1);
      raw[kz][ky] = issue6(fine, fz * (int)gf.plane + fy * gf.pitch + p.xl, p);
    }
  }
#pragma unroll
  for (int kz = 0; kz < 4; ++kz) {
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
      const float wr = wy[ky] * wz[kz];
      const V6<float> v = finish6<float, float>(raw[kz][ky], p);
      // voxels beyond the fine row only ever meet zero weights, but may hold anything: mask them
      float m[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) { const int fx = p.xt - 1 + k; m[k] = (fx >= 0 && fx < gf.nx) ? v.v[k] : 0.f; }
      a0 += wr * (wxa[0] * m[0] + wxa[1] * m[1] + wxa[2] * m[2] + wxa[3] * m[3]);
      a1 += wr * (wxb[0] * m[2] + wxb[1] * m[3] + wxb[2] * m[4] + wxb[3] * m[5]);
    }
  }
  const long long o = (long long)zc * gc.plane + (long long)yc * gc.pitch + xc0;
  if (xc0 + 1 < gc.nx) *reinterpret_cast<float2*>(coarse + o) = make_float2(a0, a1);
  else if (xc0 < gc.nx) coarse[o] = a0;
}

// The same operator for transfers that are cell-centred along all three axes (every level of a power-of-two volume), marching
// along z: a warp owns a coarse row, a thread its coarse voxels 2t, 2t+1; every fine plane is reduced ONCE to its xy-restricted
// row (four fine rows, one 16-byte load + two shuffles each) and enters the two coarse planes it contributes to from registers
// -- two fine planes loaded per coarse plane instead of the four k_fast_restrict re-reads through L1 / L2.
// grid = (ceil(nxf/128), ceil(nyc/WY), ceil(nzc/zcc)), block = (32, WY); needs nxf, nyf, nzf even.
template <int WY>
__global__ void __launch_bounds__(32 * WY) k_fast_restrict_cell(Geom gf, Geom gc, const float* __restrict__ fine, float* __restrict__ coarse, int zcc)
{
  Pos p;
  p.lane = threadIdx.x;
  p.xt = blockIdx.x * TX + p.lane * 4;
  p.xl = p.xt < gf.nx ? p.xt : 0;
  p.edge = p.lane == 0 || p.lane == 31;
  p.dh = (p.lane == 0 ? max(p.xt - 1, 0) : min(p.xt + 4, gf.nx - 1)) - p.xl;
  const int yc = blockIdx.y * WY + threadIdx.y;
  if (yc >= gc.ny) return;  // whole warp
  const int k0 = blockIdx.z * zcc, k1 = min(k0 + zcc, gc.nz);
  float wy[4], wxa[4], wxb[4];
  restrict_taps(yc, gc.ny, 1, wy);
  const int xc0 = p.xt >> 1;  // coarse voxels xc0, xc0 + 1
  restrict_taps(min(xc0, gc.nx - 1), gc.nx, 1, wxa);
  restrict_taps(min(xc0 + 1, gc.nx - 1), gc.nx, 1, wxb);
  int ro[4];  // the four contributing fine rows (clamped rows carry weight 0)
#pragma unroll
  for (int ky = 0; ky < 4; ++ky) ro[ky] = min(max(2 * yc + ky - 1, 0), gf.ny - 1) * gf.pitch + p.xl;
  const int zlo = gf.zlo_phys ? 0 : -1, zhi = gf.zhi_phys ? gf.nz - 1 : gf.nz;  // ghost planes of a z-slab are valid
  struct P2 { float a, b; };
  // the four rows of fine plane z (clamped into the valid planes; a clamped plane only ever meets weight 0): loads only ...
  struct PR { Raw6<float> r[4]; };
  auto issue_plane = [&](int z) {
    const int zb = min(max(z, zlo), zhi) * (int)gf.plane;
    PR q;
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) q.r[ky] = issue6(fine, zb + ro[ky], p);
    return q;
  };
  // ... and their xy-restricted row
  auto finish_plane = [&](const PR& q) {
    P2 o = {0.f, 0.f};
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
      const V6<float> v = finish6<float, float>(q.r[ky], p);
      float m[6];  // voxels beyond the fine row only ever meet zero weights, but may hold anything: mask them
#pragma unroll
      for (int i = 0; i < 6; ++i) { const int fx = p.xt - 1 + i; m[i] = (fx >= 0 && fx < gf.nx) ? v.v[i] : 0.f; }
      o.a += wy[ky] * (wxa[0] * m[0] + wxa[1] * m[1] + wxa[2] * m[2] + wxa[3] * m[3]);
      o.b += wy[ky] * (wxb[0] * m[2] + wxb[1] * m[3] + wxb[2] * m[4] + wxb[3] * m[5]);
    }
    return o;
  };
  P2 pm, p0;
  {
    const PR ra = issue_plane(2 * k0 - 1), rb = issue_plane(2 * k0);
    pm = finish_plane(ra);
    p0 = finish_plane(rb);
  }
  // the loads of the next coarse plane's two fine planes are in flight while this one is reduced
  PR r1 = issue_plane(2 * k0 + 1), r2 = issue_plane(2 * k0 + 2);
  for (int k = k0; k < k1; ++k) {
    PR n1 = r1, n2 = r2;
    if (k + 1 < k1) { n1 = issue_plane(2 * k + 3); n2 = issue_plane(2 * k + 4); }
    const P2 p1 = finish_plane(r1), p2 = finish_plane(r2);
    float wz[4];
    restrict_taps(k, gc.nz, 1, wz, gc.zlo_phys != 0, gc.zhi_phys != 0);
    const float b0 = wz[0] * pm.a + wz[1] * p0.a + wz[2] * p1.a + wz[3] * p2.a;
    const float b1 = wz[0] * pm.b + wz[1] * p0.b + wz[2] * p1.b + wz[3] * p2.b;
    const long long oo = (long long)k * gc.plane + (long long)yc * gc.pitch + xc0;
    if (xc0 + 1 < gc.nx) *reinterpret_cast<float2*>(coarse + oo) = make_float2(b0, b1);
    else if (xc0 < gc.nx) coarse[oo] = b0;
    r1 = n1; r2 = n2;
    pm = p1; p0 = p2;
  }
}

}  // namespace fast
}  // namespace mad
