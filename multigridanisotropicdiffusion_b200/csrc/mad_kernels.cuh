// mad_kernels.cuh -- device code of libmadgpu (sm_100a).
//
// Data layout in HBM (per level): every scalar field is a pitched fp32 (fp64 for the two
// level-0 outer fields) volume, x fastest, rows padded to a multiple of 32 elements (128 B for
// fp32) so that every row start is float4/TMA aligned; planes are pitch*ny elements; one ghost
// plane below z=0 and one above z=nz-1 (halo planes of the z-slab decomposition).  The diffusion
// tensor is SoA: 6 (3-D: xx,xy,xz,yy,yz,zz) or 3 (2-D: xx,xy,yy) such planes.  The operator
// A = I - dt*div(D grad) is NEVER stored: its 19 (9) coefficients are evaluated from the tensor
// planes inside every kernel (row_coeffs below), which is what keeps a sweep at 36 B/voxel.
//
// Reference routines restated here (paths relative to /root/reference/include):
//   row_coeffs / apply_row  mad/itkGridsHierarchy.hxx:298-516 (GenerateDCA) in closed form
//   k_jacobi                mad/itkMultigridWeightedJacobiSmoother.hxx:33-102
//   k_gs_color              mad/itkMultigridGaussSeidelSmoother.hxx:33-111 (multicolour ordering)
//   k_residual              mad/itkMultigridGaussSeidelSmoother.hxx:114-180 + …Filter.hxx:496-515
//   k_restrict              mad/itkInterGridOperators.hxx:175-304, tables .h:115-127
//   k_prolong               mad/itkInterGridOperators.hxx:45-172,  tables .h:101-113 (gather form)
//   k_coarse_gemv           mad/itkDirectSolver.hxx:91-147
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

// dynamic shared memory of a kernel; the CPU test build (tests/mad_host/) hands out its own block
#ifndef MAD_DYNAMIC_SHARED
#define MAD_DYNAMIC_SHARED(type, name) extern __shared__ type name[]
#endif

namespace mad {

struct Geom {
  int nx, ny, nz;      // local extents (nz = 1 in 2-D)
  int pitch;           // row stride, elements
  long long plane;     // plane stride, elements
  int zlo_phys;        // local z = 0 is the physical (Neumann) boundary
  int zhi_phys;        // local z = nz-1 is the physical boundary
  int z0;              // global z of local plane 0 (colour parity)
  // z-slab decomposition with peer access: where a kernel that PRODUCES a field also stores its first / last plane --
  // the neighbour slab's upper / lower ghost plane of the same field, mapped through CUDA IPC (NVLink peer stores);
  // null on a single GPU, at the outer slabs and on the NCCL path.  Set per launch by the host.
  void* glo;
  void* ghi;
  float wx, wy, wz;    // dt / h_d^2
  float cxy, cxz, cyz; // dt / (2 h_a h_b)
  float bxx, bxy, bxz, byy, byz, bzz;  // dt / (4 h_d h_d2)
  // the same constants in fp64, used by the level-0 outer residual (stop test)
  double dwx, dwy, dwz, dcxy, dcxz, dcyz, dbxx, dbxy, dbxz, dbyy, dbyz, dbzz;
};

// Pick the constant set that matches the arithmetic type.
template <typename T>
struct GeomConst {
  T wx, wy, wz, cxy, cxz, cyz, bxx, bxy, bxz, byy, byz, bzz;
  __host__ __device__ __forceinline__ explicit GeomConst(const Geom& g)
  {
    if (sizeof(T) == 8) {
      wx = T(g.dwx); wy = T(g.dwy); wz = T(g.dwz); cxy = T(g.dcxy); cxz = T(g.dcxz); cyz = T(g.dcyz);
      bxx = T(g.dbxx); bxy = T(g.dbxy); bxz = T(g.dbxz); byy = T(g.dbyy); byz = T(g.dbyz); bzz = T(g.dbzz);
    } else {
      wx = T(g.wx); wy = T(g.wy); wz = T(g.wz); cxy = T(g.cxy); cxz = T(g.cxz); cyz = T(g.cyz);
      bxx = T(g.bxx); bxy = T(g.bxy); bxz = T(g.bxz); byy = T(g.byy); byz = T(g.byz); bzz = T(g.bzz);
    }
  }
};

struct Tensor {
  const float* p[6];
};

// component slots
enum { XX3 = 0, XY3 = 1, XZ3 = 2, YY3 = 3, YZ3 = 4, ZZ3 = 5, XX2 = 0, XY2 = 1, YY2 = 2 };

// Row of A at one voxel, interior form: with node-mirrored reads u(-1)=u(1), u(n)=u(n-2) this is
// exactly the row GenerateDCA builds by redirecting offsets (mad/itkGridsHierarchy.hxx:362-430).
template <typename T>
struct Row {
  T diag;
  T xp, xm, yp, ym, zp, zm;  // axis neighbours: -w_d D_dd +- b_d
  T exy, exz, eyz;           // edge pairs: coefficient of (u++ - u+- - u-+ + u--) = -c_ab D_ab
};

// Difference of a tensor component along one axis as GenerateDCA takes it
// (mad/itkGridsHierarchy.hxx:451-470): central in the interior (un-normalised: D+ - D-),
// second-order one-sided on the first / last node.
template <typename T>
__host__ __device__ __forceinline__ T tensor_diff(const float* __restrict__ p, long long c, long long s, bool lo, bool hi)
{
  if (lo) return T(-3) * T(p[c]) + T(4) * T(p[c + s]) - T(p[c + 2 * s]);
  if (hi) return T(3) * T(p[c]) - T(4) * T(p[c - s]) + T(p[c - 2 * s]);
  return T(p[c + s]) - T(p[c - s]);
}

template <int DIM, typename T>
__host__ __device__ __forceinline__ void row_coeffs(const Geom& g, const Tensor& D, int x, int y, int z, Row<T>& r)
{
  const long long c = (long long)z * g.plane + (long long)y * g.pitch + x;
  const GeomConst<T> k(g);
  const bool xlo = x == 0, xhi = x == g.nx - 1, ylo = y == 0, yhi = y == g.ny - 1;
  if (DIM == 3) {
    const bool zlo = (z == 0) && g.zlo_phys, zhi = (z == g.nz - 1) && g.zhi_phys;
    const T dxx = D.p[XX3][c], dyy = D.p[YY3][c], dzz = D.p[ZZ3][c];
    const T dxy = D.p[XY3][c], dxz = D.p[XZ3][c], dyz = D.p[YZ3][c];
    const T ax = k.wx * dxx, ay = k.wy * dyy, az = k.wz * dzz;
    r.diag = T(1) + T(2) * (ax + ay + az);
    const T bx = -(k.bxx * tensor_diff<T>(D.p[XX3], c, 1, xlo, xhi) + k.bxy * tensor_diff<T>(D.p[XY3], c, g.pitch, ylo, yhi) +
                   k.bxz * tensor_diff<T>(D.p[XZ3], c, g.plane, zlo, zhi));
    const T by = -(k.bxy * tensor_diff<T>(D.p[XY3], c, 1, xlo, xhi) + k.byy * tensor_diff<T>(D.p[YY3], c, g.pitch, ylo, yhi) +
                   k.byz * tensor_diff<T>(D.p[YZ3], c, g.plane, zlo, zhi));
    const T bz = -(k.bxz * tensor_diff<T>(D.p[XZ3], c, 1, xlo, xhi) + k.byz * tensor_diff<T>(D.p[YZ3], c, g.pitch, ylo, yhi) +
                   k.bzz * tensor_diff<T>(D.p[ZZ3], c, g.plane, zlo, zhi));
    r.xp = -ax + bx; r.xm = -ax - bx;
    r.yp = -ay + by; r.ym = -ay - by;
    r.zp = -az + bz; r.zm = -az - bz;
    r.exy = -k.cxy * dxy; r.exz = -k.cxz * dxz; r.eyz = -k.cyz * dyz;
  } else {
    const T dxx = D.p[XX2][c], dyy = D.p[YY2][c], dxy = D.p[XY2][c];
    const T ax = k.wx * dxx, ay = k.wy * dyy;
    r.diag = T(1) + T(2) * (ax + ay);
    const T bx = -(k.bxx * tensor_diff<T>(D.p[XX2], c, 1, xlo, xhi) + k.bxy * tensor_diff<T>(D.p[XY2], c, g.pitch, ylo, yhi));
    const T by = -(k.bxy * tensor_diff<T>(D.p[XY2], c, 1, xlo, xhi) + k.byy * tensor_diff<T>(D.p[YY2], c, g.pitch, ylo, yhi));
    r.xp = -ax + bx; r.xm = -ax - bx;
    r.yp = -ay + by; r.ym = -ay - by;
    r.zp = r.zm = T(0);
    r.exy = -k.cxy * dxy; r.exz = r.eyz = T(0);
  }
}

// Sum over the off-diagonal entries, a_k u_k, with node-mirrored reads.
template <int DIM, typename T, typename UT>
__host__ __device__ __forceinline__ T apply_offdiag(const Geom& g, const Row<T>& r, const UT* __restrict__ u, int x, int y, int z)
{
  const long long c = (long long)z * g.plane + (long long)y * g.pitch + x;
  const long long xm = (x == 0) ? 1 : -1, xp = (x == g.nx - 1) ? -1 : 1;
  const long long ym = (y == 0) ? g.pitch : -(long long)g.pitch, yp = (y == g.ny - 1) ? -(long long)g.pitch : g.pitch;
  T s = r.xp * T(u[c + xp]) + r.xm * T(u[c + xm]) + r.yp * T(u[c + yp]) + r.ym * T(u[c + ym]);
  s += r.exy * ((T(u[c + xp + yp]) - T(u[c + xp + ym])) - (T(u[c + xm + yp]) - T(u[c + xm + ym])));
  if (DIM == 3) {
    const long long zm = (z == 0 && g.zlo_phys) ? g.plane : -g.plane, zp = (z == g.nz - 1 && g.zhi_phys) ? -g.plane : g.plane;
    s += r.zp * T(u[c + zp]) + r.zm * T(u[c + zm]);
    s += r.exz * ((T(u[c + xp + zp]) - T(u[c + xp + zm])) - (T(u[c + xm + zp]) - T(u[c + xm + zm])));
    s += r.eyz * ((T(u[c + yp + zp]) - T(u[c + yp + zm])) - (T(u[c + ym + zp]) - T(u[c + ym + zm])));
  }
  return s;
}

// Scatter a row to explicit 3^DIM storage in Neighborhood raster order; offsets that leave the
// grid are redirected to their node mirror, as GenerateDCA does (mad/itkGridsHierarchy.hxx:362-430).
template <int DIM, typename T>
__host__ __device__ __forceinline__ void scatter_row(const Geom& g, const Row<T>& r, int x, int y, int z, T* S)
{
  constexpr int NS = DIM == 2 ? 9 : 27;
  for (int i = 0; i < NS; ++i) S[i] = T(0);
  const int xm = (x == 0) ? 1 : -1, xp = (x == g.nx - 1) ? -1 : 1;
  const int ym = (y == 0) ? 1 : -1, yp = (y == g.ny - 1) ? -1 : 1;
  const int zm = (z == 0 && g.zlo_phys) ? 1 : -1, zp = (z == g.nz - 1 && g.zhi_phys) ? -1 : 1;
#define MAD_AT(ox, oy, oz) S[DIM == 2 ? ((oy) + 1) * 3 + ((ox) + 1) : (((oz) + 1) * 3 + ((oy) + 1)) * 3 + ((ox) + 1)]
  MAD_AT(0, 0, 0) += r.diag;
  MAD_AT(xp, 0, 0) += r.xp; MAD_AT(xm, 0, 0) += r.xm;
  MAD_AT(0, yp, 0) += r.yp; MAD_AT(0, ym, 0) += r.ym;
  MAD_AT(xp, yp, 0) += r.exy; MAD_AT(xp, ym, 0) -= r.exy; MAD_AT(xm, yp, 0) -= r.exy; MAD_AT(xm, ym, 0) += r.exy;
  if (DIM == 3) {
    MAD_AT(0, 0, zp) += r.zp; MAD_AT(0, 0, zm) += r.zm;
    MAD_AT(xp, 0, zp) += r.exz; MAD_AT(xp, 0, zm) -= r.exz; MAD_AT(xm, 0, zp) -= r.exz; MAD_AT(xm, 0, zm) += r.exz;
    MAD_AT(0, yp, zp) += r.eyz; MAD_AT(0, yp, zm) -= r.eyz; MAD_AT(0, ym, zp) -= r.eyz; MAD_AT(0, ym, zm) += r.eyz;
  }
#undef MAD_AT
}

#if defined(__CUDACC__) || defined(MAD_HOST_EMULATION)  // the CPU test build (tests/mad_host/) runs these kernels on host fibres

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum (deterministic order); result valid in thread 0.
__device__ __forceinline__ double block_sum(double v)
{
  __shared__ double sh[32];
  const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  const int nthr = blockDim.x * blockDim.y * blockDim.z;
  v = warp_sum(v);
  if ((tid & 31) == 0) sh[tid >> 5] = v;
  __syncthreads();
  double t = 0;
  if (tid < 32) {
    t = (tid < (nthr + 31) / 32) ? sh[tid] : 0.0;
    t = warp_sum(t);
  }
  __syncthreads();
  return t;
}

// ------------------------------------------------------------------------------------------
// v1 kernels: one voxel per thread, 3-D thread blocks, neighbour reuse through L1/L2.
// ------------------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(512) k_jacobi(Geom g, Tensor D, const float* __restrict__ u, const float* __restrict__ f,
                                                float* __restrict__ out, float omega)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int z = blockIdx.z * blockDim.z + threadIdx.z;
  if (x >= g.nx || y >= g.ny || z >= g.nz) return;
  Row<float> r;
  row_coeffs<DIM, float>(g, D, x, y, z, r);
  const float off = apply_offdiag<DIM, float, float>(g, r, u, x, y, z);
  const long long c = (long long)z * g.plane + (long long)y * g.pitch + x;
  // mad/itkMultigridWeightedJacobiSmoother.hxx:88-89
  out[c] = (f[c] - off) * (omega / r.diag) + (1.0f - omega) * u[c];
}

__device__ __forceinline__ int voxel_color(int dim, int ncolors, int x, int y, int zg)
{
  if (dim == 2) return (x & 1) | ((y & 1) << 1);
  const int p = (x & 1) | ((y & 1) << 1) | ((zg & 1) << 2);
  return ncolors == 8 ? p : (p < 7 - p ? p : 7 - p);
}

// One colour of a multicolour Gauss-Seidel sweep, in place.  No voxel of a colour is a stencil
// neighbour of another voxel of the same colour (19-point: parity classes p and 7-p only differ
// along a cube diagonal, which the operator does not couple), so the update order inside a
// colour is irrelevant and the sweep is a valid Gauss-Seidel ordering.
template <int DIM>
__global__ void __launch_bounds__(512) k_gs_color(Geom g, Tensor D, float* __restrict__ u, const float* __restrict__ f, int color,
                                                  int ncolors)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int z = blockIdx.z * blockDim.z + threadIdx.z;
  if (x >= g.nx || y >= g.ny || z >= g.nz) return;
  if (voxel_color(DIM, ncolors, x, y, z + g.z0) != color) return;
  Row<float> r;
  row_coeffs<DIM, float>(g, D, x, y, z, r);
  const float off = apply_offdiag<DIM, float, float>(g, r, u, x, y, z);
  const long long c = (long long)z * g.plane + (long long)y * g.pitch + x;
  u[c] = (f[c] - off) / r.diag;  // mad/itkMultigridGaussSeidelSmoother.hxx:99
}

// r = f - A u, optional per-block partial sums of r^2 (double).
template <int DIM, typename T, typename UT, typename FT, typename RT>
__global__ void __launch_bounds__(512) k_residual(Geom g, Tensor D, const UT* __restrict__ u, const FT* __restrict__ f, RT* __restrict__ res,
                                                  double* __restrict__ partials)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int z = blockIdx.z * blockDim.z + threadIdx.z;
  double sq = 0.0;
  if (x < g.nx && y < g.ny && z < g.nz) {
    Row<T> r;
    row_coeffs<DIM, T>(g, D, x, y, z, r);
    const T off = apply_offdiag<DIM, T, UT>(g, r, u, x, y, z);
    const long long c = (long long)z * g.plane + (long long)y * g.pitch + x;
    const T v = T(f[c]) - r.diag * T(u[c]) - off;
    if (res) res[c] = RT(v);
    sq = (double)v * (double)v;
  }
  if (partials) {
    const double t = block_sum(sq);
    if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0)
      partials[(size_t)blockIdx.x + (size_t)gridDim.x * (blockIdx.y + (size_t)gridDim.y * blockIdx.z)] = t;
  }
}

// sum of squares of a pitched field (L2Norm, …Filter.hxx:496-515)
template <typename FT>
__global__ void __launch_bounds__(512) k_sumsq(Geom g, const FT* __restrict__ f, double* __restrict__ partials)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int z = blockIdx.z * blockDim.z + threadIdx.z;
  double sq = 0.0;
  if (x < g.nx && y < g.ny && z < g.nz) {
    const double v = (double)f[(long long)z * g.plane + (long long)y * g.pitch + x];
    sq = v * v;
  }
  const double t = block_sum(sq);
  if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0)
    partials[(size_t)blockIdx.x + (size_t)gridDim.x * (blockIdx.y + (size_t)gridDim.y * blockIdx.z)] = t;
}

// Peer-memory halo, bounded wait (MADGPU_P2P_WAIT=kernel): one thread polls this rank's two arrival counters until both have
// reached the wanted sequence numbers (cyclic compare, as cuStreamWaitValue32's GEQ) or `timeout_cycles` have passed.  A time-out
// does not stop the stream -- the kernels behind it run on stale ghost planes -- but it is recorded: fail[0] = 1 travels with the
// next residual-norm all-reduce to every rank, diag = {wanted lower, wanted upper, seen lower, seen upper}.
__global__ void k_halo_wait(const volatile unsigned* flags, unsigned need_lo, unsigned need_hi, int has_lo, int has_hi, long long timeout_cycles,
                            double* fail, unsigned* diag)
{
  if (fail[0] != 0.0) return;  // an earlier wait of this cycle already timed out: do not pay the time-out again
  const long long t0 = clock64();
  bool ok_lo = !has_lo, ok_hi = !has_hi;
  unsigned seen_lo = 0, seen_hi = 0;
  while (!(ok_lo && ok_hi)) {
    if (!ok_lo) { seen_lo = flags[0]; ok_lo = (int)(seen_lo - need_lo) >= 0; }
    if (!ok_hi) { seen_hi = flags[1]; ok_hi = (int)(seen_hi - need_hi) >= 0; }
    if (clock64() - t0 > timeout_cycles) break;
  }
  if (!(ok_lo && ok_hi)) {
    fail[0] = 1.0;
    diag[0] = need_lo; diag[1] = need_hi; diag[2] = seen_lo; diag[3] = seen_hi;
  }
}

// out[0] = sum(partials[0..n)) in a fixed order (one block).
__global__ void __launch_bounds__(1024) k_reduce_partials(const double* __restrict__ partials, long long n, double* __restrict__ out)
{
  double s = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) s += partials[i];
  const double t = block_sum(s);
  if (threadIdx.x == 0) out[0] = t;
}

// Explicit operator rows in Neighborhood raster order (parity test of GenerateDCA): the interior
// coefficients are scattered to their node-mirrored offsets exactly as the reference redirects them.
template <int DIM>
__global__ void k_assemble(Geom g, Tensor D, float* __restrict__ stencil /* dense nvox * 3^DIM */)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int z = blockIdx.z * blockDim.z + threadIdx.z;
  if (x >= g.nx || y >= g.ny || z >= g.nz) return;
  constexpr int NS = DIM == 2 ? 9 : 27;
  Row<float> r;
  row_coeffs<DIM, float>(g, D, x, y, z, r);
  float S[NS];
  scatter_row<DIM, float>(g, r, x, y, z, S);
  float* o = stencil + ((long long)(z * (long long)g.ny + y) * g.nx + x) * NS;
#pragma unroll
  for (int i = 0; i < NS; ++i) o[i] = S[i];
}

// ------------------------------------------------------------------------------------------
// Inter-grid transfers.  cent[d]: 0 vertex-centred (fine n odd), 1 cell-centred (fine n even).
// ------------------------------------------------------------------------------------------
struct Transfer {
  int cent[3];
};

// Gather taps of full weighting along one axis for coarse index c: fine indices 2c-1..2c+2.
// Tables: mad/itkInterGridOperators.h:115-127 (vertex ends are injection).
// lo / hi: the first / last index of this axis is a physical boundary (false at an inner face of a z-slab,
// whose ghost planes hold the neighbour's values: interior weights apply there).
__device__ __forceinline__ void restrict_taps(int c, int nc, int cent, float w[4], bool lo = true, bool hi = true)
{
  const bool first = c == 0 && lo, last = c == nc - 1 && hi;
  if (cent == 0) {
    if (first || last) { w[0] = 0.f; w[1] = 1.f; w[2] = 0.f; w[3] = 0.f; }
    else { w[0] = .25f; w[1] = .5f; w[2] = .25f; w[3] = 0.f; }
  } else {
    if (first) { w[0] = 0.f; w[1] = .5f; w[2] = .375f; w[3] = .125f; }
    else if (last) { w[0] = .125f; w[1] = .375f; w[2] = .5f; w[3] = 0.f; }
    else { w[0] = .125f; w[1] = .375f; w[2] = .375f; w[3] = .125f; }
  }
}

template <int DIM, typename TI, typename TO>
__global__ void __launch_bounds__(256) k_restrict(Geom gf, Geom gc, Transfer t, const TI* __restrict__ fine, TO* __restrict__ coarse)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int z = blockIdx.z * blockDim.z + threadIdx.z;
  if (x >= gc.nx || y >= gc.ny || z >= gc.nz) return;
  float wx[4], wy[4], wz[4];
  restrict_taps(x, gc.nx, t.cent[0], wx);
  restrict_taps(y, gc.ny, t.cent[1], wy);
  if (DIM == 3) restrict_taps(z, gc.nz, t.cent[2], wz, gc.zlo_phys != 0, gc.zhi_phys != 0);
  typedef typename std::conditional<std::is_same<TI, double>::value, double, float>::type A;
  A acc = 0;
  constexpr int KZ = DIM == 3 ? 4 : 1;
#pragma unroll
  for (int kz = 0; kz < KZ; ++kz) {
    const float wzz = DIM == 3 ? wz[kz] : 1.f;
    if (wzz == 0.f) continue;
    const int fz = DIM == 3 ? 2 * z + kz - 1 : 0;
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
      const float wyz = wy[ky] * wzz;
      if (wyz == 0.f) continue;
      const int fy = 2 * y + ky - 1;
      const TI* row = fine + (long long)fz * gf.plane + (long long)fy * gf.pitch;
#pragma unroll
      for (int kx = 0; kx < 4; ++kx) {
        if (wx[kx] == 0.f) continue;
        acc += A(wx[kx] * wyz) * A(row[2 * x + kx - 1]);
      }
    }
  }
  coarse[(long long)z * gc.plane + (long long)y * gc.pitch + x] = TO(acc);
}

// Gather form of the reference's scatter interpolation along one axis: fine index x takes
// w0*c[i0] + w1*c[i1].  vertex: f[2i]=c[i], f[2i+1]=(c[i]+c[i+1])/2.  cell: f[2i]=3/4 c[i]+1/4 c[i-1],
// f[2i+1]=3/4 c[i]+1/4 c[i+1], f[0]=c[0], f[n-1]=c[nc-1]  (mad/itkInterGridOperators.h:101-113).
__device__ __forceinline__ void prolong_taps(int x, int nf, int nc, int cent, int& i0, int& i1, float& w0, float& w1, bool lo = true,
                                             bool hi = true)
{
  const int i = x >> 1;
  if (cent == 0) {
    if ((x & 1) == 0) { i0 = i1 = i; w0 = 1.f; w1 = 0.f; }
    else { i0 = i; i1 = i + 1; w0 = .5f; w1 = .5f; }
  } else {
    if (x == 0 && lo) { i0 = i1 = 0; w0 = 1.f; w1 = 0.f; }
    else if (x == nf - 1 && hi) { i0 = i1 = nc - 1; w0 = 1.f; w1 = 0.f; }
    else if ((x & 1) == 0) { i0 = i; i1 = i - 1; w0 = .75f; w1 = .25f; }
    else { i0 = i; i1 = i + 1; w0 = .75f; w1 = .25f; }
  }
}

// fine (+)= P coarse
template <int DIM, typename TO, bool ADD>
__global__ void __launch_bounds__(256) k_prolong(Geom gc, Geom gf, Transfer t, const float* __restrict__ coarse, TO* __restrict__ fine)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int z = blockIdx.z * blockDim.z + threadIdx.z;
  if (x >= gf.nx || y >= gf.ny || z >= gf.nz) return;
  int x0, x1, y0, y1, z0 = 0, z1 = 0;
  float wx0, wx1, wy0, wy1, wz0 = 1.f, wz1 = 0.f;
  prolong_taps(x, gf.nx, gc.nx, t.cent[0], x0, x1, wx0, wx1);
  prolong_taps(y, gf.ny, gc.ny, t.cent[1], y0, y1, wy0, wy1);
  if (DIM == 3) prolong_taps(z, gf.nz, gc.nz, t.cent[2], z0, z1, wz0, wz1, gf.zlo_phys != 0, gf.zhi_phys != 0);
  const float* p00 = coarse + (long long)z0 * gc.plane + (long long)y0 * gc.pitch;
  const float* p01 = coarse + (long long)z0 * gc.plane + (long long)y1 * gc.pitch;
  float v = wy0 * (wx0 * p00[x0] + wx1 * p00[x1]) + wy1 * (wx0 * p01[x0] + wx1 * p01[x1]);
  if (DIM == 3) {
    const float* p10 = coarse + (long long)z1 * gc.plane + (long long)y0 * gc.pitch;
    const float* p11 = coarse + (long long)z1 * gc.plane + (long long)y1 * gc.pitch;
    const float v1 = wy0 * (wx0 * p10[x0] + wx1 * p10[x1]) + wy1 * (wx0 * p11[x0] + wx1 * p11[x1]);
    v = wz0 * v + wz1 * v1;
  }
  const long long c = (long long)z * gf.plane + (long long)y * gf.pitch + x;
  if (ADD) fine[c] += TO(v);
  else fine[c] = TO(v);
}

// e = Ainv f on the coarsest grid; Ainv dense row-major fp64 in LexPosition order
// (mad/itkDirectSolver.h:89-99); f, e pitched.  One warp per row.
__global__ void __launch_bounds__(256) k_coarse_gemv(Geom g, const double* __restrict__ Ainv, const float* __restrict__ f, float* __restrict__ e,
                                                     int n)
{
  MAD_DYNAMIC_SHARED(double, sf);
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const int x = j % g.nx, y = (j / g.nx) % g.ny, z = j / (g.nx * g.ny);
    sf[j] = (double)f[(long long)z * g.plane + (long long)y * g.pitch + x];
  }
  __syncthreads();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const double* a = Ainv + (size_t)row * n;
  double s = 0.0;
  for (int j = lane; j < n; j += 32) s += a[j] * sf[j];
  s = warp_sum(s);
  if (lane == 0) {
    const int x = row % g.nx, y = (row / g.nx) % g.ny, z = row / (g.nx * g.ny);
    e[(long long)z * g.plane + (long long)y * g.pitch + x] = (float)s;
  }
}

// ---- dense inverse of the coarsest-grid operator, built on the device (replaces the vnl_sparse_lu factorisation of
// ---- mad/itkDirectSolver.hxx:44-86; a host LU of 512 unknowns cost 50-300 ms per tensor, this costs a few) -----------------
// M = [A | I], row-major, leading dimension ld = 2n, zero-initialised by the caller.  Row = LexPosition (mad/itkDirectSolver.h:89-99).
template <int DIM>
__global__ void k_coarse_matrix(Geom g, Tensor D, double* __restrict__ M, int n, int ld)
{
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  const int x = row % g.nx, y = (row / g.nx) % g.ny, z = row / (g.nx * g.ny);
  Row<double> r;
  row_coeffs<DIM, double>(g, D, x, y, z, r);
  double S[DIM == 2 ? 9 : 27];
  scatter_row<DIM, double>(g, r, x, y, z, S);
  const int zl = DIM == 3 ? -1 : 0, zh = DIM == 3 ? 1 : 0;
  double* m = M + (size_t)row * ld;
  for (int oz = zl; oz <= zh; ++oz)
    for (int oy = -1; oy <= 1; ++oy)
      for (int ox = -1; ox <= 1; ++ox) {
        const int xx = x + ox, yy = y + oy, zz = z + oz;
        if (xx < 0 || xx >= g.nx || yy < 0 || yy >= g.ny || zz < 0 || zz >= g.nz) continue;  // folded onto the mirror by scatter_row
        const int si = DIM == 2 ? (oy + 1) * 3 + (ox + 1) : ((oz + 1) * 3 + (oy + 1)) * 3 + (ox + 1);
        m[((size_t)zz * g.ny + yy) * g.nx + xx] = S[si];
      }
  m[n + row] = 1.0;
}

// Gauss-Jordan step k, part 1 (one block): partial pivoting over rows k..n-1 of column k, row swap, pivot row scaled to a unit
// pivot, column k saved to colk[] (the elimination overwrites it).  singular[0] is set when no pivot is left.
__global__ void __launch_bounds__(1024) k_gj_pivot(double* __restrict__ M, int n, int ld, int k, double* __restrict__ colk, int* __restrict__ singular)
{
  __shared__ double sv[32];
  __shared__ int si[32];
  __shared__ int prow;
  __shared__ double pinv;
  const int tid = threadIdx.x;
  double best = -1.0;
  int bi = k;
  for (int i = k + tid; i < n; i += blockDim.x) {
    const double v = fabs(M[(size_t)i * ld + k]);
    if (v > best) { best = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if ((tid & 31) == 0) { sv[tid >> 5] = best; si[tid >> 5] = bi; }
  __syncthreads();
  if (tid < 32) {
    best = tid < (int)((blockDim.x + 31) >> 5) ? sv[tid] : -1.0;
    bi = tid < (int)((blockDim.x + 31) >> 5) ? si[tid] : k;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (tid == 0) {
      prow = bi;
      if (!(best > 0.0)) { singular[0] = 1; pinv = 0.0; }
      else pinv = 1.0 / M[(size_t)bi * ld + k];
    }
  }
  __syncthreads();
  const int p = prow;
  const double inv = pinv;
  double* rk = M + (size_t)k * ld;
  double* rp = M + (size_t)p * ld;
  for (int j = tid; j < ld; j += blockDim.x) {
    const double a = rk[j], b = rp[j];
    rk[j] = b * inv;
    if (p != k) rp[j] = a;
  }
  __syncthreads();
  for (int i = tid; i < n; i += blockDim.x) colk[i] = i == k ? 0.0 : M[(size_t)i * ld + k];
}

// part 2: rows i != k lose their column-k entry, M[i][:] -= colk[i] * M[k][:].  grid = (ceil(ld/256), ceil(n/8)), block = 256
__global__ void __launch_bounds__(256) k_gj_eliminate(double* __restrict__ M, int n, int ld, int k, const double* __restrict__ colk)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ld) return;
  const double pk = M[(size_t)k * ld + j];
  const int i0 = blockIdx.y * 8;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int i = i0 + r;
    if (i < n && i != k) {
      const double c = colk[i];
      if (c != 0.0) M[(size_t)i * ld + j] -= c * pk;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Element-wise helpers (pitched <-> dense, casts, axpy).
// ------------------------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void k_dense_to_pitched(Geom g, const TI* __restrict__ in, TO* __restrict__ out)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int z = blockIdx.z * blockDim.z + threadIdx.z;
  if (x >= g.nx || y >= g.ny || z >= g.nz) return;
  out[(long long)z * g.plane + (long long)y * g.pitch + x] = TO(in[((long long)z * g.ny + y) * g.nx + x]);
}

// static_cast<OutputPixelType>(double) of the reference (…Filter.hxx:277): truncation toward zero for
// integer pixels (values outside the pixel range are clamped instead of undefined).
template <typename TO>
__device__ __forceinline__ TO cast_out(double v) { return TO(v); }
template <>
__device__ __forceinline__ uint8_t cast_out<uint8_t>(double v)
{
  const double t = trunc(v);
  return (uint8_t)(t < 0.0 ? 0.0 : (t > 255.0 ? 255.0 : t));
}
template <>
__device__ __forceinline__ int16_t cast_out<int16_t>(double v)
{
  const double t = trunc(v);
  return (int16_t)(t < -32768.0 ? -32768.0 : (t > 32767.0 ? 32767.0 : t));
}

template <typename TI, typename TO>
__global__ void k_pitched_to_dense(Geom g, const TI* __restrict__ in, TO* __restrict__ out)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int z = blockIdx.z * blockDim.z + threadIdx.z;
  if (x >= g.nx || y >= g.ny || z >= g.nz) return;
  out[((long long)z * g.ny + y) * g.nx + x] = cast_out<TO>((double)in[(long long)z * g.plane + (long long)y * g.pitch + x]);
}

// u64 += e32  (correction add of the outer defect-correction loop); four voxels per thread (rows are padded to a
// multiple of 32 elements, so the 16/32-byte accesses never leave the row's allocation)
__global__ void __launch_bounds__(256) k_axpy_f64_f32(Geom g, double* __restrict__ u, const float* __restrict__ e)
{
  const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int z = blockIdx.z;
  if (x >= g.nx || y >= g.ny) return;
  const long long c = (long long)z * g.plane + (long long)y * g.pitch + x;
  const float4 q = *reinterpret_cast<const float4*>(e + c);
  double2 a = *reinterpret_cast<double2*>(u + c), b = *reinterpret_cast<double2*>(u + c + 2);
  a.x += (double)q.x; a.y += (double)q.y; b.x += (double)q.z; b.y += (double)q.w;
  *reinterpret_cast<double2*>(u + c) = a;
  *reinterpret_cast<double2*>(u + c + 2) = b;
  double* gp = z == 0 ? (double*)g.glo : z == g.nz - 1 ? (double*)g.ghi : nullptr;
  if (gp) {
    const long long r = (long long)y * g.pitch + x;
    *reinterpret_cast<double2*>(gp + r) = a;
    *reinterpret_cast<double2*>(gp + r + 2) = b;
  }
  if (g.nz == 1 && g.glo && g.ghi) {  // a one-plane slab feeds both neighbours
    const long long r = (long long)y * g.pitch + x;
    *reinterpret_cast<double2*>((double*)g.ghi + r) = a;
    *reinterpret_cast<double2*>((double*)g.ghi + r + 2) = b;
  }
}

// AoS tensor chunk (ITK SymmetricSecondRankTensor buffer) -> SoA fp32 pitched planes.
// first = linear dense voxel index of the first voxel of the chunk.
template <typename TI, int NCOMP>
__global__ void k_tensor_ingest(Geom g, const TI* __restrict__ aos, long long first, long long count, float* const* __restrict__ planes_unused,
                                float* p0, float* p1, float* p2, float* p3, float* p4, float* p5)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const long long v = first + i;
  const int x = (int)(v % g.nx);
  const long long t = v / g.nx;
  const int y = (int)(t % g.ny);
  const long long z = t / g.ny;
  const long long c = z * g.plane + (long long)y * g.pitch + x;
  const TI* a = aos + i * NCOMP;
  p0[c] = (float)a[0]; p1[c] = (float)a[1]; p2[c] = (float)a[2];
  if (NCOMP == 6) { p3[c] = (float)a[3]; p4[c] = (float)a[4]; p5[c] = (float)a[5]; }
}

#endif  // __CUDACC__

}  // namespace mad
