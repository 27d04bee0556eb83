// ved_kernels.cuh -- the CUDA kernels of the VED tensor front-end (ved.cu) together with their launch geometry and the pass
// structure of the separable Hessian.  Everything that decides WHAT is computed WHERE lives here, so that the CPU suite can run
// this very source: tests/ved_host_harness.cpp includes this file after tests/mad_host/fiber_shim.h, which maps the CUDA
// built-ins (threadIdx, __shared__, __syncwarp, the <<<>>> launch behind VED_LAUNCH) onto host fibres.  ved.cu adds only the context, the
// C-ABI and the copies.  Arithmetic: ved_math.h.  Reference citations: ved.cu / ved_math.h.
#ifndef MADGPU_VED_KERNELS_CUH
#define MADGPU_VED_KERNELS_CUH

#include <stdint.h>

#include "ved_math.h"

#ifndef VED_LAUNCH  // the host shim supplies its own
#define VED_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#endif

namespace vedk
{
template <int K>
struct RgArgs {
  ved::RgCoefs c[K];
  float* out[K];
  double scale[K];  // applied to causal + anticausal (1 / (spacing_a * spacing_b) on the last pass of a Hessian component)
};

// ---- recursive Gaussian along y or z ---------------------------------------------------------------------------------
// line t: first element (t / inner) * outer_stride + (t % inner), n elements `stride` apart.
//   y pass: inner = nx, outer_stride = nx * ny, stride = nx, lines = nx * nz;   z pass: inner = lines = nx * ny, stride = nx * ny.
template <int K>
__global__ void __launch_bounds__(128) k_rg_lines(const float* __restrict__ in, RgArgs<K> a, int n, long long stride, long long inner,
                                                   long long outer_stride, long long nlines)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nlines) return;
  const long long base = (t / inner) * outer_stride + (t % inner);
  float* out[K];
#pragma unroll
  for (int k = 0; k < K; ++k) out[k] = a.out[k] + base;
  ved::rg_line<K>(in + base, stride, n, a.c, out, a.scale);
}

// ---- recursive Gaussian along x ---------------------------------------------------------------------------------------
constexpr int RG_ROW_WARPS = 2;  // warps per CTA; (1 + K) tiles of 32 x 33 floats per warp: 33.8 KB of shared memory at K = 3

template <int K>
__global__ void __launch_bounds__(32 * RG_ROW_WARPS) k_rg_rows(const float* __restrict__ in, RgArgs<K> a, int nx, long long nrows)
{
  __shared__ float tin[RG_ROW_WARPS][32][33];
  __shared__ float tout[RG_ROW_WARPS][K][32][33];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const long long row0 = ((long long)blockIdx.x * RG_ROW_WARPS + wp) * 32;
  if (row0 >= nrows) return;  // warp-uniform: the warps of a CTA never synchronise with each other
  const long long myrow = row0 + lane;
  const bool mine = myrow < nrows;
  ved::RgState s[K];
  const int nchunks = (nx + 31) / 32;

  const double e0 = mine ? (double)in[myrow * nx] : 0.0;
#pragma unroll
  for (int k = 0; k < K; ++k) ved::rg_causal_init(s[k], a.c[k], e0);
  for (int ch = 0; ch < nchunks; ++ch) {
    const int c0 = ch * 32, w = min(32, nx - c0);
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      const long long row = row0 + r;
      tin[wp][r][lane] = (row < nrows && lane < w) ? in[row * nx + c0 + lane] : 0.f;
    }
    __syncwarp();
    for (int j = 0; j < w; ++j) {
      const double xi = (double)tin[wp][lane][j];
#pragma unroll
      for (int k = 0; k < K; ++k) tout[wp][k][lane][j] = (float)ved::rg_causal_step(s[k], a.c[k], xi);
    }
    __syncwarp();
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      const long long row = row0 + r;
      if (row < nrows && lane < w) {
#pragma unroll
        for (int k = 0; k < K; ++k) a.out[k][row * nx + c0 + lane] = tout[wp][k][r][lane];
      }
    }
    __syncwarp();
  }

  const double e1 = mine ? (double)in[myrow * nx + nx - 1] : 0.0;
#pragma unroll
  for (int k = 0; k < K; ++k) ved::rg_anti_init(s[k], a.c[k], e1);
  for (int ch = nchunks - 1; ch >= 0; --ch) {
    const int c0 = ch * 32, w = min(32, nx - c0);
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      const long long row = row0 + r;
      const bool ok = row < nrows && lane < w;
      tin[wp][r][lane] = ok ? in[row * nx + c0 + lane] : 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) tout[wp][k][r][lane] = ok ? a.out[k][row * nx + c0 + lane] : 0.f;
    }
    __syncwarp();
    for (int j = w - 1; j >= 0; --j) {
      const double xi = (double)tin[wp][lane][j];
#pragma unroll
      for (int k = 0; k < K; ++k)
        tout[wp][k][lane][j] = (float)(((double)tout[wp][k][lane][j] + ved::rg_anti_step(s[k], a.c[k], xi)) * a.scale[k]);
    }
    __syncwarp();
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      const long long row = row0 + r;
      if (row < nrows && lane < w) {
#pragma unroll
        for (int k = 0; k < K; ++k) a.out[k][row * nx + c0 + lane] = tout[wp][k][r][lane];
      }
    }
    __syncwarp();
  }
}

// ---- eigen-system + vesselness + tensor ----------------------------------------------------------------------------------
struct TensorPlanes {
  float* p[6];
};
struct HessianPlanes {
  const float* p[6];
};

// SOA: six fp32 planes of this context (offset 0); otherwise a chunk of the caller's AoS fp64 buffer starting at voxel `first_voxel`
template <bool SOA>
__global__ void __launch_bounds__(128) k_ved_update(long long first_voxel, long long count, HessianPlanes hs, const double* __restrict__ aos, int first,
                                                     ved::Params P, double* __restrict__ response, TensorPlanes T)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const long long v = first_voxel + i;
  double h[6];
  if (SOA) {
#pragma unroll
    for (int k = 0; k < 6; ++k) h[k] = (double)hs.p[k][v];
  } else {
#pragma unroll
    for (int k = 0; k < 6; ++k) h[k] = aos[i * 6 + k];
  }
  double resp = first ? 0.0 : response[v];
  double t[6];
  if (ved::update_voxel(h, first != 0, P, resp, t)) {
    response[v] = resp;
#pragma unroll
    for (int k = 0; k < 6; ++k) T.p[k][v] = (float)t[k];
  }
}

template <typename TI>
__global__ void k_cast_in(const TI* __restrict__ in, float* __restrict__ out, long long n)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)in[i];
}

static __global__ void k_planes_to_aos_f64(HessianPlanes src, double* __restrict__ out, long long first_voxel, long long count)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
#pragma unroll
  for (int k = 0; k < 6; ++k) out[i * 6 + k] = (double)src.p[k][first_voxel + i];
}

// ---- launch geometry ------------------------------------------------------------------------------------------------------
struct Volume {
  long long nx, ny, nz;
  double h[3];
};

inline unsigned blocks_for(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

template <int K, typename Stream>
void launch_rows(Stream st, const Volume& v, const float* in, const RgArgs<K>& a)
{
  const long long nrows = v.ny * v.nz;
  const long long warps = (nrows + 31) / 32;
  VED_LAUNCH(k_rg_rows<K>, blocks_for(warps, RG_ROW_WARPS), 32 * RG_ROW_WARPS, st, in, a, (int)v.nx, nrows);
}

// axis 1 (y) or 2 (z)
template <int K, typename Stream>
void launch_lines(Stream st, const Volume& v, int axis, const float* in, const RgArgs<K>& a)
{
  const long long nlines = axis == 1 ? v.nx * v.nz : v.nx * v.ny;
  const long long inner = axis == 1 ? v.nx : v.nx * v.ny;
  const long long stride = axis == 1 ? v.nx : v.nx * v.ny;
  VED_LAUNCH(k_rg_lines<K>, blocks_for(nlines, 128), 128, st, in, a, (int)(axis == 1 ? v.ny : v.nz), stride, inner, v.nx * v.ny, nlines);
}

// ComputeHessian, itkVEDMultigridImageFilter.hxx:158-173 (HessianRecursiveGaussianImageFilter with NormalizeAcrossScale on).
// Separable: H_ab = (d_a d_b G) * I.  The x pass yields G0x, G1x, G2x of the image in one read; the y pass the six xy products
// from three reads; the z pass the six components, scaled by 1 / (h_a h_b).  W: twelve work volumes; H receives the six component
// planes in the order (0,0),(0,1),(0,2),(1,1),(1,2),(2,2) (they alias W[0..2] and W[9..11]).  Returns the number of launches.
template <typename Stream>
int hessian_passes(Stream st, const Volume& v, double sigma, const float* image, float* const* W, const float** H)
{
  ved::RgCoefs c[3][3];  // [axis][order]
  for (int ax = 0; ax < 3; ++ax)
    for (int o = 0; o < 3; ++o) ved::rg_setup(sigma, v.h[ax], o, true, c[ax][o]);
  float *G0x = W[0], *G1x = W[1], *G2x = W[2];
  float *Pxx = W[3], *Pxy = W[4], *Pxz = W[5], *Pyy = W[6], *Pyz = W[7], *Pzz = W[8];
  int launches = 0;
  {  // x pass
    RgArgs<3> a;
    for (int o = 0; o < 3; ++o) { a.c[o] = c[0][o]; a.scale[o] = 1.0; }
    a.out[0] = G0x; a.out[1] = G1x; a.out[2] = G2x;
    launch_rows<3>(st, v, image, a);
    ++launches;
  }
  {  // y pass
    RgArgs<3> a3;
    for (int o = 0; o < 3; ++o) { a3.c[o] = c[1][o]; a3.scale[o] = 1.0; }
    a3.out[0] = Pzz; a3.out[1] = Pyz; a3.out[2] = Pyy;  // G0x -> G0y, G1y, G2y
    launch_lines<3>(st, v, 1, G0x, a3);
    RgArgs<2> a2;
    for (int o = 0; o < 2; ++o) { a2.c[o] = c[1][o]; a2.scale[o] = 1.0; }
    a2.out[0] = Pxz; a2.out[1] = Pxy;  // G1x -> G0y, G1y
    launch_lines<2>(st, v, 1, G1x, a2);
    RgArgs<1> a1;
    a1.c[0] = c[1][0]; a1.scale[0] = 1.0;
    a1.out[0] = Pxx;  // G2x -> G0y
    launch_lines<1>(st, v, 1, G2x, a1);
    launches += 3;
  }
  {  // z pass; the x-pass volumes are free again and take three of the outputs
    const double* h = v.h;
    struct { const float* in; float* out; int order; double factor; } z[6] = {
        {Pxx, W[0], 0, h[0] * h[0]}, {Pxy, W[1], 0, h[0] * h[1]}, {Pxz, W[2], 1, h[0] * h[2]},
        {Pyy, W[9], 0, h[1] * h[1]}, {Pyz, W[10], 1, h[1] * h[2]}, {Pzz, W[11], 2, h[2] * h[2]}};
    for (int k = 0; k < 6; ++k) {
      RgArgs<1> a;
      a.c[0] = c[2][z[k].order];
      a.scale[0] = 1.0 / z[k].factor;
      a.out[0] = z[k].out;
      launch_lines<1>(st, v, 2, z[k].in, a);
      H[k] = z[k].out;
      ++launches;
    }
  }
  return launches;
}

// UpdateVesselness on six fp32 planes (the context's Hessian) ...
template <typename Stream>
void launch_update_planes(Stream st, long long nvox, const float* const* H, bool first, const ved::Params& P, double* response, float* const* T)
{
  HessianPlanes hs;
  TensorPlanes tp;
  for (int k = 0; k < 6; ++k) { hs.p[k] = H[k]; tp.p[k] = T[k]; }
  VED_LAUNCH(k_ved_update<true>, blocks_for(nvox, 128), 128, st, 0ll, nvox, hs, (const double*)nullptr, first ? 1 : 0, P, response, tp);
}

// ... or on a chunk of `count` voxels of a caller's AoS fp64 Hessian (device copy `aos`), voxels [first_voxel, first_voxel + count)
template <typename Stream>
void launch_update_aos(Stream st, long long first_voxel, long long count, const double* aos, bool first, const ved::Params& P, double* response,
                       float* const* T)
{
  HessianPlanes hs = {};
  TensorPlanes tp;
  for (int k = 0; k < 6; ++k) tp.p[k] = T[k];
  VED_LAUNCH(k_ved_update<false>, blocks_for(count, 128), 128, st, first_voxel, count, hs, aos, first ? 1 : 0, P, response, tp);
}

template <typename TI, typename Stream>
void launch_cast_in(Stream st, const TI* in, float* out, long long n)
{
  VED_LAUNCH(k_cast_in<TI>, blocks_for(n, 256), 256, st, in, out, n);
}

template <typename Stream>
void launch_planes_to_aos(Stream st, const float* const* planes, double* out, long long first_voxel, long long count)
{
  HessianPlanes src;
  for (int k = 0; k < 6; ++k) src.p[k] = planes[k];
  VED_LAUNCH(k_planes_to_aos_f64, blocks_for(count, 256), 256, st, src, out, first_voxel, count);
}
}  // namespace vedk

#endif  // MADGPU_VED_KERNELS_CUH
