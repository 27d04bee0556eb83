// ved_host_harness.cpp -- TEST INFRASTRUCTURE.  Compiles the VED front-end's kernel source for the HOST, so that the
// `-m "not gpu"` suite can check it against the oracle (oracle/ved_oracle.c) without a GPU:
//   csrc/ved_math.h       the arithmetic (coefficient set-up, line recursion with fp32 intermediate storage, eigen-solver,
//                         vesselness, per-voxel update), plain host functions here;
//   csrc/ved_kernels.cuh  the UNMODIFIED __global__ kernels, their launch geometry and the pass structure of the separable
//                         Hessian, run on host fibres through tests/mad_host/fiber_shim.h (threadIdx, __shared__, __syncwarp, the
//                         launch behind VED_LAUNCH) -- indexing, tiling and buffer reuse are exercised exactly as written.
// Built by tests/test_cpu_ved.py with g++ into tests/_build/.  The product never loads this file; the CUDA path never runs on
// the CPU (ved.cu, the context / C-ABI / copies around these kernels, needs a device).
#include <cstdint>
#include <cstring>
#include <vector>

#include "mad_host/fiber_shim.h"

#include "../multigridanisotropicdiffusion_b200/csrc/ved_kernels.cuh"

extern "C" {

void vh_rg_setup(double sigma, double spacing, int order, int normalize, double* out14)
{
  ved::RgCoefs c;
  ved::rg_setup(sigma, spacing, order, normalize != 0, c);
  std::memcpy(out14, &c, sizeof c);
}

// one line, one filter; fp32 in / out like the kernels
void vh_rg_line(double sigma, double spacing, int order, int normalize, const float* x, int n, float* y, double scale)
{
  ved::RgCoefs c;
  ved::rg_setup(sigma, spacing, order, normalize != 0, c);
  float* out[1] = {y};
  ved::rg_line<1>(x, 1, n, &c, out, &scale);
}

// hessian() of ved.cu: the kernels and launch geometry of ved_kernels.cuh executed by host fibres.
// n = (nx, ny, nz); image nvox floats; H six planes of nvox floats.  Returns the number of kernel launches.
int vh_hessian(const int* n, const double* h, double sigma, const float* image, float* H)
{
  const vedk::Volume v = {n[0], n[1], n[2], {h[0], h[1], h[2]}};
  const long long nvox = v.nx * v.ny * v.nz;
  std::vector<std::vector<float> > W(12, std::vector<float>(nvox, -12345.f));  // poisoned: every output voxel must be written
  float* w[12];
  for (int i = 0; i < 12; ++i) w[i] = W[i].data();
  const float* planes[6];
  const int launches = vedk::hessian_passes(0, v, sigma, image, w, planes);
  for (int k = 0; k < 6; ++k) std::memcpy(H + k * nvox, planes[k], sizeof(float) * nvox);
  return launches;
}

void vh_eig3_top(const double* h6, double* w, double* t) { ved::eig3_top(h6, w, t); }

double vh_vesselness(const double* e, double alpha, double beta, double gamma)
{
  const ved::Params P = {alpha, beta, gamma, 0.01, 5.0, 10.0};
  return ved::vesselness(e, P);
}

// k_ved_update through its launchers.  soa != 0: H is six fp32 planes; else hessian_aos holds 6 doubles per voxel and is consumed
// in chunks of `chunk` voxels like madved_update_vesselness_host_f64 does.  response nvox doubles, T six fp32 planes.
void vh_update(long long nvox, int soa, const float* H, const double* hessian_aos, int first, const double* params6, double* response, float* T,
               long long chunk)
{
  const ved::Params P = {params6[0], params6[1], params6[2], params6[3], params6[4], params6[5]};
  float* t[6];
  for (int k = 0; k < 6; ++k) t[k] = T + k * nvox;
  if (soa) {
    const float* hp[6];
    for (int k = 0; k < 6; ++k) hp[k] = H + k * nvox;
    vedk::launch_update_planes(0, nvox, hp, first != 0, P, response, t);
  } else {
    for (long long v0 = 0; v0 < nvox; v0 += chunk) {
      const long long cnt = std::min(chunk, nvox - v0);
      vedk::launch_update_aos(0, v0, cnt, hessian_aos + v0 * 6, first != 0, P, response, t);
    }
  }
}

// k_cast_in / k_planes_to_aos_f64 through their launchers
void vh_cast_in_i16(const int16_t* in, float* out, long long n) { vedk::launch_cast_in(0, in, out, n); }
void vh_cast_in_u8(const uint8_t* in, float* out, long long n) { vedk::launch_cast_in(0, in, out, n); }
void vh_cast_in_f64(const double* in, float* out, long long n) { vedk::launch_cast_in(0, in, out, n); }
void vh_planes_to_aos(const float* planes6, long long nvox, double* out, long long first_voxel, long long count)
{
  const float* p[6];
  for (int k = 0; k < 6; ++k) p[k] = planes6 + k * nvox;
  vedk::launch_planes_to_aos(0, p, out, first_voxel, count);
}

}  // extern "C"
