// ved_cabi_host.cpp -- TEST INFRASTRUCTURE.  multigridanisotropicdiffusion_b200/csrc/ved.cu compiled UNMODIFIED for the host:
// its kernels run on host fibres (tests/mad_host/fiber_shim.h), its CUDA runtime calls land in tests/fake_cuda/cuda_runtime.h, and the
// five madgpu_* entry points madved_run needs are answered by a stand-in solver backed by the CPU oracle (oracle/mad_oracle.c).
// The result, tests/_build/libmadved_host.so, exports the same madved_* C-ABI as libmadgpu.so, so tests/test_cpu_ved_cabi.py can
// drive the real context / staging / call-sequence code without a GPU.  Never part of the product.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "mad_host/fiber_shim.h"

#include "../multigridanisotropicdiffusion_b200/csrc/ved.cu"

// ---- oracle (test infrastructure) ----
extern "C" {
struct mo_hier;
mo_hier* mo_create(int dim, const int* n0, const double* h0, double dt, const double* tensor_aos, int smoother, double omega, int nu, int max_coarse);
void mo_destroy(mo_hier* H);
int mo_solve(mo_hier* H, int cycle, double tolerance, int max_cycles, int number_of_steps, double* image, int* cycles_per_step, double* relres_hist,
             int faithful);
}

// ---- stand-in for the solver context of madgpu.h: same entry points, the oracle underneath ----
struct madgpu_ctx {
  madgpu_params p;
  std::vector<double> tensor;  // AoS
  std::vector<double> result;  // the fp64 iterate of the last solve
  bool tensor_set;
  std::string err;
  int solves;
};

extern "C" {

madgpu_ctx* fake_solver_create(const int* n, const double* h, double dt, int smoother, int nu, int cycle, double tol, int steps)
{
  madgpu_ctx* c = new madgpu_ctx();
  std::memset(&c->p, 0, sizeof c->p);
  c->p.dim = 3;
  for (int d = 0; d < 3; ++d) { c->p.size[d] = n[d]; c->p.spacing[d] = h[d]; }
  c->p.time_step = dt; c->p.smoother = smoother; c->p.iterations_per_grid = nu; c->p.cycle = cycle; c->p.tolerance = tol;
  c->p.number_of_steps = steps; c->p.max_cycles = 100; c->p.omega = 2.0 / 3.0;
  c->tensor_set = false;
  c->solves = 0;
  return c;
}
void fake_solver_destroy(madgpu_ctx* c) { delete c; }
int fake_solver_solves(const madgpu_ctx* c) { return c->solves; }
long long fake_cuda_live_allocs() { return fake_cuda::g_live_allocs; }
void fake_cuda_set_devices(int n) { fake_cuda::g_devices = n; }

const char* madgpu_last_error(const madgpu_ctx* c) { return c ? c->err.c_str() : ""; }

int madgpu_level_info(const madgpu_ctx* c, int32_t level, int32_t size[3], double spacing[3], int32_t centering[3])
{
  if (!c || level != 0) return MADGPU_EINVAL;
  for (int d = 0; d < 3; ++d) { size[d] = c->p.size[d]; spacing[d] = c->p.spacing[d]; centering[d] = 0; }
  return 0;
}

int madgpu_set_tensor_device_f32(madgpu_ctx* c, const float* const* planes)
{
  const size_t nv = (size_t)c->p.size[0] * c->p.size[1] * c->p.size[2];
  c->tensor.resize(nv * 6);
  for (size_t v = 0; v < nv; ++v)
    for (int k = 0; k < 6; ++k) c->tensor[v * 6 + k] = (double)planes[k][v];
  c->tensor_set = true;
  return 0;
}

int madgpu_solve_device_f32(madgpu_ctx* c, const float* d_in, float* d_out, madgpu_stats* stats)
{
  if (!c->tensor_set) { c->err = "tensor not set"; return MADGPU_ESTATE; }
  const size_t nv = (size_t)c->p.size[0] * c->p.size[1] * c->p.size[2];
  if (d_in) {
    c->result.resize(nv);
    for (size_t v = 0; v < nv; ++v) c->result[v] = (double)d_in[v];
  } else if (c->result.size() != nv) { c->err = "no previous solve to continue from"; return MADGPU_ESTATE; }  // d_in == NULL: the fp64 result of the previous solve
  mo_hier* H = mo_create(3, c->p.size, c->p.spacing, c->p.time_step, c->tensor.data(), c->p.smoother, c->p.omega, c->p.iterations_per_grid, 0);
  if (!H) { c->err = "mo_create failed"; return MADGPU_ECUDA; }
  std::vector<int> cyc(c->p.number_of_steps > 0 ? c->p.number_of_steps : 1);
  const int rc = mo_solve(H, c->p.cycle, c->p.tolerance, c->p.max_cycles, c->p.number_of_steps, c->result.data(), cyc.data(), nullptr, 0);
  mo_destroy(H);
  if (rc) { c->err = "mo_solve failed"; return MADGPU_ECUDA; }
  for (size_t v = 0; v < nv; ++v) d_out[v] = (float)c->result[v];
  ++c->solves;
  if (stats) {
    stats->steps = c->p.number_of_steps;
    stats->total_cycles = 0;
    for (int s = 0; s < c->p.number_of_steps && s < MADGPU_MAX_STEPS; ++s) { stats->cycles_per_step[s] = cyc[s]; stats->total_cycles += cyc[s]; }
    stats->kernel_launches = 1000;  // recognisable in the front-end's launch count
  }
  return 0;
}

int madgpu_fetch_output(madgpu_ctx* c, int32_t out_type, void* out)
{
  const size_t nv = c->result.size();
  if (!nv) { c->err = "no solve yet"; return MADGPU_ESTATE; }
  for (size_t v = 0; v < nv; ++v) {
    const double x = c->result[v];
    switch (out_type) {  // static_cast< OutputPixelType >
      case MADGPU_PIX_U8: ((uint8_t*)out)[v] = (uint8_t)x; break;
      case MADGPU_PIX_I16: ((int16_t*)out)[v] = (int16_t)x; break;
      case MADGPU_PIX_F32: ((float*)out)[v] = (float)x; break;
      default: ((double*)out)[v] = x;
    }
  }
  return 0;
}

}  // extern "C"
