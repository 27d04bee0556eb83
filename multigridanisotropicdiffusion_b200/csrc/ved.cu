// ved.cu -- VED tensor front-end of libmadgpu.so (C-ABI in include/madved.h): Hessian at several scales by recursive
// Gaussian filtering, per-voxel eigen-system + vesselness + arg-max over scales, diffusion tensor, and the whole
// VEDMultigridImageFilter::GenerateData loop on the device.  sm_100a only, no CPU path.
//
// Reference: /root/reference/include/itkVEDMultigridImageFilter.hxx (cited per function).  The arithmetic lives in
// ved_math.h (shared with the CPU test harness); this file is data movement and launch geometry.
//
// HBM layout: dense fp32 volumes, x fastest (no pitch: the recursive filters walk whole lines, and the solver re-pitches the
// tensor when it ingests it).  Per context: image, 12 work volumes (3 x-pass outputs G0x G1x G2x, 6 xy products, 3 more so that
// the 6 Hessian planes never alias a z-pass input), fp64 response, 6 tensor planes: 19 fp32 + 1 fp64 volumes = 84 B / voxel
// (11.3 GB at 512^3).
//
// Kernels (all HBM-bound streaming; nothing here is a contraction):
//   k_rg_rows<K>   recursive Gaussian along x: a warp owns 32 rows and walks them in 32-column tiles transposed through shared
//                  memory, so global accesses are coalesced while each lane runs the recursion of its own row
//   k_rg_lines<K>  recursive Gaussian along y or z: one thread per line, neighbouring threads = neighbouring x (coalesced)
//                  K outputs (filter orders) per input read.  Causal pass writes, anticausal pass adds: 2 + 3K accesses / voxel.
//   k_ved_update   eigen + vesselness + tensor per voxel (fp64 arithmetic, ved::update_voxel)
//   k_cast_in/out  pixel casts of GenerateData (:70-100, :141)
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/madved.h"
#include "ved_kernels.cuh"

namespace
{
std::string g_ved_create_error;

}  // namespace

struct madved_ctx {
  madved_params p;
  long long nvox;
  cudaStream_t stream;
  cudaEvent_t ev_a, ev_b;
  float* image;
  float* work[12];
  const float* H[6];  // Hessian planes of the last madved_hessian (aliases into work[])
  float* T[6];
  double* response;
  void* stage;  // host<->device staging, grown on demand
  size_t stage_bytes;
  long long stage_voxels;  // voxels per staging chunk of the AoS fp64 paths (MADVED_STAGE_VOXELS, default 2 Mi = 96 MB)
  bool first, have_image, have_hessian, have_tensor;
  madved_stats st;
  std::string err;
};

namespace
{
int vfail(madved_ctx* c, int code, const char* fmt, ...)
{
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  else g_ved_create_error = buf;
  return code;
}

#define VCU(call)                                                                                                  \
  do {                                                                                                             \
    cudaError_t e_ = (call);                                                                                       \
    if (e_ != cudaSuccess)                                                                                         \
      return vfail(ctx, e_ == cudaErrorMemoryAllocation ? MADGPU_ENOMEM : MADGPU_ECUDA, "%s:%d %s: %s", __FILE__,  \
                   __LINE__, #call, cudaGetErrorString(e_));                                                       \
  } while (0)

vedk::Volume volume_of(const madved_ctx* ctx)
{
  const madved_params& p = ctx->p;
  return vedk::Volume{p.size[0], p.size[1], p.size[2], {p.spacing[0], p.spacing[1], p.spacing[2]}};
}

size_t vpix_size(int t) { return t == MADGPU_PIX_U8 ? 1 : t == MADGPU_PIX_I16 ? 2 : t == MADGPU_PIX_F32 ? 4 : 8; }

int ensure_vstage(madved_ctx* ctx, size_t bytes)
{
  if (ctx->stage_bytes >= bytes) return 0;
  if (ctx->stage) { VCU(cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->stage); ctx->stage = nullptr; ctx->stage_bytes = 0; }
  VCU(cudaMalloc(&ctx->stage, bytes));
  ctx->stage_bytes = bytes;
  return 0;
}

// device time of what was enqueued between the two calls, added to *acc
int timer_begin(madved_ctx* ctx)
{
  VCU(cudaEventRecord(ctx->ev_a, ctx->stream));
  return 0;
}
int timer_end(madved_ctx* ctx, double* acc)
{
  VCU(cudaEventRecord(ctx->ev_b, ctx->stream));
  VCU(cudaEventSynchronize(ctx->ev_b));
  float ms = 0.f;
  VCU(cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b));
  *acc += ms;
  VCU(cudaGetLastError());
  return 0;
}

// ComputeHessian, itkVEDMultigridImageFilter.hxx:158-173 (HessianRecursiveGaussianImageFilter, NormalizeAcrossScale on).
// Separable: H_ab = (d_a d_b G) * I.  The x pass yields G0x, G1x, G2x of the image in one read; the y pass the six xy products;
// the z pass the six components, scaled by 1 / (h_a h_b).
int hessian(madved_ctx* ctx, double sigma)
{
  if (!ctx->have_image) return vfail(ctx, MADGPU_ESTATE, "no image (madved_set_image first)");
  if (!(sigma > 0.0)) return vfail(ctx, MADGPU_EINVAL, "sigma must be positive");
  int rc = timer_begin(ctx);
  if (rc) return rc;
  ctx->st.kernel_launches += vedk::hessian_passes(ctx->stream, volume_of(ctx), sigma, ctx->image, ctx->work, ctx->H);
  rc = timer_end(ctx, &ctx->st.hessian_ms);
  if (rc) return rc;
  ctx->have_hessian = true;
  return 0;
}

int update_from_planes(madved_ctx* ctx)
{
  if (!ctx->have_hessian) return vfail(ctx, MADGPU_ESTATE, "no Hessian (madved_hessian first)");
  const ved::Params P = {ctx->p.alpha, ctx->p.beta, ctx->p.gamma, ctx->p.epsilon, ctx->p.omega, ctx->p.sensitivity};
  int rc = timer_begin(ctx);
  if (rc) return rc;
  vedk::launch_update_planes(ctx->stream, ctx->nvox, ctx->H, ctx->first, P, ctx->response, ctx->T);
  ctx->st.kernel_launches++;
  rc = timer_end(ctx, &ctx->st.vesselness_ms);
  if (rc) return rc;
  ctx->first = false;
  ctx->have_tensor = true;
  ctx->st.scales++;
  return 0;
}

int upload_image(madved_ctx* ctx, int type, const void* host)
{
  const size_t bytes = (size_t)ctx->nvox * vpix_size(type);
  if (type == MADGPU_PIX_F32) {
    VCU(cudaMemcpyAsync(ctx->image, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  } else {
    int rc = ensure_vstage(ctx, bytes);
    if (rc) return rc;
    VCU(cudaMemcpyAsync(ctx->stage, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    switch (type) {
      case MADGPU_PIX_U8: vedk::launch_cast_in(ctx->stream, (const uint8_t*)ctx->stage, ctx->image, ctx->nvox); break;
      case MADGPU_PIX_I16: vedk::launch_cast_in(ctx->stream, (const int16_t*)ctx->stage, ctx->image, ctx->nvox); break;
      case MADGPU_PIX_F64: vedk::launch_cast_in(ctx->stream, (const double*)ctx->stage, ctx->image, ctx->nvox); break;
      default: return vfail(ctx, MADGPU_EINVAL, "bad pixel type %d", type);
    }
    ctx->st.kernel_launches++;
  }
  VCU(cudaStreamSynchronize(ctx->stream));
  VCU(cudaGetLastError());
  ctx->have_image = true;
  ctx->have_hessian = false;
  return 0;
}

void begin(madved_ctx* ctx)
{
  ctx->first = true;
  ctx->have_tensor = false;
  memset(&ctx->st, 0, sizeof ctx->st);
  ctx->st.struct_size = (int32_t)sizeof(madved_stats);
}
}  // namespace

// =====================================================================================================================
//                                                       C-ABI
// =====================================================================================================================
extern "C" {

void madved_params_default(madved_params* p)
{
  memset(p, 0, sizeof *p);
  p->struct_size = (int32_t)sizeof *p;
  p->spacing[0] = p->spacing[1] = p->spacing[2] = 1.0;
  p->alpha = 0.5;         // itkVEDMultigridImageFilter.hxx:36
  p->beta = 0.5;          // :37
  p->gamma = 5.0;         // :38
  p->epsilon = 0.01;      // :39
  p->omega = 5.0;         // :40
  p->sensitivity = 10.0;  // :41
  p->device = 0;
}

const char* madved_last_error(const madved_ctx* ctx) { return ctx ? ctx->err.c_str() : g_ved_create_error.c_str(); }

void madved_destroy(madved_ctx* ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->p.device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  cudaFree(ctx->image);
  for (float* w : ctx->work) cudaFree(w);
  for (float* t : ctx->T) cudaFree(t);
  cudaFree(ctx->response);
  cudaFree(ctx->stage);
  if (ctx->ev_a) cudaEventDestroy(ctx->ev_a);
  if (ctx->ev_b) cudaEventDestroy(ctx->ev_b);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int madved_create(const madved_params* p, madved_ctx** out)
{
  madved_ctx* ctx = nullptr;  // errors before the context exists go to the create-error slot
  if (!p || !out) return vfail(ctx, MADGPU_EINVAL, "null argument");
  *out = nullptr;
  if (p->struct_size != (int32_t)sizeof(madved_params)) return vfail(ctx, MADGPU_EINVAL, "madved_params.struct_size %d != %zu", p->struct_size, sizeof(madved_params));
  for (int d = 0; d < 3; ++d) {
    if (p->size[d] < 4) return vfail(ctx, MADGPU_EINVAL, "size[%d] = %d: the recursive Gaussian needs lines of at least 4 samples", d, p->size[d]);
    if (!(p->spacing[d] > 0.0)) return vfail(ctx, MADGPU_EINVAL, "spacing[%d] must be positive", d);
  }
  if (!(p->sensitivity != 0.0)) return vfail(ctx, MADGPU_EINVAL, "sensitivity must not be zero");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return vfail(ctx, MADGPU_ECUDA, "no CUDA device (there is no CPU fallback)"); }
  if (p->device < 0 || p->device >= ndev) return vfail(ctx, MADGPU_EINVAL, "device %d out of range (%d devices)", p->device, ndev);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, p->device) != cudaSuccess || prop.major < 10) { cudaGetLastError(); return vfail(ctx, MADGPU_ECUDA, "device %d is not sm_100 (this library is built for B200 only)", p->device); }
  madved_ctx* c = new madved_ctx();
  c->p = *p;
  c->nvox = (long long)p->size[0] * p->size[1] * p->size[2];
  c->stream = nullptr; c->ev_a = c->ev_b = nullptr;
  c->image = nullptr; c->response = nullptr; c->stage = nullptr; c->stage_bytes = 0;
  for (auto& w : c->work) w = nullptr;
  for (auto& t : c->T) t = nullptr;
  for (auto& h : c->H) h = nullptr;
  c->have_image = c->have_hessian = false;
  c->stage_voxels = 2ll << 20;
  if (const char* e = getenv("MADVED_STAGE_VOXELS")) c->stage_voxels = std::max(1ll, atoll(e));
  begin(c);
  ctx = c;
  auto bail = [&](cudaError_t e, const char* what) {
    const int code = e == cudaErrorMemoryAllocation ? MADGPU_ENOMEM : MADGPU_ECUDA;
    g_ved_create_error = std::string(what) + ": " + cudaGetErrorString(e);
    cudaGetLastError();
    madved_destroy(c);
    return code;
  };
  cudaError_t e;
  if ((e = cudaSetDevice(p->device)) != cudaSuccess) return bail(e, "cudaSetDevice");
  if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
  if ((e = cudaEventCreate(&c->ev_a)) != cudaSuccess || (e = cudaEventCreate(&c->ev_b)) != cudaSuccess) return bail(e, "cudaEventCreate");
  const size_t fb = (size_t)c->nvox * sizeof(float);
  if ((e = cudaMalloc((void**)&c->image, fb)) != cudaSuccess) return bail(e, "cudaMalloc(image)");
  for (auto& w : c->work)
    if ((e = cudaMalloc((void**)&w, fb)) != cudaSuccess) return bail(e, "cudaMalloc(work volume)");
  for (auto& t : c->T)
    if ((e = cudaMalloc((void**)&t, fb)) != cudaSuccess) return bail(e, "cudaMalloc(tensor plane)");
  if ((e = cudaMalloc((void**)&c->response, (size_t)c->nvox * sizeof(double))) != cudaSuccess) return bail(e, "cudaMalloc(response)");
  *out = c;
  return 0;
}

int madved_set_params(madved_ctx* ctx, double alpha, double beta, double gamma, double epsilon, double omega, double sensitivity)
{
  if (!ctx) return MADGPU_EINVAL;
  if (!(sensitivity != 0.0)) return vfail(ctx, MADGPU_EINVAL, "sensitivity must not be zero");
  ctx->p.alpha = alpha; ctx->p.beta = beta; ctx->p.gamma = gamma;
  ctx->p.epsilon = epsilon; ctx->p.omega = omega; ctx->p.sensitivity = sensitivity;
  return 0;
}

int madved_set_image(madved_ctx* ctx, int32_t type, const void* image)
{
  if (!ctx) return MADGPU_EINVAL;
  if (!image) return vfail(ctx, MADGPU_EINVAL, "null image pointer");
  if (type < 0 || type > 3) return vfail(ctx, MADGPU_EINVAL, "bad pixel type %d", type);
  VCU(cudaSetDevice(ctx->p.device));
  return upload_image(ctx, type, image);
}

int madved_set_image_device_f32(madved_ctx* ctx, const float* d_image)
{
  if (!ctx) return MADGPU_EINVAL;
  if (!d_image) return vfail(ctx, MADGPU_EINVAL, "null image pointer");
  VCU(cudaSetDevice(ctx->p.device));
  if (d_image != ctx->image) VCU(cudaMemcpyAsync(ctx->image, d_image, (size_t)ctx->nvox * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
  VCU(cudaStreamSynchronize(ctx->stream));
  ctx->have_image = true;
  ctx->have_hessian = false;
  return 0;
}

int madved_image_device(madved_ctx* ctx, float** d_image)
{
  if (!ctx || !d_image) return MADGPU_EINVAL;
  *d_image = ctx->image;
  return 0;
}

int madved_begin(madved_ctx* ctx)
{
  if (!ctx) return MADGPU_EINVAL;
  begin(ctx);
  return 0;
}

int madved_hessian(madved_ctx* ctx, double sigma)
{
  if (!ctx) return MADGPU_EINVAL;
  VCU(cudaSetDevice(ctx->p.device));
  return hessian(ctx, sigma);
}

int madved_update_vesselness(madved_ctx* ctx)
{
  if (!ctx) return MADGPU_EINVAL;
  VCU(cudaSetDevice(ctx->p.device));
  return update_from_planes(ctx);
}

int madved_add_scale(madved_ctx* ctx, double sigma)
{
  int rc = madved_hessian(ctx, sigma);
  if (rc) return rc;
  return update_from_planes(ctx);
}

int madved_update_vesselness_host_f64(madved_ctx* ctx, const double* hessian_aos)
{
  if (!ctx) return MADGPU_EINVAL;
  if (!hessian_aos) return vfail(ctx, MADGPU_EINVAL, "null Hessian pointer");
  VCU(cudaSetDevice(ctx->p.device));
  const long long chunk = std::min<long long>(ctx->nvox, ctx->stage_voxels);
  int rc = ensure_vstage(ctx, (size_t)chunk * 6 * sizeof(double));
  if (rc) return rc;
  const ved::Params P = {ctx->p.alpha, ctx->p.beta, ctx->p.gamma, ctx->p.epsilon, ctx->p.omega, ctx->p.sensitivity};
  for (long long v0 = 0; v0 < ctx->nvox; v0 += chunk) {
    const long long cnt = std::min(chunk, ctx->nvox - v0);
    VCU(cudaMemcpyAsync(ctx->stage, hessian_aos + v0 * 6, (size_t)cnt * 6 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    vedk::launch_update_aos(ctx->stream, v0, cnt, (const double*)ctx->stage, ctx->first, P, ctx->response, ctx->T);
    ctx->st.kernel_launches++;
    VCU(cudaStreamSynchronize(ctx->stream));  // the single staging buffer is reused by the next chunk
  }
  VCU(cudaGetLastError());
  ctx->first = false;
  ctx->have_tensor = true;
  ctx->st.scales++;
  return 0;
}

int madved_tensor_planes(madved_ctx* ctx, const float** planes)
{
  if (!ctx || !planes) return MADGPU_EINVAL;
  if (!ctx->have_tensor) return vfail(ctx, MADGPU_ESTATE, "no tensor yet (no Hessian consumed since madved_begin)");
  for (int k = 0; k < 6; ++k) planes[k] = ctx->T[k];
  return 0;
}

static int planes_to_host_aos(madved_ctx* ctx, const float* const* planes, double* out)
{
  const long long chunk = std::min<long long>(ctx->nvox, ctx->stage_voxels);
  int rc = ensure_vstage(ctx, (size_t)chunk * 6 * sizeof(double));
  if (rc) return rc;
  for (long long v0 = 0; v0 < ctx->nvox; v0 += chunk) {
    const long long cnt = std::min(chunk, ctx->nvox - v0);
    vedk::launch_planes_to_aos(ctx->stream, planes, (double*)ctx->stage, v0, cnt);
    VCU(cudaMemcpyAsync(out + v0 * 6, ctx->stage, (size_t)cnt * 6 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    VCU(cudaStreamSynchronize(ctx->stream));
  }
  VCU(cudaGetLastError());
  return 0;
}

int madved_get_tensor_f64(madved_ctx* ctx, double* tensor_aos)
{
  if (!ctx) return MADGPU_EINVAL;
  if (!tensor_aos) return vfail(ctx, MADGPU_EINVAL, "null output pointer");
  if (!ctx->have_tensor) return vfail(ctx, MADGPU_ESTATE, "no tensor yet (no Hessian consumed since madved_begin)");
  VCU(cudaSetDevice(ctx->p.device));
  return planes_to_host_aos(ctx, ctx->T, tensor_aos);
}

int madved_get_hessian_f64(madved_ctx* ctx, double* hessian_aos)
{
  if (!ctx) return MADGPU_EINVAL;
  if (!hessian_aos) return vfail(ctx, MADGPU_EINVAL, "null output pointer");
  if (!ctx->have_hessian) return vfail(ctx, MADGPU_ESTATE, "no Hessian (madved_hessian first)");
  VCU(cudaSetDevice(ctx->p.device));
  return planes_to_host_aos(ctx, ctx->H, hessian_aos);
}

int madved_get_response_f64(madved_ctx* ctx, double* response)
{
  if (!ctx) return MADGPU_EINVAL;
  if (!response) return vfail(ctx, MADGPU_EINVAL, "null output pointer");
  if (!ctx->have_tensor) return vfail(ctx, MADGPU_ESTATE, "no vesselness yet (no Hessian consumed since madved_begin)");
  VCU(cudaSetDevice(ctx->p.device));
  VCU(cudaMemcpyAsync(response, ctx->response, (size_t)ctx->nvox * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  VCU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int madved_get_stats(madved_ctx* ctx, madved_stats* stats)
{
  if (!ctx || !stats) return MADGPU_EINVAL;
  const size_t n = std::min((size_t)(stats->struct_size > 0 ? stats->struct_size : (int32_t)sizeof(madved_stats)), sizeof(madved_stats));
  memcpy(stats, &ctx->st, n);
  return 0;
}

// VEDMultigridImageFilter::GenerateData, itkVEDMultigridImageFilter.hxx:63-155.
int madved_run(madved_ctx* ctx, madgpu_ctx* solver, int32_t in_type, const void* in, int32_t out_type, void* out, const double* scales,
               int32_t nscales, int32_t iterations, madgpu_stats* solver_stats)
{
  if (!ctx) return MADGPU_EINVAL;
  if (!solver || !in || !out || !scales) return vfail(ctx, MADGPU_EINVAL, "null argument");
  if (nscales < 1 || iterations < 0) return vfail(ctx, MADGPU_EINVAL, "need at least one scale and a non-negative iteration count");
  if (in_type < 0 || in_type > 3 || out_type < 0 || out_type > 3) return vfail(ctx, MADGPU_EINVAL, "bad pixel type");
  int32_t sz[3];
  double sp[3];
  int32_t cent[3];
  if (madgpu_level_info(solver, 0, sz, sp, cent) != 0 || sz[0] != ctx->p.size[0] || sz[1] != ctx->p.size[1] || sz[2] != ctx->p.size[2])
    return vfail(ctx, MADGPU_EINVAL, "the solver context has a different volume size");
  VCU(cudaSetDevice(ctx->p.device));
  auto t0 = std::chrono::steady_clock::now();
  int rc = upload_image(ctx, in_type, in);  // :70-100
  if (rc) return rc;
  const double h2d = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  double hessian_ms = 0, vesselness_ms = 0, diffusion_ms = 0;
  int64_t launches = ctx->st.kernel_launches;
  int32_t nsc = 0;
  madgpu_stats sst;
  memset(&sst, 0, sizeof sst);
  sst.struct_size = (int32_t)sizeof sst;
  for (int it = 0; it < iterations; ++it) {  // :105
    begin(ctx);                              // the state of the previous outer iteration was dropped at :121-123
    for (int s = 0; s < nscales; ++s) {      // :108-117
      rc = hessian(ctx, scales[s]);
      if (rc) return rc;
      rc = update_from_planes(ctx);
      if (rc) return rc;
    }
    hessian_ms += ctx->st.hessian_ms; vesselness_ms += ctx->st.vesselness_ms; launches += ctx->st.kernel_launches; nsc += ctx->st.scales;
    // DiffusionStep, :381-402: the tensor goes to the solver in HBM, the image is solved in place
    t0 = std::chrono::steady_clock::now();
    const float* planes[6];
    for (int k = 0; k < 6; ++k) planes[k] = ctx->T[k];
    VCU(cudaStreamSynchronize(ctx->stream));
    if (madgpu_set_tensor_device_f32(solver, planes) != 0) return vfail(ctx, MADGPU_ECUDA, "solver: %s", madgpu_last_error(solver));
    // the first DiffusionStep starts from the (fp32) input image, the following ones from the solver's own fp64 result; the fp32 copy that
    // comes back feeds the next Hessian
    if (madgpu_solve_device_f32(solver, it == 0 ? ctx->image : nullptr, ctx->image, &sst) != 0) return vfail(ctx, MADGPU_ECUDA, "solver: %s", madgpu_last_error(solver));
    ctx->have_hessian = false;  // the image changed
    launches += sst.kernel_launches;
    diffusion_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }
  t0 = std::chrono::steady_clock::now();
  if (iterations > 0) {
    // the result is cast from the solver's fp64 iterate straight to the output pixel type (static_cast, :141)
    if (madgpu_fetch_output(solver, out_type, out) != 0) return vfail(ctx, MADGPU_ECUDA, "solver: %s", madgpu_last_error(solver));
  } else {
    // no iteration: the output is the cast input (:141 on the untouched internal image)
    if (in_type == out_type) memcpy(out, in, (size_t)ctx->nvox * vpix_size(in_type));
    else return vfail(ctx, MADGPU_EINVAL, "iterations == 0 needs in_type == out_type");
  }
  ctx->st.d2h_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  ctx->st.h2d_ms = h2d;
  ctx->st.hessian_ms = hessian_ms; ctx->st.vesselness_ms = vesselness_ms; ctx->st.diffusion_ms = diffusion_ms;
  ctx->st.kernel_launches = launches; ctx->st.scales = nsc;
  if (solver_stats) {
    const size_t n = std::min((size_t)(solver_stats->struct_size > 0 ? solver_stats->struct_size : (int32_t)sizeof(madgpu_stats)), sizeof(madgpu_stats));
    memcpy(solver_stats, &sst, n);
  }
  return 0;
}

}  // extern "C"
