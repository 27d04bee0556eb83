// madgpu.cu -- host driver and C-ABI of libmadgpu.so (see include/madgpu.h).
//
// Replaces, on one B200, the solve path of itk::MultigridAnisotropicDiffusionImageFilter:
//   hierarchy        mad/itkGridsHierarchy.hxx:30-204
//   V-cycle          itkMultigridAnisotropicDiffusionImageFilter.hxx:341-493
//   FMG              itkMultigridAnisotropicDiffusionImageFilter.hxx:300-338
//   time-step loop   itkMultigridAnisotropicDiffusionImageFilter.hxx:104-297
// (paths relative to /root/reference/include).
//
// Precision plan: all multigrid work is fp32; the level-0 iterate u and right-hand side f are
// kept in fp64 and every cycle is applied in defect-correction form
//      r = f - A u (fp64 arithmetic)  ->  e = Cycle(0, r) (fp32)  ->  u += e,
// which is algebraically the reference's u = VCycle(u, f) (both smoothers and the transfers are
// linear) but lets the relative residual reach the 1e-10 the reference tests ask for.  The fp64
// residual doubles as the stop test of the previous cycle, so it costs one pass, not two.
#include "../../include/madgpu.h"
#include "mad_kernels.cuh"
#include "mad_fast.cuh"
#include "mad_fast2d.cuh"

// Every kernel launch of this file goes through MAD_LAUNCH((kernel<...>), grid, block, shared bytes, stream, args...) -- the kernel
// name in parentheses so that template commas survive the preprocessor.  Here it is the plain <<<>>> launch; the CPU test build
// (tests/mad_host/) defines it beforehand to run the same kernel source on host fibres.
#ifndef MAD_LAUNCH
#define MAD_UNPAREN(...) __VA_ARGS__
#define MAD_LAUNCH(kernel, grid, block, smem, stream, ...) MAD_UNPAREN kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <string>
#include <vector>

using namespace mad;

namespace {

thread_local std::string g_create_error;

struct Level {
  Geom g;
  int n[3];
  double h[3];
  int cent[3];
  int gnz;       // global number of planes of this level (== n[2] unless this context is one z-slab of it)
  int zb;        // global z of local plane 0
  size_t elems;  // allocation size in elements (incl. 2 ghost planes)
  float *u, *f, *tmp;  // plane-0 pointers
  float* D[6];
  uint4* coef16;      // packed fp16 operator rows for the Gauss-Seidel smoother (mad_fast.cuh), built on first use
  bool coef16_valid;
  int tb_flip;        // temporal blocking: the next fused pass uses the tile grid shifted by half a tile in y and z (alternates per pass)
  bool coef16_off;    // some diagonal of this level is beyond the range of the packed rows (fast::COEF_DIAG_MAX): exact-row sweeps instead
  std::vector<void*> allocs;
};

// NCCL is bound at run time (dlopen): libmadgpu.so has no load-time dependency on it and single-GPU users never
// touch it.  Only the handful of entry points the halo exchange needs are declared (ABI of nccl.h 2.x).
struct Nccl {
  typedef struct ncclComm* comm_t;
  struct UniqueId { char internal[128]; };
  enum { Float32 = 7, Float64 = 8, Sum = 0 };
  void* lib;
  int (*GetUniqueId)(UniqueId*);
  int (*CommInitRank)(comm_t*, int, UniqueId, int);
  int (*CommDestroy)(comm_t);
  int (*Send)(const void*, size_t, int, int, comm_t, cudaStream_t);
  int (*Recv)(void*, size_t, int, int, comm_t, cudaStream_t);
  int (*AllReduce)(const void*, void*, size_t, int, int, comm_t, cudaStream_t);
  int (*Broadcast)(const void*, void*, size_t, int, int, comm_t, cudaStream_t);
  int (*GroupStart)();
  int (*GroupEnd)();
  const char* (*GetErrorString)(int);
};

struct ProfEvent {
  int cls;
  cudaEvent_t e0, e1;
};

}  // namespace

struct madgpu_ctx {
  madgpu_params p;
  int dim, ncomp, nlevels;
  Level lv[MADGPU_MAX_LEVELS];
  double *u64, *f64;  // level-0 outer fields (plane-0 pointers)
  std::vector<void*> allocs;
  bool have_result;  // run_steps has completed on this context: f64 holds its fp64 result
  double* Ainv;  // device, ncoarse^2
  double* gjM;   // device work matrix [A | I] of the Gauss-Jordan inverse (+ saved column, singular flag), kept between tensors
  int gj_n;
  cudaGraphExec_t gj_exec;  // the inverse's launches, captured once
  bool gj_graph_bad;
  int ncoarse;
  bool coarse_direct;
  double* partials;
  size_t npartials;
  double* d_scalar;
  double* h_scalar;  // pinned
  cudaStream_t stream;
  cudaEvent_t ev_a, ev_b, ev_c;
  bool tensor_set;
  std::string err;
  std::vector<double> relres_hist;
  madgpu_stats st;
  int profiling;  // bit mask of kernel classes
  double rhs_norm;  // of the current cycles_begin
  std::vector<ProfEvent> prof;
  std::vector<cudaEvent_t> ev_pool;
  int64_t launches;
  // z-slab decomposition (world > 1): this context owns planes [lv[l].zb, lv[l].zb + lv[l].n[2]) of levels 0..nlevels-1;
  // the last of them (the agglomeration level) is gathered to rank 0, whose `sub` context holds the rest of the hierarchy
  void* stage[2];        // persistent host<->device staging (tensor chunks, image), grown on demand
  size_t stage_bytes[2];
  // peer-memory halo (z-slabs with CUDA IPC): every shareable allocation of this context, the neighbours' mappings of
  // the same allocations, two arrival counters the neighbours write with stream memory operations
  struct Shared { void* local; size_t bytes; void* lo; void* hi; uint32_t produced; };
  std::vector<Shared> shared;
  uint32_t* flags;        // [0] written by the lower neighbour, [1] by the upper one
  uint32_t* flags_lo;     // the lower neighbour's flags (mapped), we write [1]
  uint32_t* flags_hi;     // the upper neighbour's flags (mapped), we write [0]
  uint32_t halo_seq;      // number of fields produced so far (identical on every rank: same program order)
  bool p2p;
  int p2p_wait_kernel;          // arrival counters are awaited by k_halo_wait (bounded; default) or, MADGPU_P2P_WAIT=memop, by cuStreamWaitValue32
  long long p2p_timeout_cycles; // MADGPU_P2P_TIMEOUT_MS (default 10 s at ~2 GHz)
  int p2p_drop_rank, p2p_drop_seq;  // test hook MADGPU_P2P_TEST_DROP_SIGNAL=rank:seq: from that sequence number on the rank's signals are not sent
  int rank, world;
  std::string sticky;  // first collective error inside an operator
  bool borrowed_stream;  // sub-context of a slab context: runs on the parent's stream
  Nccl::comm_t comm;
  madgpu_ctx* sub;
  float* gather_buf;   // rank 0: dense agglomeration-level field (all slabs)
  float* slab_buf;     // every rank: dense local slab of the agglomeration level, +2 ghost planes
  int total_levels;    // levels of the whole hierarchy (distributed + serial)
  int gsize[MADGPU_MAX_LEVELS][3], gcent[MADGPU_MAX_LEVELS][3];  // global level schedule
  int pf_dist;      // L2 prefetch distance (planes) of the streaming kernels, 0 = off
  int gs_pairs;     // packed-row Gauss-Seidel: one warp per row pair (k_coef_gs2) where ny is even
  int gs_coef16;    // fused Gauss-Seidel reads pre-evaluated fp16 operator rows (default) instead of the tensor planes
  int gs_fused;     // 3-D Gauss-Seidel as one fused pass per sweep (default) instead of one pass per colour
  int fast_cfg;     // CTA shape / register cap of the streaming kernels (tuning hook)
  int res64_smem;   // MADGPU_RES64_SMEM: 0 (default) = all-register fp64 residual (k_fast_sweep<MODE_RES_C32>), 3 / 4 = k_fast_res64 capped for 3 / 4 CTAs per SM
                    // (measured on B200 at 512^3: 1.66 ms all-register, 2.16 / 3.44 ms with the shared-memory ring -- kept as an A/B hook)
  int res64_minb;   // MADGPU_RES64_MINB=3: register cap of the fp64 residual for 3 CTAs per SM
  int res64_c32;    // level-0 fp64 residual with the operator row evaluated in fp32 -- the row the fp32 sweeps relax -- and applied in fp64 (default; MADGPU_RES64_COEF32=0: fp64 row)
  int fast_min_nx;  // 3-D levels with nx >= this use the streaming kernels of mad_fast.cuh
  int fast2d;       // MADGPU_FAST2D=0: 2-D levels keep the generic one-pixel-per-thread kernels (A/B and cross-check hook)
  long long fast2d_min_pixels;  // MADGPU_FAST2D_MIN_PIXELS (default 2^20): see use_fast2
  // CUDA graphs of the launch-bound part of a V-cycle: vcycle(l) for the first level of at most graph_voxels voxels (and with it
  // everything below) is captured once per (level, zero guess, solver settings, ping-pong state) and replayed
  struct CycleGraph {
    int level, zero_guess, smoother, nu;
    double omega;
    std::vector<const void*> state;  // u pointers of levels >= level at entry (the sweeps swap u / tmp)
    cudaGraphExec_t exec;
    int launches, seen;
    bool bad;
  };
  std::vector<CycleGraph> graphs;
  long long graph_voxels;  // MADGPU_GRAPH_VOXELS (0 = no graphs)
  int gs_private;          // MADGPU_GS_PRIVATE: 2 (default) = the row-pair sweep with warp-private tiles (128 x 2 x zc, no CTA barriers) on a grid of
                           // pairs that alternates between even and odd first rows from sweep to sweep; 1 = the same on a fixed grid; 0 = tiles of
                           // 128 x 8 x zc shared by the four warps of a CTA (two barriers per plane step; the default until call q of round 2)
  int gs_tb_single;        // MADGPU_GS_TB_SINGLE=1: every sweep through k_coef_gs_tb<1> (shared-memory ring fed by cp.async, tiles of 128 x 16) instead of k_coef_gs2: A/B hook
  int gs_tb;               // temporal blocking of the Gauss-Seidel sweeps of a leg: up to this many sweeps per pass (MADGPU_GS_TB = 2 or 3; default 1 = off:
                           // measured on B200 at 512^3 a fused pass of 3 sweeps takes 2.04 ms against 3 x 0.77 ms -- the packed rows of the older planes
                           // come from L2, not HBM, but still cross the L2 -> SM fabric once per sweep -- and its frozen tile faces cost 3 V-cycles in 6)
  int coarse_host;         // MADGPU_COARSE_HOST=1: assemble and invert the coarsest operator on the host (round-1 path; cross-check)
  long long coarse_direct_max;  // coarsest grids of up to this many unknowns get the dense inverse (MADGPU_COARSE_DIRECT_MAX, default 4096)
  int prolong_cell;        // MADGPU_PROLONG_CELL=0: keep the generic streaming prolongation for cell-centred transfers too (A/B hook)
  int restrict_cell;       // MADGPU_RESTRICT_CELL=0: likewise for the restriction
  bool capturing;
};

namespace {

int fail(madgpu_ctx* c, int code, const char* fmt, ...)
{
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  else g_create_error = buf;
  return code;
}

#define CU(call)                                                                                                   \
  do {                                                                                                             \
    cudaError_t e_ = (call);                                                                                       \
    if (e_ != cudaSuccess)                                                                                         \
      return fail(ctx, e_ == cudaErrorMemoryAllocation ? MADGPU_ENOMEM : MADGPU_ECUDA, "%s:%d %s: %s", __FILE__,   \
                  __LINE__, #call, cudaGetErrorString(e_));                                                        \
  } while (0)

template <typename T>
int dalloc(madgpu_ctx* ctx, std::vector<void*>& owner, T** out, size_t n)
{
  void* p = nullptr;
  CU(cudaMalloc(&p, n * sizeof(T)));
  CU(cudaMemsetAsync(p, 0, n * sizeof(T), ctx->stream));
  owner.push_back(p);
  *out = (T*)p;
  return 0;
}

dim3 block3(int dim) { return dim == 3 ? dim3(32, 4, 4) : dim3(32, 16, 1); }
dim3 grid3(const Geom& g, dim3 b) { return dim3((g.nx + b.x - 1) / b.x, (g.ny + b.y - 1) / b.y, (g.nz + b.z - 1) / b.z); }

Tensor tensor_of(const Level& L);
void set_ghosts(madgpu_ctx* ctx, Geom& g, const Level& L, const void* field, size_t es);
void halo_signal(madgpu_ctx* ctx, const void* field);
void halo_dirty(madgpu_ctx* ctx, const void* field);

// ---- streaming (mad_fast.cuh) launch geometry -------------------------------------------------
bool use_fast(const madgpu_ctx* ctx, const Level& L)
{
  return ctx->dim == 3 && L.g.nx >= ctx->fast_min_nx && L.g.nz >= 3 && L.g.ny >= 3 && L.elems < (1ull << 31);  // 32-bit element offsets
}
bool gs_pairs(const madgpu_ctx* ctx, const Level& L) { return ctx->gs_pairs && L.g.ny % 2 == 0 && L.g.ny >= 4; }
// planes per CTA: enough CTAs for ~8 waves of the resident set, at least 8 planes so that the two start-up planes stay cheap
int fast_zc(const Geom& g, int wy)
{
  const long long cxy = (long long)((g.nx + fast::TX - 1) / fast::TX) * ((g.ny + wy - 1) / wy);
  const long long want = 148ll * 3 * 8;
  long long chunks = std::max(1ll, want / std::max(1ll, cxy));
  int zc = (int)std::max(8ll, (g.nz + chunks - 1) / chunks);
  return std::min(zc, std::max(g.nz, 1));
}
dim3 fast_grid(const Geom& g, int wy, int zc) { return dim3((g.nx + fast::TX - 1) / fast::TX, (g.ny + wy - 1) / wy, (g.nz + zc - 1) / zc); }

// ---- 2-D streaming kernels (mad_fast2d.cuh): one warp per strip of 128 columns and chunk of yc rows ----------------------------
// Gauss-Seidel: at every size (one launch per sweep instead of four colour passes).  Weighted Jacobi, residuals and transfers: only
// on levels of at least fast2d_min_pixels pixels -- a warp marches its rows one after the other, so a small level (512^2: 256
// warps) is latency-bound where the one-pixel-per-thread kernels spread it over the whole GPU (measured on B200, 512^2: a weighted-
// Jacobi V(3,3) cycle 0.34 ms with the streaming kernels everywhere against 0.21 ms; 8192^2: 4.8 ms against 10.1 ms).
bool use_fast2_gs(const madgpu_ctx* ctx, const Level& L)
{
  return ctx->dim == 2 && ctx->fast2d && L.g.nx >= ctx->fast_min_nx && L.g.ny >= 4 && L.elems < (1ull << 31);
}
bool use_fast2(const madgpu_ctx* ctx, const Level& L)
{
  return use_fast2_gs(ctx, L) && (long long)L.g.nx * L.g.ny >= ctx->fast2d_min_pixels;
}
constexpr int F2_WY = 4;  // warps per CTA (independent of each other)
// rows per warp: enough warps for ~8 per SM sub-partition on large images, at least 8 rows so that the two start-up rows stay cheap
int fast2_yc(const Geom& g)
{
  const long long strips = (g.nx + fast::TX - 1) / fast::TX;
  const long long chunks = std::max(1ll, (148ll * 32) / std::max(1ll, strips));
  const int yc = (int)std::max(8ll, (g.ny + chunks - 1) / chunks);
  return std::min(yc, std::max(g.ny, 1));
}
template <int MODE, typename T, typename UT, typename FT>
size_t launch_fast2(madgpu_ctx* ctx, const Level& L, const UT* u, const FT* f, float* out, double* partials, float omega, int uzero = 0)
{
  const int yc = fast2_yc(L.g);
  const int chunks = (L.g.ny + yc - 1) / yc;
  const dim3 fg((L.g.nx + fast::TX - 1) / fast::TX, (chunks + F2_WY - 1) / F2_WY);
  MAD_LAUNCH((fast::k2_sweep<MODE, T, UT, FT, F2_WY>), fg, dim3(32, F2_WY), 0, ctx->stream, L.g, tensor_of(L), u, f, out, partials, omega, yc, uzero);
  return (size_t)fg.x * fg.y;
}

// One streaming pass (MODE_WJ / MODE_RES) over a level; returns the number of CTAs (= partial sums written).
// ctx->fast_cfg selects the CTA shape / register cap (tuning hook MADGPU_FAST_CFG).
template <int MODE, typename T, typename UT, typename FT, typename OT>
size_t launch_fast(madgpu_ctx* ctx, const Level& L, const UT* u, const FT* f, OT* out, double* partials, float omega, int uzero = 0)
{
  const Tensor D = tensor_of(L);
  Geom gg = L.g;
  if (MODE != fast::MODE_COEF) set_ghosts(ctx, gg, L, out, sizeof(OT));
#define MAD_FAST_LAUNCH(WY, MINB, PF)                                                                                        \
  do {                                                                                                                       \
    const int zc = fast_zc(L.g, WY);                                                                                         \
    const dim3 fg = fast_grid(L.g, WY, zc);                                                                                  \
    MAD_LAUNCH((fast::k_fast_sweep<MODE, T, UT, FT, OT, WY, MINB, PF>), fg, dim3(32, WY), 0, ctx->stream, gg, D, u, f, out, partials, omega, zc, ctx->pf_dist, uzero); \
    if (MODE != fast::MODE_COEF) halo_signal(ctx, out);                                                                      \
    return (size_t)fg.x * fg.y * fg.z;                                                                                       \
  } while (0)
  if (sizeof(T) == 8) {
    if (ctx->fast_cfg == 4) MAD_FAST_LAUNCH(4, 2, true);
    if (ctx->res64_minb == 3) MAD_FAST_LAUNCH(4, 3, false);  // 168 registers (spills) for 12 instead of 8 warps per SM: A/B hook MADGPU_RES64_MINB
    MAD_FAST_LAUNCH(4, 2, false);
  }
  switch (ctx->fast_cfg) {
    case 1: MAD_FAST_LAUNCH(8, 1, false);
    case 4: MAD_FAST_LAUNCH(4, 2, true);
    case 5: MAD_FAST_LAUNCH(8, 1, true);
    default: MAD_FAST_LAUNCH(4, 3, false);
  }
#undef MAD_FAST_LAUNCH
}

// ---- launch bookkeeping -------------------------------------------------------------------
cudaEvent_t get_event(madgpu_ctx* ctx)
{
  if (!ctx->ev_pool.empty()) {
    cudaEvent_t e = ctx->ev_pool.back();
    ctx->ev_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

struct Scope {
  madgpu_ctx* ctx;
  ProfEvent pe;
  bool on;
  Scope(madgpu_ctx* c, int cls, int nlaunch = 1) : ctx(c), on(((c->profiling >> cls) & 1) != 0)
  {
    ctx->launches += nlaunch;
    if (on) {
      pe.cls = cls;
      pe.e0 = get_event(ctx);
      pe.e1 = get_event(ctx);
      cudaEventRecord(pe.e0, ctx->stream);
      ctx->st.prof_launches[cls] += nlaunch;
    }
  }
  ~Scope()
  {
    if (on) {
      cudaEventRecord(pe.e1, ctx->stream);
      ctx->prof.push_back(pe);
    }
  }
};

void prof_collect(madgpu_ctx* ctx)
{
  for (auto& pe : ctx->prof) {
    float ms = 0.f;
    cudaEventSynchronize(pe.e1);
    cudaEventElapsedTime(&ms, pe.e0, pe.e1);
    ctx->st.prof_ms[pe.cls] += ms;
    ctx->ev_pool.push_back(pe.e0);
    ctx->ev_pool.push_back(pe.e1);
  }
  ctx->prof.clear();
}

Tensor tensor_of(const Level& L)
{
  Tensor t;
  for (int c = 0; c < 6; ++c) t.p[c] = L.D[c];
  return t;
}

// ---- z-slab decomposition: NCCL binding, halo exchange, agglomeration ------------------------------
Nccl g_nccl = {};

const char* nccl_load()
{
  if (g_nccl.lib) return nullptr;
  const char* names[] = {getenv("MADGPU_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};  // MADGPU_NCCL_LIB: a specific build, e.g. the one bundled with torch
  void* h = nullptr;
  for (const char* n : names)
    if (n && *n && (h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
  if (!h) return "libnccl.so.2 not found (needed only for world_size > 1)";
#define MAD_SYM(field, name)                                             \
  *(void**)(&g_nccl.field) = dlsym(h, name);                             \
  if (!g_nccl.field) return "libnccl lacks " name;
  MAD_SYM(GetUniqueId, "ncclGetUniqueId")
  MAD_SYM(CommInitRank, "ncclCommInitRank")
  MAD_SYM(CommDestroy, "ncclCommDestroy")
  MAD_SYM(Send, "ncclSend")
  MAD_SYM(Recv, "ncclRecv")
  MAD_SYM(AllReduce, "ncclAllReduce")
  MAD_SYM(Broadcast, "ncclBroadcast")
  MAD_SYM(GroupStart, "ncclGroupStart")
  MAD_SYM(GroupEnd, "ncclGroupEnd")
  MAD_SYM(GetErrorString, "ncclGetErrorString")
#undef MAD_SYM
  g_nccl.lib = h;
  return nullptr;
}

// Collective errors inside the (void) operator functions are latched here and reported by the entry point.
void latch(madgpu_ctx* ctx, int r, const char* what)
{
  if (r == 0 || !ctx->sticky.empty()) return;
  const bool nccl = strncmp(what, "g_nccl", 6) == 0;
  ctx->sticky = std::string(what) + ": " + (nccl && g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : ("error " + std::to_string(r)));
}
#define NCV(call) latch(ctx, (call), #call)

// ---- peer-memory halo: the kernel that produces a field stores its boundary planes straight into the neighbours' ghost
// ---- planes (Geom::glo / ghi, NVLink peer stores) and the stream then bumps a counter in the neighbours' memory
// ---- (cuStreamWriteValue32); a kernel that reads ghost planes first waits for the counters (cuStreamWaitValue32).  No
// ---- separate exchange operation, no host round trip.  The NCCL send/recv path remains for fields produced otherwise.
typedef int (*StreamValueFn)(cudaStream_t, unsigned long long, uint32_t, unsigned int);
StreamValueFn g_write_value = nullptr, g_wait_value = nullptr;

const char* stream_memops_load()
{
  if (g_write_value && g_wait_value) return nullptr;
  cudaDriverEntryPointQueryResult q;
  void* f = nullptr;
  if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &f, cudaEnableDefault, &q) != cudaSuccess || !f) return "cuStreamWriteValue32 unavailable";
  g_write_value = (StreamValueFn)f;
  if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &f, cudaEnableDefault, &q) != cudaSuccess || !f) return "cuStreamWaitValue32 unavailable";
  g_wait_value = (StreamValueFn)f;
  return nullptr;
}

madgpu_ctx::Shared* shared_of(madgpu_ctx* ctx, const void* field)
{
  for (auto& s : ctx->shared)
    if ((const char*)field >= (const char*)s.local && (const char*)field < (const char*)s.local + s.bytes) return &s;
  return nullptr;
}

// Ghost targets of a field this rank is about to produce (plane-0 pointer `field`, element size es).
void set_ghosts(madgpu_ctx* ctx, Geom& g, const Level& L, const void* field, size_t es)
{
  g.glo = g.ghi = nullptr;
  if (!ctx->p2p) return;
  madgpu_ctx::Shared* s = shared_of(ctx, field);
  if (!s) return;
  const size_t off = (const char*)field - (const char*)s->local;  // bytes from the allocation start to plane 0
  const size_t plane = (size_t)L.g.plane * es;
  if (s->lo) g.glo = (char*)s->lo + off + (size_t)L.g.nz * plane;  // lower neighbour's upper ghost plane
  if (s->hi) g.ghi = (char*)s->hi + off - plane;                   // upper neighbour's lower ghost plane
}

// after the producing kernel has been launched: publish "field version halo_seq is complete" to both neighbours
void halo_signal(madgpu_ctx* ctx, const void* field)
{
  if (!ctx->p2p) return;
  madgpu_ctx::Shared* s = shared_of(ctx, field);
  if (!s) return;
  s->produced = ++ctx->halo_seq;
  if (ctx->rank == ctx->p2p_drop_rank && ctx->p2p_drop_seq >= 0 && (int)s->produced >= ctx->p2p_drop_seq) return;  // test hook: this rank's signals stop arriving
  if (ctx->flags_lo) latch(ctx, g_write_value(ctx->stream, (unsigned long long)(uintptr_t)(ctx->flags_lo + 1), s->produced, 0), "cuStreamWriteValue32");
  if (ctx->flags_hi) latch(ctx, g_write_value(ctx->stream, (unsigned long long)(uintptr_t)(ctx->flags_hi + 0), s->produced, 0), "cuStreamWriteValue32");
}

// a writer that does NOT store into the neighbours' ghost planes: consumers must exchange explicitly
void halo_dirty(madgpu_ctx* ctx, const void* field)
{
  if (!ctx->p2p) return;
  if (madgpu_ctx::Shared* s = shared_of(ctx, field)) s->produced = 0;
}

// before a kernel that reads the ghost planes of `field`: true when the neighbours' peer stores are being waited for,
// false when the caller has to run the explicit exchange
bool halo_wait(madgpu_ctx* ctx, const void* field)
{
  if (!ctx->p2p) return false;
  madgpu_ctx::Shared* s = shared_of(ctx, field);
  if (!s || s->produced == 0) return false;
  if (ctx->p2p_wait_kernel) {  // bounded wait: a lost signal costs a time-out and an error, not a hung stream
    MAD_LAUNCH((k_halo_wait), 1, 1, 0, ctx->stream, (const volatile unsigned*)ctx->flags, s->produced, s->produced, ctx->flags_lo != nullptr,
               ctx->flags_hi != nullptr, ctx->p2p_timeout_cycles, ctx->d_scalar + 1, (unsigned*)ctx->flags + 8);
    ctx->launches++;
    return true;
  }
  const unsigned int GEQ = 1;  // CU_STREAM_WAIT_VALUE_GEQ
  if (ctx->flags_lo) latch(ctx, g_wait_value(ctx->stream, (unsigned long long)(uintptr_t)(ctx->flags + 0), s->produced, GEQ), "cuStreamWaitValue32");
  if (ctx->flags_hi) latch(ctx, g_wait_value(ctx->stream, (unsigned long long)(uintptr_t)(ctx->flags + 1), s->produced, GEQ), "cuStreamWaitValue32");
  return true;
}

// Ghost planes of a level field: plane -1 <- last plane of rank-1, plane nz <- first plane of rank+1 (one ncclSend/Recv
// pair per neighbour on the solver's stream).  The outer ranks keep their Neumann mirror handling (zlo_phys / zhi_phys).
template <typename T>
void exchange_halo(madgpu_ctx* ctx, const Level& L, T* field)
{
  if (ctx->world == 1) return;
  if (halo_wait(ctx, field)) return;  // produced with peer stores: only the arrival counters have to be awaited
  Scope s(ctx, MADGPU_K_HALO, 0);
  const size_t cnt = (size_t)L.g.plane * (sizeof(T) / sizeof(float));
  float* f = reinterpret_cast<float*>(field);
  const size_t pl = (size_t)L.g.plane * (sizeof(T) / sizeof(float));
  NCV(g_nccl.GroupStart());
  if (ctx->rank > 0) {
    NCV(g_nccl.Send(f, cnt, Nccl::Float32, ctx->rank - 1, ctx->comm, ctx->stream));
    NCV(g_nccl.Recv(f - pl, cnt, Nccl::Float32, ctx->rank - 1, ctx->comm, ctx->stream));
  }
  if (ctx->rank < ctx->world - 1) {
    NCV(g_nccl.Send(f + (size_t)(L.g.nz - 1) * pl, cnt, Nccl::Float32, ctx->rank + 1, ctx->comm, ctx->stream));
    NCV(g_nccl.Recv(f + (size_t)L.g.nz * pl, cnt, Nccl::Float32, ctx->rank + 1, ctx->comm, ctx->stream));
  }
  NCV(g_nccl.GroupEnd());
}

// dense slab <-> pitched field of the agglomeration level
void pack_slab(madgpu_ctx* ctx, const Level& L, const float* pitched, float* dense)
{
  const dim3 b = block3(3), g = grid3(L.g, b);
  MAD_LAUNCH((k_pitched_to_dense<float, float>), g, b, 0, ctx->stream, L.g, pitched, dense);
  ctx->launches++;
}
void unpack_slab(madgpu_ctx* ctx, const Level& L, const float* dense, float* pitched)
{
  const dim3 b = block3(3), g = grid3(L.g, b);
  MAD_LAUNCH((k_dense_to_pitched<float, float>), g, b, 0, ctx->stream, L.g, dense, pitched);
  halo_dirty(ctx, pitched);
  ctx->launches++;
}

// every rank's dense slab (slab_buf) -> rank 0's dense global field (gather_buf); equal slab sizes
void gather_slabs(madgpu_ctx* ctx, const Level& L)
{
  const size_t cnt = (size_t)L.n[0] * L.n[1] * L.n[2];
  NCV(g_nccl.GroupStart());
  if (ctx->rank == 0) {
    for (int r = 1; r < ctx->world; ++r) NCV(g_nccl.Recv(ctx->gather_buf + (size_t)r * cnt, cnt, Nccl::Float32, r, ctx->comm, ctx->stream));
  } else {
    NCV(g_nccl.Send(ctx->slab_buf, cnt, Nccl::Float32, 0, ctx->comm, ctx->stream));
  }
  NCV(g_nccl.GroupEnd());
  if (ctx->rank == 0) cudaMemcpyAsync(ctx->gather_buf, ctx->slab_buf, cnt * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream);
}
void scatter_slabs(madgpu_ctx* ctx, const Level& L)
{
  const size_t cnt = (size_t)L.n[0] * L.n[1] * L.n[2];
  NCV(g_nccl.GroupStart());
  if (ctx->rank == 0) {
    for (int r = 1; r < ctx->world; ++r) NCV(g_nccl.Send(ctx->gather_buf + (size_t)r * cnt, cnt, Nccl::Float32, r, ctx->comm, ctx->stream));
  } else {
    NCV(g_nccl.Recv(ctx->slab_buf, cnt, Nccl::Float32, 0, ctx->comm, ctx->stream));
  }
  NCV(g_nccl.GroupEnd());
  if (ctx->rank == 0) cudaMemcpyAsync(ctx->slab_buf, ctx->gather_buf, cnt * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream);
}

void vcycle(madgpu_ctx* ctx, int l, bool zero_guess = false);
void op_zero(madgpu_ctx* ctx, Level& L, float* p);

// The agglomeration level (last level of a slab context): its right-hand side is gathered onto rank 0, which runs the
// rest of the V-cycle on its serial sub-hierarchy, and the correction is scattered back.
void agglomerated_solve(madgpu_ctx* ctx)
{
  Level& L = ctx->lv[ctx->nlevels - 1];
  Scope s(ctx, MADGPU_K_COARSE, 0);
  pack_slab(ctx, L, L.f, ctx->slab_buf);
  gather_slabs(ctx, L);
  if (ctx->rank == 0) {
    madgpu_ctx* S = ctx->sub;
    Level& S0 = S->lv[0];
    unpack_slab(S, S0, ctx->gather_buf, S0.f);
    vcycle(S, 0, true);  // zero initial guess
    pack_slab(S, S0, S0.u, ctx->gather_buf);
    ctx->launches += S->launches;
    S->launches = 0;
  }
  scatter_slabs(ctx, L);
  unpack_slab(ctx, L, ctx->slab_buf, L.u);
}

// ---- operators ----------------------------------------------------------------------------
void op_zero(madgpu_ctx* ctx, Level& L, float* p)
{
  Scope s(ctx, MADGPU_K_MISC);
  halo_dirty(ctx, p);
  cudaMemsetAsync(p, 0, (size_t)L.g.plane * L.g.nz * sizeof(float), ctx->stream);
}

// ---- temporal blocking of the Gauss-Seidel sweeps (fast::k_coef_gs_tb) -----------------------------------------------------
// warps (row pairs) per CTA: tiles of 128 x 16 voxels; the staged single-sweep variant (MADGPU_GS_TB_SINGLE=2) uses 128 x 8 so that
// two CTAs fit the shared memory of an SM
int tb_wp(const madgpu_ctx* ctx) { return ctx->gs_tb_single == 2 ? 4 : 8; }
bool use_tb(const madgpu_ctx* ctx, const Level& L)
{
  return (ctx->gs_tb > 1 || ctx->gs_tb_single) && ctx->gs_fused && ctx->gs_coef16 && !L.coef16_off && use_fast(ctx, L) && gs_pairs(ctx, L) && L.g.ny >= 8 && L.g.nz >= 8;
}
// planes per CTA: ~4 waves of the two resident CTAs per SM; long chunks keep the fill / drain steps of the sweep pipeline (2 per
// fused sweep) and the frozen z faces rare
int tb_zc(const Geom& g, int wp)
{
  const long long cxy = (long long)((g.nx + fast::TX - 1) / fast::TX) * ((g.ny + 2 * wp - 1) / (2 * wp));
  const long long chunks = std::max(1ll, (148ll * 2 * 4) / std::max(1ll, cxy));
  int zc = (int)std::max(16ll, (g.nz + chunks - 1) / chunks);
  zc = (zc + 1) & ~1;
  return std::min(zc, (g.nz + 1) & ~1);
}
// sweeps fused by the next pass when `remaining` sweeps of the leg are left
int tb_fuse(const madgpu_ctx* ctx, int remaining) { return std::min(remaining >= 3 ? 3 : remaining, ctx->gs_tb); }

template <int S, int WP, int MINB, bool STAGE>
void launch_tb(madgpu_ctx* ctx, Level& L, const Geom& gg, int uz)
{
  static bool attr_set = false;  // per instantiation; the attribute is a property of the function
  const size_t smem = fast::tb_smem_bytes(S, WP) + (STAGE ? fast::tb_stage_bytes(WP) : 0);
  if (!attr_set) { cudaFuncSetAttribute(fast::k_coef_gs_tb<S, WP, MINB, STAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr_set = true; }
  const int zc = tb_zc(L.g, WP);
  const int oy = L.tb_flip ? WP : 0, oz = L.tb_flip ? zc / 2 : 0;
  const dim3 grid((L.g.nx + fast::TX - 1) / fast::TX, (L.g.ny + oy + 2 * WP - 1) / (2 * WP), (L.g.nz + oz + zc - 1) / zc);
  MAD_LAUNCH((fast::k_coef_gs_tb<S, WP, MINB, STAGE>), grid, dim3(32, WP), smem, ctx->stream, gg, L.coef16, L.u, L.f, L.tmp, zc, oy, oz, ctx->pf_dist, uz);
  L.tb_flip ^= 1;
}

double read_scalar(madgpu_ctx* ctx);

// Packed fp16 operator rows of a level (fast::MODE_COEF), built once per tensor.  False when the level cannot use them: out of
// device memory (latched as an error) or a diagonal beyond the range of the packed 1/diag (the level then keeps exact rows).
bool build_coef16(madgpu_ctx* ctx, Level& L)
{
  if (L.coef16_valid) return true;
  if (!L.coef16) {
    void* q = nullptr;
    const size_t bytes = (size_t)L.g.nz * L.g.ny * (size_t)(L.g.pitch >> 2) * fast::COEF_WORDS * sizeof(uint4);
    if (cudaMalloc(&q, bytes) != cudaSuccess) {
      if (ctx->sticky.empty()) ctx->sticky = "out of device memory for the packed Gauss-Seidel rows";
      cudaGetLastError();
      L.coef16_off = true;
      return false;
    }
    L.allocs.push_back(q);
    L.coef16 = (uint4*)q;
  }
  {
    Scope sb(ctx, MADGPU_K_MISC);
    cudaMemsetAsync(ctx->d_scalar + 2, 0, sizeof(double), ctx->stream);
    launch_fast<fast::MODE_COEF, float, float, float, float>(ctx, L, L.u, L.f, reinterpret_cast<float*>(L.coef16), ctx->d_scalar + 2, 0.f);
  }
  double flag = 0.0;  // once per tensor and level: a host round trip is affordable here
  cudaMemcpyAsync(ctx->h_scalar + 2, ctx->d_scalar + 2, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  flag = ctx->h_scalar[2];
  if (flag != 0.0) { L.coef16_off = true; return false; }
  L.coef16_valid = true;
  return true;
}

// n_iter smoother iterations on (L.u, L.f); result in L.u (pointers may be swapped with L.tmp).
// zero_first: the iterate is identically zero on entry (every V-cycle leg starts like that): the streaming kernels then
// skip reading it (and the memset that would have produced it); the generic kernels get the memset.
void op_smooth(madgpu_ctx* ctx, int l, int smoother, int n_iter, bool zero_first = false)
{
  Level& L = ctx->lv[l];
  const dim3 b = block3(ctx->dim), g = grid3(L.g, b);
  const int cls = l == 0 ? MADGPU_K_SMOOTH0 : MADGPU_K_SMOOTHC;
  const Tensor D = tensor_of(L);
  const bool streaming = (use_fast(ctx, L) && (smoother == MADGPU_SMOOTHER_WJ || ctx->gs_fused)) || (smoother == MADGPU_SMOOTHER_WJ ? use_fast2(ctx, L) : use_fast2_gs(ctx, L));
  if (zero_first && (!streaming || n_iter == 0)) { op_zero(ctx, L, L.u); zero_first = false; }
  for (int it = 0; it < n_iter; ++it) {
    const int uz = zero_first && it == 0;
    if (!uz) exchange_halo(ctx, L, L.u);
    Geom gg = L.g;  // + the neighbours' ghost planes of the field this sweep produces (L.tmp)
    set_ghosts(ctx, gg, L, L.tmp, sizeof(float));
    if (smoother == MADGPU_SMOOTHER_WJ) {
      Scope s(ctx, cls);
      if (use_fast(ctx, L)) launch_fast<fast::MODE_WJ, float, float, float, float>(ctx, L, L.u, L.f, L.tmp, nullptr, (float)ctx->p.omega, uz);
      else if (use_fast2(ctx, L)) launch_fast2<fast::M2_WJ, float, float, float>(ctx, L, L.u, L.f, L.tmp, nullptr, (float)ctx->p.omega, uz);
      else if (ctx->dim == 3) MAD_LAUNCH((k_jacobi<3>), g, b, 0, ctx->stream, L.g, D, L.u, L.f, L.tmp, (float)ctx->p.omega);
      else MAD_LAUNCH((k_jacobi<2>), g, b, 0, ctx->stream, L.g, D, L.u, L.f, L.tmp, (float)ctx->p.omega);
      if (!use_fast(ctx, L)) halo_dirty(ctx, L.tmp);
      std::swap(L.u, L.tmp);
    } else if (use_fast2_gs(ctx, L)) {
      // 2-D: one pass, rows in y order, even then odd columns, exact inside a tile of 128 columns x yc rows (mad_fast2d.cuh)
      Scope s(ctx, cls);
      launch_fast2<fast::M2_GS, float, float, float>(ctx, L, L.u, L.f, L.tmp, nullptr, 0.f, uz);
      std::swap(L.u, L.tmp);
    } else if (use_fast(ctx, L) && ctx->gs_fused && ctx->gs_coef16 && !L.coef16_off && build_coef16(ctx, L)) {
      // fused sweep fed by pre-evaluated fp16 operator rows (built once per tensor and level)
      Scope s(ctx, cls);
      const int fuse = use_tb(ctx, L) ? tb_fuse(ctx, n_iter - it) : 0;
      if (fuse >= 1) {  // `fuse` sweeps of this leg in one pass over shared-memory plane rings (temporal blocking; 1: MADGPU_GS_TB_SINGLE)
        if (fuse == 3) launch_tb<3, 8, 2, false>(ctx, L, gg, uz);
        else if (fuse == 2) launch_tb<2, 8, 2, false>(ctx, L, gg, uz);
        else if (ctx->gs_tb_single == 2) launch_tb<1, 4, 2, true>(ctx, L, gg, uz);   // operands staged by cp.async, tiles of 128 x 8, two CTAs per SM
        else if (ctx->gs_tb_single == 3) launch_tb<1, 8, 1, true>(ctx, L, gg, uz);   // the same on tiles of 128 x 16, one CTA per SM
        else launch_tb<1, 8, 2, false>(ctx, L, gg, uz);
        it += fuse - 1;
      } else if (gs_pairs(ctx, L)) {  // one warp per row pair: tile 128 x 8 x zc (128 x 2 x zc with MADGPU_GS_PRIVATE: no barriers)
        const int zc = fast_zc(L.g, 8);
        // (capped at 128 registers for 16 warps per SM both variants spill and take 0.89 instead of 0.77 / 0.73 ms: profiles/r02o_*)
        if (ctx->gs_private) {
          // MADGPU_GS_PRIVATE=2: the pairs start at odd rows in every other sweep, so that the tile faces of one sweep lie inside the
          // pairs of the next (the grid then has one more pair: (-1, 0) ... (ny-1, ny))
          const int ysh = ctx->gs_private == 2 ? L.tb_flip : 0;
          dim3 fg = fast_grid(L.g, 8, zc);
          fg.y = (L.g.ny + ysh + 7) / 8;
          MAD_LAUNCH((fast::k_coef_gs2<4, 3, true>), fg, dim3(32, 4), 0, ctx->stream, gg, L.coef16, L.u, L.f, L.tmp, zc, ctx->pf_dist, uz, ysh);
          if (ctx->gs_private == 2) L.tb_flip ^= 1;
        } else MAD_LAUNCH((fast::k_coef_gs2<4, 3, false>), fast_grid(L.g, 8, zc), dim3(32, 4), 0, ctx->stream, gg, L.coef16, L.u, L.f, L.tmp, zc, ctx->pf_dist, uz, 0);
      } else {
        const int zc = fast_zc(L.g, 4);
        MAD_LAUNCH((fast::k_coef_gs<4, 4>), fast_grid(L.g, 4, zc), dim3(32, 4), 0, ctx->stream, gg, L.coef16, L.u, L.f, L.tmp, zc, ctx->pf_dist, uz);
      }
      std::swap(L.u, L.tmp);
      halo_signal(ctx, L.u);
    } else if (use_fast(ctx, L) && ctx->gs_fused) {
      // one pass: z-ordered planes, four in-plane colours, exact inside a CTA tile (mad_fast.cuh)
      Scope s(ctx, cls);
      if (ctx->fast_cfg == 1) {
        const int zc = fast_zc(L.g, 8);
        MAD_LAUNCH((fast::k_fast_gs<8, 1, false, false>), fast_grid(L.g, 8, zc), dim3(32, 8), 0, ctx->stream, gg, D, L.u, L.f, L.tmp, zc, ctx->pf_dist, uz);
      } else if (ctx->fast_cfg == 4) {
        const int zc = fast_zc(L.g, 4);
        MAD_LAUNCH((fast::k_fast_gs<4, 2, true, false>), fast_grid(L.g, 4, zc), dim3(32, 4), 0, ctx->stream, gg, D, L.u, L.f, L.tmp, zc, ctx->pf_dist, uz);
      } else if (ctx->fast_cfg == 5) {
        const int zc = fast_zc(L.g, 8);
        MAD_LAUNCH((fast::k_fast_gs<8, 1, true, false>), fast_grid(L.g, 8, zc), dim3(32, 8), 0, ctx->stream, gg, D, L.u, L.f, L.tmp, zc, ctx->pf_dist, uz);
      } else if (ctx->fast_cfg == 7) {
        const int zc = fast_zc(L.g, 4);
        MAD_LAUNCH((fast::k_fast_gs<4, 3, false, true>), fast_grid(L.g, 4, zc), dim3(32, 4), 0, ctx->stream, gg, D, L.u, L.f, L.tmp, zc, ctx->pf_dist, uz);
      } else if (ctx->fast_cfg == 8) {
        const int zc = fast_zc(L.g, 8);
        MAD_LAUNCH((fast::k_fast_gs<8, 1, false, true>), fast_grid(L.g, 8, zc), dim3(32, 8), 0, ctx->stream, gg, D, L.u, L.f, L.tmp, zc, ctx->pf_dist, uz);
      } else if (ctx->fast_cfg == 6) {
        const int zc = fast_zc(L.g, 2);
        MAD_LAUNCH((fast::k_fast_gs<2, 6, false, false>), fast_grid(L.g, 2, zc), dim3(32, 2), 0, ctx->stream, gg, D, L.u, L.f, L.tmp, zc, ctx->pf_dist, uz);
      } else {
        const int zc = fast_zc(L.g, 4);
        MAD_LAUNCH((fast::k_fast_gs<4, 3, false, false>), fast_grid(L.g, 4, zc), dim3(32, 4), 0, ctx->stream, gg, D, L.u, L.f, L.tmp, zc, ctx->pf_dist, uz);
      }
      std::swap(L.u, L.tmp);
      halo_signal(ctx, L.u);
    } else {
      const int nc = ctx->dim == 2 ? 4 : (ctx->p.gs_colors == 8 ? 8 : 4);
      Scope s(ctx, cls, nc);
      for (int c = 0; c < nc; ++c) {
        if (ctx->dim == 3) MAD_LAUNCH((k_gs_color<3>), g, b, 0, ctx->stream, L.g, D, L.u, L.f, c, nc);
        else MAD_LAUNCH((k_gs_color<2>), g, b, 0, ctx->stream, L.g, D, L.u, L.f, c, nc);
      }
      halo_dirty(ctx, L.u);
    }
  }
}

double read_scalar(madgpu_ctx* ctx)
{
  const bool watch = ctx->p2p && ctx->p2p_wait_kernel;  // d_scalar[1]: "a halo wait timed out on some rank" (all-reduced with the norm)
  cudaMemcpyAsync(ctx->h_scalar, ctx->d_scalar, (watch ? 2 : 1) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  if (watch && ctx->h_scalar[1] != 0.0) {
    unsigned d[4] = {0, 0, 0, 0};
    cudaMemcpy(d, ctx->flags + 8, sizeof d, cudaMemcpyDeviceToHost);
    char msg[256];
    snprintf(msg, sizeof msg, "peer-memory halo: an arrival-counter wait timed out on %d rank(s); rank %d last wanted %u / %u from its lower / upper "
             "neighbour and saw %u / %u -- results are invalid, the context must be recreated (NCCL halo)", (int)ctx->h_scalar[1], ctx->rank, d[0], d[1], d[2], d[3]);
    if (ctx->sticky.empty()) ctx->sticky = msg;
    ctx->p2p = false;
  }
  return *ctx->h_scalar;
}

// sum of the per-block partials -> d_scalar (device); caller reads it back when needed
void reduce_partials(madgpu_ctx* ctx, size_t n)
{
  MAD_LAUNCH((k_reduce_partials), 1, 1024, 0, ctx->stream, ctx->partials, (long long)n, ctx->d_scalar);
  if (ctx->world > 1) NCV(g_nccl.AllReduce(ctx->d_scalar, ctx->d_scalar, ctx->p2p && ctx->p2p_wait_kernel ? 2 : 1, Nccl::Float64, Nccl::Sum, ctx->comm, ctx->stream));
}

// L.tmp = L.f - A L.u  (fp32)
void op_residual32(madgpu_ctx* ctx, int l, float* out, bool norm)
{
  Level& L = ctx->lv[l];
  const dim3 b = block3(ctx->dim), g = grid3(L.g, b);
  exchange_halo(ctx, L, L.u);
  Scope s(ctx, MADGPU_K_RESTRICT, norm ? 2 : 1);
  const Tensor D = tensor_of(L);
  double* part = norm ? ctx->partials : nullptr;
  if (!norm && ctx->p.smoother == MADGPU_SMOOTHER_GS && use_fast(ctx, L) && ctx->gs_fused && ctx->gs_coef16 && !L.coef16_off && L.coef16_valid) {
    // inside a Gauss-Seidel V-cycle: residual with the packed rows the sweeps use
    const int zc = fast_zc(L.g, 4);
    Geom gg = L.g;
    set_ghosts(ctx, gg, L, out, sizeof(float));
    MAD_LAUNCH((fast::k_coef_residual<4, 4>), fast_grid(L.g, 4, zc), dim3(32, 4), 0, ctx->stream, gg, L.coef16, L.u, L.f, out, zc, ctx->pf_dist);
    halo_signal(ctx, out);
    return;
  }
  if (use_fast(ctx, L)) {
    const size_t nb = launch_fast<fast::MODE_RES, float, float, float, float>(ctx, L, L.u, L.f, out, part, 0.f);
    if (norm) reduce_partials(ctx, nb);
    return;
  }
  if (use_fast2(ctx, L)) {
    const size_t nb = launch_fast2<fast::M2_RES, float, float, float>(ctx, L, L.u, L.f, out, part, 0.f);
    if (norm) reduce_partials(ctx, nb);
    return;
  }
  if (ctx->dim == 3) MAD_LAUNCH((k_residual<3, float, float, float, float>), g, b, 0, ctx->stream, L.g, D, L.u, L.f, out, part);
  else MAD_LAUNCH((k_residual<2, float, float, float, float>), g, b, 0, ctx->stream, L.g, D, L.u, L.f, out, part);
  halo_dirty(ctx, out);
  if (norm) reduce_partials(ctx, (size_t)g.x * g.y * g.z);
}

// level-0 outer residual in fp64 arithmetic: r32 = f64 - A u64, ||r||^2 -> d_scalar
void op_residual64(madgpu_ctx* ctx, float* r32_or_null, double* r64_or_null)
{
  Level& L = ctx->lv[0];
  const dim3 b = block3(ctx->dim), g = grid3(L.g, b);
  exchange_halo(ctx, L, ctx->u64);
  Scope s(ctx, MADGPU_K_RESID0, 2);
  const Tensor D = tensor_of(L);
  if (r64_or_null) {
    if (ctx->dim == 3) MAD_LAUNCH((k_residual<3, double, double, double, double>), g, b, 0, ctx->stream, L.g, D, ctx->u64, ctx->f64, r64_or_null, ctx->partials);
    else MAD_LAUNCH((k_residual<2, double, double, double, double>), g, b, 0, ctx->stream, L.g, D, ctx->u64, ctx->f64, r64_or_null, ctx->partials);
  } else if (use_fast(ctx, L) && ctx->res64_c32 && ctx->res64_smem) {
    // y-neighbour rows through a shared-memory ring (k_fast_res64): a third of the registers of the all-register form
    constexpr int WY = 4;
    const int zc = fast_zc(L.g, WY);
    const dim3 fg = fast_grid(L.g, WY, zc);
    Geom gg = L.g;
    set_ghosts(ctx, gg, L, r32_or_null, sizeof(float));
    if (ctx->res64_smem == 4) MAD_LAUNCH((fast::k_fast_res64<WY, 4>), fg, dim3(32, WY), 0, ctx->stream, gg, D, (const double*)ctx->u64, (const double*)ctx->f64, r32_or_null, ctx->partials, zc, ctx->pf_dist);
    else MAD_LAUNCH((fast::k_fast_res64<WY, 3>), fg, dim3(32, WY), 0, ctx->stream, gg, D, (const double*)ctx->u64, (const double*)ctx->f64, r32_or_null, ctx->partials, zc, ctx->pf_dist);
    halo_signal(ctx, r32_or_null);
    reduce_partials(ctx, (size_t)fg.x * fg.y * fg.z);
    return;
  } else if (use_fast(ctx, L)) {
    const size_t nb = ctx->res64_c32 ? launch_fast<fast::MODE_RES_C32, double, double, double, float>(ctx, L, ctx->u64, ctx->f64, r32_or_null, ctx->partials, 0.f)
                                     : launch_fast<fast::MODE_RES, double, double, double, float>(ctx, L, ctx->u64, ctx->f64, r32_or_null, ctx->partials, 0.f);
    reduce_partials(ctx, nb);
    return;
  } else if (use_fast2(ctx, L)) {
    reduce_partials(ctx, launch_fast2<fast::M2_RES, double, double, double>(ctx, L, ctx->u64, ctx->f64, r32_or_null, ctx->partials, 0.f));
    return;
  } else {
    if (ctx->dim == 3) MAD_LAUNCH((k_residual<3, double, double, double, float>), g, b, 0, ctx->stream, L.g, D, ctx->u64, ctx->f64, r32_or_null, ctx->partials);
    else MAD_LAUNCH((k_residual<2, double, double, double, float>), g, b, 0, ctx->stream, L.g, D, ctx->u64, ctx->f64, r32_or_null, ctx->partials);
  }
  reduce_partials(ctx, (size_t)g.x * g.y * g.z);
}

Transfer transfer_of(const Level& coarse)
{
  Transfer t;
  for (int d = 0; d < 3; ++d) t.cent[d] = coarse.cent[d];
  return t;
}

template <typename TI>
void op_restrict(madgpu_ctx* ctx, int lf, const TI* fine, float* coarse, int cls = MADGPU_K_RESTRICT)
{
  Level& F = ctx->lv[lf];
  Level& C = ctx->lv[lf + 1];
  const dim3 b = ctx->dim == 3 ? dim3(32, 4, 2) : dim3(32, 8, 1);
  const dim3 g = grid3(C.g, b);
  exchange_halo(ctx, F, const_cast<TI*>(fine));
  Scope s(ctx, cls);
  if constexpr (std::is_same<TI, float>::value) {
    if (use_fast(ctx, F) && F.g.nx >= 8) {
      constexpr int WY = 8;
      if (C.cent[0] == 1 && C.cent[1] == 1 && C.cent[2] == 1 && ctx->restrict_cell) {
        // cell-centred along all axes (power-of-two volumes): the marching kernel, every fine plane reduced once
        const int gx = (F.g.nx + fast::TX - 1) / fast::TX, gy = (C.g.ny + WY - 1) / WY;
        const int chunks = std::max(1, std::min(C.g.nz, (148 * 8 + gx * gy - 1) / (gx * gy)));
        const int zcc = (C.g.nz + chunks - 1) / chunks;
        MAD_LAUNCH((fast::k_fast_restrict_cell<WY>), dim3(gx, gy, (C.g.nz + zcc - 1) / zcc), dim3(32, WY), 0, ctx->stream, F.g, C.g, fine, coarse, zcc);
        return;
      }
      const dim3 fg((F.g.nx + fast::TX - 1) / fast::TX, (C.g.ny + WY - 1) / WY, C.g.nz);
      MAD_LAUNCH((fast::k_fast_restrict<WY>), fg, dim3(32, WY), 0, ctx->stream, F.g, C.g, transfer_of(C), fine, coarse);
      return;
    }
    if (use_fast2(ctx, F)) {
      constexpr int WY = 4;
      const dim3 fg((F.g.nx + fast::TX - 1) / fast::TX, (C.g.ny + WY - 1) / WY);
      MAD_LAUNCH((fast::k2_restrict<WY>), fg, dim3(32, WY), 0, ctx->stream, F.g, C.g, transfer_of(C), fine, coarse);
      return;
    }
  }
  if (ctx->dim == 3) MAD_LAUNCH((k_restrict<3, TI, float>), g, b, 0, ctx->stream, F.g, C.g, transfer_of(C), fine, coarse);
  else MAD_LAUNCH((k_restrict<2, TI, float>), g, b, 0, ctx->stream, F.g, C.g, transfer_of(C), fine, coarse);
}

template <typename TO, bool ADD>
void op_prolong(madgpu_ctx* ctx, int lf, const float* coarse, TO* fine)
{
  Level& F = ctx->lv[lf];
  Level& C = ctx->lv[lf + 1];
  const dim3 b = ctx->dim == 3 ? dim3(32, 4, 2) : dim3(32, 8, 1);
  const dim3 g = grid3(F.g, b);
  exchange_halo(ctx, C, const_cast<float*>(coarse));
  Scope s(ctx, MADGPU_K_PROLONG);
  if constexpr (std::is_same<TO, float>::value) {
    if (use_fast(ctx, F)) {
      constexpr int WY = 8;
      Geom gg = F.g;
      set_ghosts(ctx, gg, F, fine, sizeof(float));
      if (C.cent[0] == 1 && C.cent[1] == 1 && C.cent[2] == 1 && ctx->prolong_cell) {
        // cell-centred along all axes (power-of-two volumes): the blocked kernel, one coarse plane loaded per two fine planes
        const int gx = (F.g.nx + fast::TX - 1) / fast::TX, gy = (C.g.ny + WY - 1) / WY;
        const int chunks = std::max(1, std::min(C.g.nz, (148 * 8 + gx * gy - 1) / (gx * gy)));
        const int zcc = (C.g.nz + chunks - 1) / chunks;
        MAD_LAUNCH((fast::k_fast_prolong_cell<ADD, WY>), dim3(gx, gy, (C.g.nz + zcc - 1) / zcc), dim3(32, WY), 0, ctx->stream, C.g, gg, coarse, fine, zcc);
        halo_signal(ctx, fine);
        return;
      }
      const dim3 fg((F.g.nx + fast::TX - 1) / fast::TX, (F.g.ny + WY - 1) / WY, (F.g.nz + fast::PROLONG_ZB - 1) / fast::PROLONG_ZB);
      MAD_LAUNCH((fast::k_fast_prolong<ADD, WY>), fg, dim3(32, WY), 0, ctx->stream, C.g, gg, transfer_of(C), coarse, fine);
      halo_signal(ctx, fine);
      return;
    }
    if (use_fast2(ctx, F)) {
      constexpr int WY = 4;
      const dim3 fg((F.g.nx + fast::TX - 1) / fast::TX, (F.g.ny + WY - 1) / WY);
      MAD_LAUNCH((fast::k2_prolong<ADD, WY>), fg, dim3(32, WY), 0, ctx->stream, C.g, F.g, transfer_of(C), coarse, fine);
      return;
    }
  }
  if (ctx->dim == 3) MAD_LAUNCH((k_prolong<3, TO, ADD>), g, b, 0, ctx->stream, C.g, F.g, transfer_of(C), coarse, fine);
  else MAD_LAUNCH((k_prolong<2, TO, ADD>), g, b, 0, ctx->stream, C.g, F.g, transfer_of(C), coarse, fine);
  halo_dirty(ctx, fine);
}

void op_coarse_solve(madgpu_ctx* ctx)
{
  Level& L = ctx->lv[ctx->nlevels - 1];
  if (ctx->coarse_direct) {
    Scope s(ctx, MADGPU_K_COARSE);
    const int n = ctx->ncoarse;
    const int rows_per_block = 8;
    MAD_LAUNCH((k_coarse_gemv), (n + rows_per_block - 1) / rows_per_block, 32 * rows_per_block, n * sizeof(double), ctx->stream, L.g, ctx->Ainv, L.f, L.u, n);
  } else {
    // Coarsest grid too large for a dense inverse (thin volumes stop coarsening early): iterate the
    // multicolour Gauss-Seidel smoother to fp32 convergence instead of vnl_sparse_lu.
    const int l = ctx->nlevels - 1;
    const dim3 b = block3(ctx->dim), g = grid3(L.g, b);
    {
      Scope s(ctx, MADGPU_K_COARSE, 2);
      MAD_LAUNCH((k_sumsq<float>), g, b, 0, ctx->stream, L.g, L.f, ctx->partials);
      reduce_partials(ctx, (size_t)g.x * g.y * g.z);
    }
    const double f_norm = std::sqrt(read_scalar(ctx));
    op_zero(ctx, L, L.u);
    if (f_norm == 0.0) return;
    // fp32 Gauss-Seidel cannot push the fp32 residual below a few 1e-7 ||f||: stop at 1e-6, or as soon as a chunk of sweeps
    // no longer halves it (stagnation at the rounding floor) -- the outer defect correction absorbs what is left
    double prev = f_norm;
    for (int chunk = 0; chunk < 400; ++chunk) {
      op_smooth(ctx, l, MADGPU_SMOOTHER_GS, 16);
      op_residual32(ctx, l, L.tmp, true);
      const double r = std::sqrt(read_scalar(ctx));
      if (!(r > 1e-6 * f_norm) || (r > 0.5 * prev && r < 1e-4 * f_norm)) break;
      prev = r;
    }
  }
}

void vcycle_body(madgpu_ctx* ctx, int l, bool zero_guess)
{
  if (l == ctx->nlevels - 1) {  // :356-371  (both solvers overwrite the whole iterate)
    if (ctx->world > 1) agglomerated_solve(ctx);
    else op_coarse_solve(ctx);
    return;
  }
  Level& L = ctx->lv[l];
  Level& C = ctx->lv[l + 1];
  const int nu = ctx->p.iterations_per_grid;
  op_smooth(ctx, l, ctx->p.smoother, nu, zero_guess);     // :384-387
  op_residual32(ctx, l, L.tmp, false);                    // :389 (last one only)
  op_restrict<float>(ctx, l, L.tmp, C.f);                 // :413
  vcycle(ctx, l + 1, true);                               // :415-420, zero coarse guess
  op_prolong<float, true>(ctx, l, C.u, L.u);              // :422-435
  op_smooth(ctx, l, ctx->p.smoother, nu);                 // :460-463
}

void drop_graphs(madgpu_ctx* ctx)
{
  for (auto& g : ctx->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  ctx->graphs.clear();
}

// The coarse part of a cycle is dozens of launches of a few microseconds each (66 of the 96 launches of a 512^3 cycle): from the
// first level of at most graph_voxels voxels down it is captured into a CUDA graph -- every shape, pointer and loop count below
// that level is static between two tensors -- and replayed with one launch.  The first call with a given key runs uncaptured
// (it builds the packed rows of the levels, which allocates), the second is captured.  Single-context levels only: a z-slab
// context keeps its distributed levels outside (stream memory operations, NCCL), its serial sub-hierarchy on rank 0 qualifies.
bool vcycle_graphed(madgpu_ctx* ctx, int l, bool zero_guess)
{
  if (ctx->capturing || ctx->graph_voxels <= 0 || ctx->world > 1 || !ctx->coarse_direct || l == ctx->nlevels - 1) return false;
  const Level& L = ctx->lv[l];
  if ((long long)L.n[0] * L.n[1] * L.n[2] > ctx->graph_voxels) return false;
  if (l > 0) {  // only the first qualifying level owns a graph; the levels below are part of it
    const Level& P = ctx->lv[l - 1];
    if ((long long)P.n[0] * P.n[1] * P.n[2] <= ctx->graph_voxels) return false;
  }
  std::vector<const void*> state;
  for (int k = l; k < ctx->nlevels; ++k) { state.push_back(ctx->lv[k].u); state.push_back((const void*)(uintptr_t)ctx->lv[k].tb_flip); }
  madgpu_ctx::CycleGraph* g = nullptr;
  for (auto& c : ctx->graphs)
    if (c.level == l && c.zero_guess == (int)zero_guess && c.smoother == ctx->p.smoother && c.nu == ctx->p.iterations_per_grid &&
        c.omega == ctx->p.omega && c.state == state) { g = &c; break; }
  if (!g) {
    if (ctx->graphs.size() >= 16) drop_graphs(ctx);  // callers that keep changing the settings: start over
    ctx->graphs.push_back({l, (int)zero_guess, ctx->p.smoother, ctx->p.iterations_per_grid, ctx->p.omega, state, nullptr, 0, 1, false});
    return false;  // first sight: plain launches (lazy allocations happen here)
  }
  if (g->bad) return false;
  if (!g->exec) {
    const int prof = ctx->profiling;
    const int64_t before = ctx->launches;
    ctx->profiling = 0;
    ctx->capturing = true;
    cudaGraph_t graph = nullptr;
    bool ok = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
      vcycle_body(ctx, l, zero_guess);
      ok = cudaStreamEndCapture(ctx->stream, &graph) == cudaSuccess && graph != nullptr;
    }
    ctx->capturing = false;
    ctx->profiling = prof;
    g->launches = (int)(ctx->launches - before);
    ctx->launches = before;
    if (ok) ok = cudaGraphInstantiate(&g->exec, graph, 0) == cudaSuccess;
    if (graph) cudaGraphDestroy(graph);
    // the body ran its pointer swaps on the host while capturing; they come back to the entry state after a whole cycle
    bool same = true;
    for (int k = l; k < ctx->nlevels; ++k) same = same && ctx->lv[k].u == state[2 * (k - l)] && (const void*)(uintptr_t)ctx->lv[k].tb_flip == state[2 * (k - l) + 1];
    if (!ok || !same) {
      cudaGetLastError();
      if (g->exec) { cudaGraphExecDestroy(g->exec); g->exec = nullptr; }
      g->bad = true;
      if (!same) {  // cannot happen with nu pre- and nu post-sweeps; restore and run plainly
        for (int k = l; k < ctx->nlevels; ++k) {
          if (ctx->lv[k].u != state[2 * (k - l)]) std::swap(ctx->lv[k].u, ctx->lv[k].tmp);
          ctx->lv[k].tb_flip = (int)(uintptr_t)state[2 * (k - l) + 1];
        }
      }
      return false;
    }
  }
  Scope s(ctx, MADGPU_K_GRAPH, g->launches);
  ctx->st.graph_launches++;
  if (cudaGraphLaunch(g->exec, ctx->stream) != cudaSuccess) {
    cudaGetLastError();
    if (ctx->sticky.empty()) ctx->sticky = "cudaGraphLaunch of the coarse-level cycle failed";
  }
  return true;
}

// V-cycle at level l on (L.u, L.f): itkMultigridAnisotropicDiffusionImageFilter.hxx:341-493 without the
// reference's logging-only residual/norm passes (:389-411, :437-439, :464-487).
// zero_guess: the iterate of level l is identically zero on entry and need not have been written (the correction
// equations of the coarser levels, …Filter.hxx:415-416, and of level 0 in defect-correction form).
void vcycle(madgpu_ctx* ctx, int l, bool zero_guess)
{
  if (vcycle_graphed(ctx, l, zero_guess)) return;
  vcycle_body(ctx, l, zero_guess);
}

void op_axpy(madgpu_ctx* ctx)
{
  Level& L = ctx->lv[0];
  const dim3 b(32, 8, 1), g((L.g.nx + 127) / 128, (L.g.ny + 7) / 8, L.g.nz);
  Scope s(ctx, MADGPU_K_MISC);
  Geom gg = L.g;
  set_ghosts(ctx, gg, L, ctx->u64, sizeof(double));
  MAD_LAUNCH((k_axpy_f64_f32), g, b, 0, ctx->stream, gg, ctx->u64, L.u);
  halo_signal(ctx, ctx->u64);
}

// One outer iteration in defect-correction form.  On entry lv[0].f holds r = f64 - A u64.
// On exit u64 is updated, lv[0].f holds the new residual and its squared norm is in d_scalar.
void outer_iteration(madgpu_ctx* ctx, bool smoother_only)
{
  Level& L = ctx->lv[0];
  if (smoother_only) op_smooth(ctx, 0, ctx->p.smoother, 1, true);  // …Filter.hxx:213
  else vcycle(ctx, 0, true);                                       // …Filter.hxx:235
  op_axpy(ctx);
  op_residual64(ctx, L.f, nullptr);                          // …Filter.hxx:215-217 / :237-239
}

void fmg(madgpu_ctx* ctx);

// FullMultiGrid below the agglomeration level of a slab context: the restricted right-hand side of that level is gathered onto
// rank 0, whose serial sub-hierarchy runs its own FullMultiGrid on it (its level 0 IS the agglomeration level, so the nu V-cycles
// of that level, …Filter.hxx:332, are part of it), and the iterate is scattered back.  (2, 4 and 8 GPUs: within 1e-14 of the one-GPU FMG solve.)
void agglomerated_fmg(madgpu_ctx* ctx)
{
  Level& L = ctx->lv[ctx->nlevels - 1];
  Scope s(ctx, MADGPU_K_COARSE, 0);
  pack_slab(ctx, L, L.f, ctx->slab_buf);
  gather_slabs(ctx, L);
  if (ctx->rank == 0) {
    madgpu_ctx* S = ctx->sub;
    Level& S0 = S->lv[0];
    const dim3 b = block3(3), g = grid3(S0.g, b);
    MAD_LAUNCH((k_dense_to_pitched<float, double>), g, b, 0, S->stream, S0.g, ctx->gather_buf, S->f64);
    fmg(S);  // -> S->u64
    MAD_LAUNCH((k_pitched_to_dense<double, float>), g, b, 0, S->stream, S0.g, S->u64, ctx->gather_buf);
    ctx->launches += S->launches + 2;
    S->launches = 0;
  }
  scatter_slabs(ctx, L);
  unpack_slab(ctx, L, ctx->slab_buf, L.u);
}

// Full multigrid prologue: itkMultigridAnisotropicDiffusionImageFilter.hxx:300-338.  Produces u64.
void fmg(madgpu_ctx* ctx)
{
  const int Lmax = ctx->nlevels - 1;
  const int nu = ctx->p.iterations_per_grid;
  if (Lmax == 0) {
    // single level: nu "V-cycles" = direct solves of f (…Filter.hxx:311-314)
    Level& L = ctx->lv[0];
    cudaMemsetAsync(ctx->u64, 0, (size_t)L.g.plane * L.g.nz * sizeof(double), ctx->stream);
    op_residual64(ctx, L.f, nullptr);
    op_zero(ctx, L, L.u);
    vcycle(ctx, 0);
    op_axpy(ctx);
    return;
  }
  // rhs restricted all the way down (:324-326)
  op_restrict<double>(ctx, 0, ctx->f64, ctx->lv[1].f);
  for (int l = 1; l < Lmax; ++l) op_restrict<float>(ctx, l, ctx->lv[l].f, ctx->lv[l + 1].f);
  // coarsest: zero guess, nu V-cycles == direct solve (:311-314); on z-slabs the rest of the recursion runs on rank 0
  if (ctx->world > 1) agglomerated_fmg(ctx);
  else vcycle(ctx, Lmax);
  for (int l = Lmax - 1; l >= 1; --l) {
    op_prolong<float, false>(ctx, l, ctx->lv[l + 1].u, ctx->lv[l].u);  // :330
    for (int it = 0; it < nu; ++it) vcycle(ctx, l);                     // :332
  }
  op_prolong<double, false>(ctx, 0, ctx->lv[1].u, ctx->u64);
  for (int it = 0; it < nu; ++it) {
    op_residual64(ctx, ctx->lv[0].f, nullptr);
    op_zero(ctx, ctx->lv[0], ctx->lv[0].u);
    vcycle(ctx, 0);
    op_axpy(ctx);
  }
}

// ---- setup --------------------------------------------------------------------------------
int level_schedule(int dim, const int* n0, int sizes[][3], int cent[][3])
{
  // mad/itkGridsHierarchy.hxx:36-106
  long g[3] = {1, 1, 1};
  for (int d = 0; d < dim; ++d) g[d] = n0[d];
  bool coarsest = false;
  int nlev = 1;
  while (!coarsest) {
    for (int d = 0; d < dim; ++d) {
      g[d] = (g[d] % 2 == 0) ? g[d] / 2 : ((g[d] - 1) / 2) + 1;
      if (g[d] < 6) coarsest = true;
    }
    ++nlev;
  }
  --nlev;
  if (nlev > MADGPU_MAX_LEVELS) return -1;
  for (int d = 0; d < 3; ++d) { sizes[0][d] = d < dim ? n0[d] : 1; cent[0][d] = 0; }
  for (int l = 1; l < nlev; ++l)
    for (int d = 0; d < 3; ++d) {
      const int nf = sizes[l - 1][d];
      if (d >= dim) { sizes[l][d] = 1; cent[l][d] = 0; }
      else if (nf % 2 == 0) { sizes[l][d] = nf / 2; cent[l][d] = 1; }
      else { sizes[l][d] = (nf - 1) / 2 + 1; cent[l][d] = 0; }
    }
  return nlev;
}

// z-slab plan.  Levels 0..La-1 are distributed: every rank owns gnz_l / world consecutive planes, slab
// starts even at every level so that cell-centred transfers need a one-plane halo; level La is the agglomeration
// level (each rank still holds its slab of it as the target of the last restriction, rank 0 gathers it and owns the
// rest of the hierarchy).  Returns La (>= 1) or -1 when the volume cannot be cut this way.
int plan_slabs(int nlev, const int sizes[][3], int world, std::string& why)
{
  if (nlev < 2) { why = "the volume has a single level, nothing to distribute"; return -1; }
  if (sizes[0][2] % world != 0) { why = "size[2] must be divisible by world_size"; return -1; }
  int lnz = sizes[0][2] / world;
  int La = 0;
  long long small = 64ll * 64 * 64;  // levels of this many voxels or fewer are latency-bound: agglomerate (MADGPU_AGGLOMERATE_VOXELS)
  if (const char* e = getenv("MADGPU_AGGLOMERATE_VOXELS")) small = atoll(e);
  for (int l = 0; l + 1 < nlev; ++l) {
    // can level l be distributed, i.e. can its slabs be restricted to slabs of level l+1?
    const bool ok = sizes[l][2] % 2 == 0 && lnz % 2 == 0 && lnz >= 4;
    const long long vox = (long long)sizes[l][0] * sizes[l][1] * sizes[l][2];
    if (!ok || (l > 0 && vox <= small)) break;
    lnz /= 2;
    La = l + 1;
  }
  if (La < 1) { why = "planes per rank must be even and >= 4 on the finest level"; return -1; }
  return La;
}

void fill_geom(Level& L, int dim, double dt)
{
  Geom& g = L.g;
  g.nx = L.n[0]; g.ny = L.n[1]; g.nz = L.n[2];
  g.pitch = (g.nx + 31) / 32 * 32;
  g.plane = (long long)g.pitch * g.ny;
  g.zlo_phys = 1; g.zhi_phys = 1; g.z0 = 0;
  g.glo = nullptr; g.ghi = nullptr;
  const double hx = L.h[0], hy = L.h[1], hz = dim == 3 ? L.h[2] : 1.0;
  const bool d3 = dim == 3;
  g.dwx = dt / (hx * hx); g.dwy = dt / (hy * hy); g.dwz = d3 ? dt / (hz * hz) : 0.0;
  g.dcxy = dt / (2 * hx * hy); g.dcxz = d3 ? dt / (2 * hx * hz) : 0.0; g.dcyz = d3 ? dt / (2 * hy * hz) : 0.0;
  g.dbxx = dt / (4 * hx * hx); g.dbxy = dt / (4 * hx * hy); g.dbyy = dt / (4 * hy * hy);
  g.dbxz = d3 ? dt / (4 * hx * hz) : 0.0; g.dbyz = d3 ? dt / (4 * hy * hz) : 0.0; g.dbzz = d3 ? dt / (4 * hz * hz) : 0.0;
  g.wx = (float)g.dwx; g.wy = (float)g.dwy; g.wz = (float)g.dwz;
  g.cxy = (float)g.dcxy; g.cxz = (float)g.dcxz; g.cyz = (float)g.dcyz;
  g.bxx = (float)g.dbxx; g.bxy = (float)g.dbxy; g.bxz = (float)g.dbxz; g.byy = (float)g.dbyy; g.byz = (float)g.dbyz; g.bzz = (float)g.dbzz;
  L.elems = (size_t)g.plane * (g.nz + 2);
}

// dense <-> pitched copies of one level field between HOST dense fp32 and device
int upload_field(madgpu_ctx* ctx, const Level& L, const float* host, float* dev)
{
  halo_dirty(ctx, dev);  // written without peer stores: a consumer must run the explicit exchange, not wait for arrival counters
  CU(cudaMemcpy2DAsync(dev, (size_t)L.g.pitch * sizeof(float), host, (size_t)L.g.nx * sizeof(float), (size_t)L.g.nx * sizeof(float),
                       (size_t)L.g.ny * L.g.nz, cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}
int download_field(madgpu_ctx* ctx, const Level& L, const float* dev, float* host)
{
  CU(cudaMemcpy2DAsync(host, (size_t)L.g.nx * sizeof(float), dev, (size_t)L.g.pitch * sizeof(float), (size_t)L.g.nx * sizeof(float),
                       (size_t)L.g.ny * L.g.nz, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// Dense inverse of the coarsest operator (fp64, host): LU with partial pivoting that skips structural
// zeros, then one pair of triangular solves per unit vector.  Replaces vnl_sparse_lu
// (mad/itkDirectSolver.hxx:81-86).
int invert_dense(std::vector<double>& A, int n, std::vector<double>& inv)
{
  std::vector<int> piv(n);
  std::vector<int> ulast(n, 0);
  for (int i = 0; i < n; ++i) {
    int last = 0;
    for (int j = 0; j < n; ++j)
      if (A[(size_t)i * n + j] != 0.0) last = j;
    ulast[i] = last;
  }
  for (int k = 0; k < n; ++k) {
    int p = k;
    double best = std::fabs(A[(size_t)k * n + k]);
    for (int i = k + 1; i < n; ++i) {
      const double v = std::fabs(A[(size_t)i * n + k]);
      if (v > best) { best = v; p = i; }
    }
    if (best == 0.0) return MADGPU_ESINGULAR;
    piv[k] = p;
    if (p != k) {
      for (int j = 0; j < n; ++j) std::swap(A[(size_t)k * n + j], A[(size_t)p * n + j]);
      std::swap(ulast[k], ulast[p]);
    }
    const double inv_p = 1.0 / A[(size_t)k * n + k];
    const int kl = ulast[k];
    for (int i = k + 1; i < n; ++i) {
      double m = A[(size_t)i * n + k];
      if (m == 0.0) continue;
      m *= inv_p;
      A[(size_t)i * n + k] = m;
      double* ri = &A[(size_t)i * n];
      const double* rk = &A[(size_t)k * n];
      for (int j = k + 1; j <= kl; ++j) ri[j] -= m * rk[j];
      if (kl > ulast[i]) ulast[i] = kl;
    }
  }
  inv.assign((size_t)n * n, 0.0);
  std::vector<double> x(n);
  for (int c = 0; c < n; ++c) {
    std::fill(x.begin(), x.end(), 0.0);
    x[c] = 1.0;
    for (int k = 0; k < n; ++k) {
      if (piv[k] != k) std::swap(x[k], x[piv[k]]);
    }
    // note: the row swaps above were applied to whole rows (including the stored multipliers), so
    // the permutation can be applied to the right-hand side up front (LAPACK getrs convention).
    int first = 0;
    while (first < n && x[first] == 0.0) ++first;
    for (int i = first + 1; i < n; ++i) {
      const double* ri = &A[(size_t)i * n];
      double s = x[i];
      for (int j = first; j < i; ++j) s -= ri[j] * x[j];
      x[i] = s;
    }
    for (int i = n - 1; i >= 0; --i) {
      const double* ri = &A[(size_t)i * n];
      double s = x[i];
      const int jl = ulast[i];
      for (int j = i + 1; j <= jl; ++j) s -= ri[j] * x[j];
      x[i] = s / ri[i];
    }
    for (int i = 0; i < n; ++i) inv[(size_t)i * n + c] = x[i];
  }
  return 0;
}

// MADGPU_SETUP_TRACE=1: where the time of a tensor set-up goes (stream-synchronised wall clock per phase, on stderr)
struct SetupTrace {
  madgpu_ctx* ctx;
  bool on;
  std::chrono::steady_clock::time_point t;
  explicit SetupTrace(madgpu_ctx* c) : ctx(c), on(getenv("MADGPU_SETUP_TRACE") != nullptr), t(std::chrono::steady_clock::now()) {}
  void mark(const char* what)
  {
    if (!on) return;
    cudaStreamSynchronize(ctx->stream);
    const auto n = std::chrono::steady_clock::now();
    fprintf(stderr, "[madgpu set-up] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
    t = n;
  }
};

// Direct solver of the coarsest grid (mad/itkDirectSolver.hxx:32-88): the operator is assembled and inverted ON THE DEVICE
// (Gauss-Jordan with partial pivoting in fp64 on [A | I]); the inverse is applied as a GEMV per coarse solve.  host = true keeps
// the round-1 path (assembly + LU on the host), which the tests compare with.
int build_coarse_solver(madgpu_ctx* ctx)
{
  Level& L = ctx->lv[ctx->nlevels - 1];
  const long long nv = (long long)L.n[0] * L.n[1] * L.n[2];
  if (nv > ctx->coarse_direct_max) {  // the dense inverse is n^2 doubles; larger coarsest grids (thin volumes) are iterated instead
    ctx->coarse_direct = false;
    ctx->ncoarse = 0;
    return 0;
  }
  const int n = (int)nv;
  if (ctx->Ainv && ctx->ncoarse != n) { cudaFree(ctx->Ainv); ctx->Ainv = nullptr; }
  if (!ctx->Ainv) CU(cudaMalloc((void**)&ctx->Ainv, (size_t)n * n * sizeof(double)));
  if (!ctx->coarse_host) {
    // Work matrix [A | I] and the elimination's ~2n launches: both persist in the context.  The launches are captured ONCE into a
    // CUDA graph and replayed per tensor -- issued one by one they depend on the host keeping up for ~1000 launches of ~6 us, and
    // on B200 boxes whose host threads get descheduled that phase was measured at anything between 6 ms and 830 ms.
    const int ld = 2 * n;
    const size_t mbytes = (size_t)n * ld * sizeof(double) + (size_t)n * sizeof(double) + 16;
    if (ctx->gjM && ctx->gj_n != n) {
      if (ctx->gj_exec) { cudaGraphExecDestroy(ctx->gj_exec); ctx->gj_exec = nullptr; }
      cudaFree(ctx->gjM);
      ctx->gjM = nullptr;
    }
    if (!ctx->gjM) { CU(cudaMalloc((void**)&ctx->gjM, mbytes)); ctx->gj_n = n; }
    double* M = ctx->gjM;
    double* colk = M + (size_t)n * ld;
    int* singular = (int*)(colk + n);
    auto enqueue = [&]() {
      cudaMemsetAsync(M, 0, mbytes, ctx->stream);
      if (ctx->dim == 3) MAD_LAUNCH((k_coarse_matrix<3>), (n + 127) / 128, 128, 0, ctx->stream, L.g, tensor_of(L), M, n, ld);
      else MAD_LAUNCH((k_coarse_matrix<2>), (n + 127) / 128, 128, 0, ctx->stream, L.g, tensor_of(L), M, n, ld);
      const dim3 eg((ld + 255) / 256, (n + 7) / 8);
      for (int k = 0; k < n; ++k) {
        MAD_LAUNCH((k_gj_pivot), 1, 1024, 0, ctx->stream, M, n, ld, k, colk, singular);
        MAD_LAUNCH((k_gj_eliminate), eg, 256, 0, ctx->stream, M, n, ld, k, (const double*)colk);
      }
      cudaMemcpy2DAsync(ctx->Ainv, (size_t)n * sizeof(double), M + n, (size_t)ld * sizeof(double), (size_t)n * sizeof(double), n, cudaMemcpyDeviceToDevice, ctx->stream);
    };
    bool launched = false;
    if (ctx->graph_voxels > 0 && !ctx->gj_graph_bad) {
      if (!ctx->gj_exec) {
        cudaGraph_t graph = nullptr;
        bool ok = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
          enqueue();
          ok = cudaStreamEndCapture(ctx->stream, &graph) == cudaSuccess && graph != nullptr;
        }
        if (ok) ok = cudaGraphInstantiate(&ctx->gj_exec, graph, 0) == cudaSuccess;
        if (graph) cudaGraphDestroy(graph);
        if (!ok) { cudaGetLastError(); ctx->gj_exec = nullptr; ctx->gj_graph_bad = true; }
      }
      if (ctx->gj_exec) {
        if (cudaGraphLaunch(ctx->gj_exec, ctx->stream) == cudaSuccess) launched = true;
        else { cudaGetLastError(); ctx->gj_graph_bad = true; }
      }
    }
    if (!launched) enqueue();
    ctx->launches += 2 * n + 1;
    int sing = 0;
    cudaMemcpyAsync(ctx->h_scalar + 3, singular, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    memcpy(&sing, ctx->h_scalar + 3, sizeof(int));
    if (e != cudaSuccess) return fail(ctx, MADGPU_ECUDA, "coarsest-grid inverse: %s", cudaGetErrorString(e));
    if (sing) return fail(ctx, MADGPU_ESINGULAR, "coarsest-grid operator (%d unknowns) is singular", n);
  } else {
    // bring the coarsest tensor planes to the host (pitched layout incl. ghost planes kept)
    std::vector<std::vector<float>> hD(ctx->ncomp, std::vector<float>(L.elems));
    Geom g = L.g;
    Tensor T;
    for (int c = 0; c < 6; ++c) T.p[c] = nullptr;
    for (int c = 0; c < ctx->ncomp; ++c) {
      CU(cudaMemcpyAsync(hD[c].data(), L.D[c] - g.plane, L.elems * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
      T.p[c] = hD[c].data() + g.plane;
    }
    CU(cudaStreamSynchronize(ctx->stream));
    std::vector<double> A((size_t)n * n, 0.0), inv;
    const int zl = ctx->dim == 3 ? -1 : 0, zh = ctx->dim == 3 ? 1 : 0;
    for (int z = 0; z < g.nz; ++z)
      for (int y = 0; y < g.ny; ++y)
        for (int x = 0; x < g.nx; ++x) {
          double S[27];
          Row<double> r;
          if (ctx->dim == 3) { row_coeffs<3, double>(g, T, x, y, z, r); scatter_row<3, double>(g, r, x, y, z, S); }
          else { row_coeffs<2, double>(g, T, x, y, z, r); scatter_row<2, double>(g, r, x, y, z, S); }
          const size_t row = ((size_t)z * g.ny + y) * g.nx + x;  // LexPosition, mad/itkDirectSolver.h:89-99
          for (int oz = zl; oz <= zh; ++oz)
            for (int oy = -1; oy <= 1; ++oy)
              for (int ox = -1; ox <= 1; ++ox) {
                const int xx = x + ox, yy = y + oy, zz = z + oz;
                if (xx < 0 || xx >= g.nx || yy < 0 || yy >= g.ny || zz < 0 || zz >= g.nz) continue;
                const int si = ctx->dim == 2 ? (oy + 1) * 3 + (ox + 1) : ((oz + 1) * 3 + (oy + 1)) * 3 + (ox + 1);
                A[row * n + ((size_t)zz * g.ny + yy) * g.nx + xx] = S[si];
              }
        }
    const int rc = invert_dense(A, n, inv);
    if (rc != 0) return fail(ctx, rc, "coarsest-grid operator (%d unknowns) is singular", n);
    CU(cudaMemcpyAsync(ctx->Ainv, inv.data(), (size_t)n * n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  ctx->ncoarse = n;
  ctx->coarse_direct = true;
  if ((size_t)n * sizeof(double) > 48 * 1024)
    CU(cudaFuncSetAttribute(k_coarse_gemv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(n * sizeof(double))));
  return 0;
}

// After level-0 tensor planes are filled: restrict them down the hierarchy and build the coarse solver.
int finish_tensor(madgpu_ctx* ctx)
{
  for (int l = 0; l < ctx->nlevels; ++l) { ctx->lv[l].coef16_valid = false; ctx->lv[l].coef16_off = false; }
  SetupTrace tr(ctx);
  drop_graphs(ctx);  // the captured cycles read the packed rows of the previous tensor
  tr.mark("drop graphs");
  for (int l = 0; l + 1 < ctx->nlevels; ++l)
    for (int c = 0; c < ctx->ncomp; ++c)  // mad/itkGridsHierarchy.hxx:149-162
      op_restrict<float>(ctx, l, ctx->lv[l].D[c], ctx->lv[l + 1].D[c], MADGPU_K_MISC);
  CU(cudaGetLastError());
  tr.mark("tensor restriction");
  if (ctx->world > 1) {
    // tensor of the agglomeration level -> rank 0, which finishes the hierarchy below it
    Level& A = ctx->lv[ctx->nlevels - 1];
    for (int c = 0; c < ctx->ncomp; ++c) {
      pack_slab(ctx, A, A.D[c], ctx->slab_buf);
      gather_slabs(ctx, A);
      if (ctx->rank == 0) unpack_slab(ctx->sub, ctx->sub->lv[0], ctx->gather_buf, ctx->sub->lv[0].D[c]);
    }
    double flag = 0.0;
    if (ctx->rank == 0) {
      const int rc = finish_tensor(ctx->sub);
      if (rc != 0) { ctx->err = ctx->sub->err; flag = (double)rc; }
    }
    // every rank must learn whether rank 0 could build the coarse hierarchy
    CU(cudaMemcpyAsync(ctx->d_scalar, &flag, sizeof flag, cudaMemcpyHostToDevice, ctx->stream));
    NCV(g_nccl.Broadcast(ctx->d_scalar, ctx->d_scalar, 1, Nccl::Float64, 0, ctx->comm, ctx->stream));
    const double got = read_scalar(ctx);
    if (!ctx->sticky.empty()) return fail(ctx, MADGPU_ECUDA, "%s", ctx->sticky.c_str());
    if (got != 0.0) return ctx->rank == 0 ? (int)got : fail(ctx, (int)got, "rank 0 could not build the agglomerated coarse hierarchy");
    ctx->tensor_set = true;
    return 0;
  }
  const int rc = build_coarse_solver(ctx);
  tr.mark("coarsest-grid inverse");
  if (rc != 0) return rc;
  ctx->tensor_set = true;
  return 0;
}

int ensure_stage(madgpu_ctx* ctx, int i, size_t bytes)
{
  if (ctx->stage_bytes[i] >= bytes) return 0;
  if (ctx->stage[i]) { CU(cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->stage[i]); ctx->stage[i] = nullptr; ctx->stage_bytes[i] = 0; }
  CU(cudaMalloc(&ctx->stage[i], bytes));
  ctx->stage_bytes[i] = bytes;
  return 0;
}

template <typename T>
int set_tensor_host(madgpu_ctx* ctx, const T* aos)
{
  if (!ctx || !aos) return fail(ctx, MADGPU_EINVAL, "null argument");
  const auto t0 = std::chrono::steady_clock::now();
  CU(cudaSetDevice(ctx->p.device));
  Level& L = ctx->lv[0];
  const long long nvox = (long long)L.n[0] * L.n[1] * L.n[2];
  const long long chunk = 4ll << 20;  // voxels per staging chunk
  const int nc = ctx->ncomp;
  T* stage[2] = {nullptr, nullptr};
  cudaEvent_t done[2];
  const long long cap = std::min(chunk, nvox);
  for (int i = 0; i < 2; ++i) {
    const int rc = ensure_stage(ctx, i, (size_t)cap * nc * sizeof(T));
    if (rc) return rc;
    stage[i] = (T*)ctx->stage[i];
    CU(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
  }
  SetupTrace tr(ctx);
  tr.mark("staging buffers");
  int k = 0;
  for (long long first = 0; first < nvox; first += chunk, k ^= 1) {
    const long long cnt = std::min(chunk, nvox - first);
    CU(cudaEventSynchronize(done[k]));
    CU(cudaMemcpyAsync(stage[k], aos + first * nc, (size_t)cnt * nc * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    const int th = 256;
    const unsigned bl = (unsigned)((cnt + th - 1) / th);
    if (nc == 6) MAD_LAUNCH((k_tensor_ingest<T, 6>), bl, th, 0, ctx->stream, L.g, stage[k], first, cnt, nullptr, L.D[0], L.D[1], L.D[2], L.D[3], L.D[4], L.D[5]);
    else MAD_LAUNCH((k_tensor_ingest<T, 3>), bl, th, 0, ctx->stream, L.g, stage[k], first, cnt, nullptr, L.D[0], L.D[1], L.D[2], nullptr, nullptr, nullptr);
    CU(cudaEventRecord(done[k], ctx->stream));
  }
  CU(cudaStreamSynchronize(ctx->stream));
  tr.mark("upload + ingest");
  for (int i = 0; i < 2; ++i) cudaEventDestroy(done[i]);
  const int rc = finish_tensor(ctx);
  ctx->st.setup_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return rc;
}

size_t pix_size(int t) { return t == MADGPU_PIX_U8 ? 1 : t == MADGPU_PIX_I16 ? 2 : t == MADGPU_PIX_F32 ? 4 : 8; }

int stage_input(madgpu_ctx* ctx, int type, const void* dev_dense)
{
  ctx->have_result = false;  // f64 is about to hold an image, not a result
  Level& L = ctx->lv[0];
  const dim3 b = block3(ctx->dim), g = grid3(L.g, b);
  switch (type) {
    case MADGPU_PIX_U8: MAD_LAUNCH((k_dense_to_pitched<uint8_t, double>), g, b, 0, ctx->stream, L.g, (const uint8_t*)dev_dense, ctx->f64); break;
    case MADGPU_PIX_I16: MAD_LAUNCH((k_dense_to_pitched<int16_t, double>), g, b, 0, ctx->stream, L.g, (const int16_t*)dev_dense, ctx->f64); break;
    case MADGPU_PIX_F32: MAD_LAUNCH((k_dense_to_pitched<float, double>), g, b, 0, ctx->stream, L.g, (const float*)dev_dense, ctx->f64); break;
    case MADGPU_PIX_F64: MAD_LAUNCH((k_dense_to_pitched<double, double>), g, b, 0, ctx->stream, L.g, (const double*)dev_dense, ctx->f64); break;
    default: return fail(ctx, MADGPU_EINVAL, "bad pixel type %d", type);
  }
  ctx->launches++;
  return 0;
}

int stage_output(madgpu_ctx* ctx, int type, void* dev_dense)
{
  Level& L = ctx->lv[0];
  const dim3 b = block3(ctx->dim), g = grid3(L.g, b);
  switch (type) {
    case MADGPU_PIX_U8: MAD_LAUNCH((k_pitched_to_dense<double, uint8_t>), g, b, 0, ctx->stream, L.g, ctx->u64, (uint8_t*)dev_dense); break;
    case MADGPU_PIX_I16: MAD_LAUNCH((k_pitched_to_dense<double, int16_t>), g, b, 0, ctx->stream, L.g, ctx->u64, (int16_t*)dev_dense); break;
    case MADGPU_PIX_F32: MAD_LAUNCH((k_pitched_to_dense<double, float>), g, b, 0, ctx->stream, L.g, ctx->u64, (float*)dev_dense); break;
    case MADGPU_PIX_F64: MAD_LAUNCH((k_pitched_to_dense<double, double>), g, b, 0, ctx->stream, L.g, ctx->u64, (double*)dev_dense); break;
    default: return fail(ctx, MADGPU_EINVAL, "bad pixel type %d", type);
  }
  ctx->launches++;
  return 0;
}

// The time-step loop of GenerateData (…Filter.hxx:158-263) on f64 (already staged) -> u64.
int run_steps(madgpu_ctx* ctx)
{
  Level& L = ctx->lv[0];
  const madgpu_params& P = ctx->p;
  const size_t bytes64 = (size_t)L.g.plane * L.g.nz * sizeof(double);
  const dim3 b = block3(ctx->dim), g = grid3(L.g, b);
  ctx->relres_hist.assign((size_t)std::max(P.number_of_steps, 1) * std::max(P.max_cycles, 1), NAN);
  ctx->st.steps = 0;
  ctx->st.total_cycles = 0;
  float fmg_ms_total = 0.f, solve_ms_total = 0.f;
  for (int n = 0; n < P.number_of_steps; ++n) {
    CU(cudaEventRecord(ctx->ev_a, ctx->stream));
    if (P.cycle == MADGPU_CYCLE_FMG) fmg(ctx);                                                        // :174
    else { Scope s(ctx, MADGPU_K_MISC); halo_dirty(ctx, ctx->u64); CU(cudaMemcpyAsync(ctx->u64, ctx->f64, bytes64, cudaMemcpyDeviceToDevice, ctx->stream)); }  // :182-199
    CU(cudaEventRecord(ctx->ev_b, ctx->stream));
    // rhsNorm (:204)
    {
      Scope s(ctx, MADGPU_K_MISC, 2);
      MAD_LAUNCH((k_sumsq<double>), g, b, 0, ctx->stream, L.g, ctx->f64, ctx->partials);
      reduce_partials(ctx, (size_t)g.x * g.y * g.z);
    }
    const double rhs_norm = std::sqrt(read_scalar(ctx));
    op_residual64(ctx, L.f, nullptr);
    double relres = 0.0;
    int it = 0;
    do {                                                                                               // :207-246
      outer_iteration(ctx, P.cycle == MADGPU_CYCLE_SMOOTHER);
      relres = std::sqrt(read_scalar(ctx)) / rhs_norm;
      if (!ctx->sticky.empty()) return fail(ctx, MADGPU_ECUDA, "%s", ctx->sticky.c_str());  // e.g. a halo wait timed out: every rank sees it in the same cycle
      // a NaN would end the do-while below ("relres > tolerance" is false) and hand back a NaN image with rc 0.  (A zero image
      // keeps the reference's behaviour: 0/0 at …Filter.hxx:204,217 ends its loop after one cycle and the zero image comes back.)
      if (!std::isfinite(relres) && rhs_norm > 0.0) return fail(ctx, MADGPU_ENUMERIC, "time step %d, cycle %d: the relative residual is not finite (%g) -- diverged or overflowed", n, it + 1, relres);
      ctx->relres_hist[(size_t)n * P.max_cycles + it] = relres;
      if (P.verbose && ctx->rank == 0) {
        if (P.cycle == MADGPU_CYCLE_SMOOTHER) printf("Smoother iteration n. %d: relative residual = %g\n", it + 1, relres);
        else printf("|--- VCycle n. %d ---| relative residual = %g\n", it + 1, relres);
      }
      ++it;
    } while (relres > P.tolerance && it < P.max_cycles);
    CU(cudaEventRecord(ctx->ev_c, ctx->stream));
    { Scope s(ctx, MADGPU_K_MISC); CU(cudaMemcpyAsync(ctx->f64, ctx->u64, bytes64, cudaMemcpyDeviceToDevice, ctx->stream)); }  // :248-261
    CU(cudaEventSynchronize(ctx->ev_c));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b); fmg_ms_total += ms;
    cudaEventElapsedTime(&ms, ctx->ev_b, ctx->ev_c); solve_ms_total += ms;
    if (n < MADGPU_MAX_STEPS) { ctx->st.cycles_per_step[n] = it; ctx->st.final_relres[n] = relres; }
    ctx->st.steps = n + 1;
    ctx->st.total_cycles += it;
  }
  ctx->st.fmg_ms = fmg_ms_total;
  ctx->st.solve_ms = solve_ms_total;
  CU(cudaGetLastError());
  if (!ctx->sticky.empty()) return fail(ctx, MADGPU_ECUDA, "%s", ctx->sticky.c_str());
  ctx->have_result = P.number_of_steps > 0;  // f64 == u64 == the result (fp64)
  return 0;
}

void begin_stats(madgpu_ctx* ctx)
{
  const double setup = ctx->st.setup_ms;
  memset(&ctx->st, 0, sizeof ctx->st);
  ctx->st.struct_size = (int32_t)sizeof(madgpu_stats);
  ctx->st.setup_ms = setup;
  ctx->st.levels = ctx->total_levels;
  ctx->launches = 0;
}

void end_stats(madgpu_ctx* ctx, madgpu_stats* out)
{
  prof_collect(ctx);
  ctx->st.kernel_launches = ctx->launches;
  if (out) {
    const size_t n = std::min((size_t)(out->struct_size > 0 ? out->struct_size : (int32_t)sizeof(madgpu_stats)), sizeof(madgpu_stats));
    memcpy(out, &ctx->st, n);
  }
}

int check_ready(madgpu_ctx* ctx)
{
  if (!ctx) return MADGPU_EINVAL;
  if (!ctx->tensor_set) return fail(ctx, MADGPU_ESTATE, "diffusion tensor not set (call madgpu_set_tensor_* first)");
  if (ctx->p.number_of_steps < 1) return fail(ctx, MADGPU_EINVAL, "number_of_steps must be >= 1");
  if (!ctx->sticky.empty()) return fail(ctx, MADGPU_ECUDA, "%s", ctx->sticky.c_str());
  return 0;
}

}  // namespace

// ============================================================================================
//                                         C-ABI
// ============================================================================================
extern "C" {

void madgpu_params_default(madgpu_params* p)
{
  memset(p, 0, sizeof *p);
  p->struct_size = (int32_t)sizeof *p;
  p->dim = 3;
  p->spacing[0] = p->spacing[1] = p->spacing[2] = 1.0;
  p->time_step = 0.01;          // …Filter.hxx:39
  p->number_of_steps = 1;       // :40
  p->cycle = MADGPU_CYCLE_V;    // :41
  p->iterations_per_grid = 2;   // :42
  p->tolerance = 1e-6;          // :43
  p->max_cycles = 100;          // :44
  p->verbose = 0;               // :45
  p->smoother = MADGPU_SMOOTHER_GS;
  p->omega = 2.0 / 3.0;
  p->gs_colors = 4;
  p->device = 0;
  p->rank = 0;
  p->world_size = 1;
}

const char* madgpu_last_error(const madgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

static int create_ctx(const madgpu_params* p, const void* nccl_id, cudaStream_t shared_stream, madgpu_ctx** out);

int madgpu_create(const madgpu_params* p, madgpu_ctx** out)
{
  if (p && p->struct_size == (int32_t)sizeof(madgpu_params) && p->world_size != 1)
    return fail(nullptr, MADGPU_EINVAL, "world_size %d: use madgpu_create_slab (needs the NCCL unique id)", p->world_size);
  return create_ctx(p, nullptr, nullptr, out);
}

int madgpu_nccl_unique_id(void* id128)
{
  if (!id128) return MADGPU_EINVAL;
  const char* e = nccl_load();
  if (e) return fail(nullptr, MADGPU_ECUDA, "%s", e);
  Nccl::UniqueId id;
  const int r = g_nccl.GetUniqueId(&id);
  if (r != 0) return fail(nullptr, MADGPU_ECUDA, "ncclGetUniqueId: %s", g_nccl.GetErrorString(r));
  memcpy(id128, &id, sizeof id);
  return 0;
}

int madgpu_create_slab(const madgpu_params* p, const void* nccl_unique_id, madgpu_ctx** out)
{
  if (!p || !out || !nccl_unique_id) return fail(nullptr, MADGPU_EINVAL, "null argument");
  if (p->struct_size != (int32_t)sizeof(madgpu_params)) return fail(nullptr, MADGPU_EINVAL, "madgpu_params size mismatch");
  if (p->world_size < 1 || p->rank < 0 || p->rank >= p->world_size) return fail(nullptr, MADGPU_EINVAL, "bad rank %d / world_size %d", p->rank, p->world_size);
  if (p->world_size == 1) return create_ctx(p, nullptr, nullptr, out);
  if (p->dim != 3) return fail(nullptr, MADGPU_EINVAL, "z-slab decomposition needs a 3-D volume");
  return create_ctx(p, nccl_unique_id, nullptr, out);
}

static int create_ctx(const madgpu_params* p, const void* nccl_id, cudaStream_t shared_stream, madgpu_ctx** out)
{
  madgpu_ctx* ctx = nullptr;  // errors before the context exists go to the thread-local slot
  if (!p || !out) return fail(ctx, MADGPU_EINVAL, "null argument");
  *out = nullptr;
  if (p->struct_size != (int32_t)sizeof(madgpu_params)) return fail(ctx, MADGPU_EINVAL, "madgpu_params size mismatch (%d vs %zu)", p->struct_size, sizeof(madgpu_params));
  if (p->dim != 2 && p->dim != 3) return fail(ctx, MADGPU_EINVAL, "dim must be 2 or 3");
  for (int d = 0; d < p->dim; ++d) {
    if (p->size[d] < 3) return fail(ctx, MADGPU_EINVAL, "size[%d]=%d: at least 3 voxels per axis are required", d, p->size[d]);
    if (!(p->spacing[d] > 0)) return fail(ctx, MADGPU_EINVAL, "spacing[%d] must be positive", d);
  }
  if (!(p->time_step > 0)) return fail(ctx, MADGPU_EINVAL, "time_step must be positive");
  if (p->smoother != MADGPU_SMOOTHER_GS && p->smoother != MADGPU_SMOOTHER_WJ) return fail(ctx, MADGPU_EINVAL, "unknown smoother %d", p->smoother);
  if (p->cycle < 0 || p->cycle > 2) return fail(ctx, MADGPU_EINVAL, "unknown cycle %d", p->cycle);
  if (p->max_cycles < 1) return fail(ctx, MADGPU_EINVAL, "max_cycles must be >= 1");
  const int world = nccl_id ? p->world_size : 1, rank = nccl_id ? p->rank : 0;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return fail(ctx, MADGPU_ECUDA, "no CUDA device: %s (libmadgpu has no CPU fallback)", cudaGetErrorString(e));
  if (p->device < 0 || p->device >= ndev) return fail(ctx, MADGPU_EINVAL, "device %d out of range (%d devices)", p->device, ndev);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, p->device)) != cudaSuccess) return fail(ctx, MADGPU_ECUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10) return fail(ctx, MADGPU_ECUDA, "device %d is sm_%d%d; libmadgpu is built for sm_100a only", p->device, prop.major, prop.minor);

  ctx = new (std::nothrow) madgpu_ctx();
  if (!ctx) return fail(nullptr, MADGPU_ENOMEM, "out of host memory");
  ctx->p = *p;
  ctx->dim = p->dim;
  ctx->ncomp = p->dim == 2 ? 3 : 6;
  ctx->Ainv = nullptr; ctx->ncoarse = 0; ctx->coarse_direct = false;
  ctx->have_result = false;
  ctx->gjM = nullptr; ctx->gj_n = 0; ctx->gj_exec = nullptr; ctx->gj_graph_bad = false;
  ctx->partials = nullptr; ctx->d_scalar = nullptr; ctx->h_scalar = nullptr;
  ctx->tensor_set = false; ctx->profiling = 0; ctx->rhs_norm = 0; ctx->launches = 0;
  ctx->flags = ctx->flags_lo = ctx->flags_hi = nullptr; ctx->halo_seq = 0; ctx->p2p = false;
  ctx->stage[0] = ctx->stage[1] = nullptr; ctx->stage_bytes[0] = ctx->stage_bytes[1] = 0;
  ctx->rank = rank; ctx->world = world; ctx->comm = nullptr; ctx->sub = nullptr; ctx->gather_buf = nullptr; ctx->slab_buf = nullptr;
  ctx->borrowed_stream = shared_stream != nullptr;
  {
    const char* e = getenv("MADGPU_FAST_MIN_NX");  // test hook: 0 forces the streaming kernels on every 3-D level
    ctx->fast_min_nx = e ? std::max(atoi(e), 8) : 64;
    e = getenv("MADGPU_FAST2D");
    ctx->fast2d = e ? atoi(e) : 1;
    e = getenv("MADGPU_FAST2D_MIN_PIXELS");
    ctx->fast2d_min_pixels = e ? atoll(e) : (1ll << 20);
    e = getenv("MADGPU_PF_DIST");
    ctx->pf_dist = e ? atoi(e) : 2;
    e = getenv("MADGPU_GS_PAIRS");
    ctx->gs_pairs = e ? atoi(e) : 1;
    e = getenv("MADGPU_GS_COEF16");
    ctx->gs_coef16 = e ? atoi(e) : 1;
    e = getenv("MADGPU_GS_FUSED");
    ctx->gs_fused = e ? atoi(e) : 1;
    e = getenv("MADGPU_P2P_WAIT");
    ctx->p2p_wait_kernel = !(e && !strcmp(e, "memop"));  // default: bounded k_halo_wait (verified on 2 and 8 B200); "memop": cuStreamWaitValue32
    e = getenv("MADGPU_P2P_TIMEOUT_MS");
    ctx->p2p_timeout_cycles = (long long)((e ? atof(e) : 10000.0) * 2.0e6);
    ctx->p2p_drop_rank = ctx->p2p_drop_seq = -1;
    if ((e = getenv("MADGPU_P2P_TEST_DROP_SIGNAL"))) sscanf(e, "%d:%d", &ctx->p2p_drop_rank, &ctx->p2p_drop_seq);
    e = getenv("MADGPU_RES64_COEF32");
    ctx->res64_c32 = e ? atoi(e) : 1;
    e = getenv("MADGPU_RES64_SMEM");
    ctx->res64_smem = e ? atoi(e) : 0;
    e = getenv("MADGPU_RES64_MINB");
    ctx->res64_minb = e ? atoi(e) : 2;
    e = getenv("MADGPU_FAST_CFG");
    ctx->fast_cfg = e ? atoi(e) : 0;
    e = getenv("MADGPU_GRAPH_VOXELS");
    ctx->graph_voxels = e ? atoll(e) : 128ll * 128 * 128;
    ctx->capturing = false;
    e = getenv("MADGPU_PROLONG_CELL");
    ctx->prolong_cell = e ? atoi(e) : 1;
    e = getenv("MADGPU_RESTRICT_CELL");
    ctx->restrict_cell = e ? atoi(e) : 1;
    e = getenv("MADGPU_GS_TB");
    ctx->gs_tb = e ? std::min(std::max(atoi(e), 1), 3) : 1;
    e = getenv("MADGPU_GS_PRIVATE");
    ctx->gs_private = e ? atoi(e) : 2;
    e = getenv("MADGPU_GS_TB_SINGLE");
    ctx->gs_tb_single = e ? atoi(e) : 0;
    e = getenv("MADGPU_COARSE_HOST");
    ctx->coarse_host = e ? atoi(e) : 0;
    e = getenv("MADGPU_COARSE_DIRECT_MAX");
    ctx->coarse_direct_max = e ? atoll(e) : 4096;
  }
  ctx->u64 = ctx->f64 = nullptr;
  memset(&ctx->st, 0, sizeof ctx->st);
  auto bail = [&](int rc) { g_create_error = ctx->err; madgpu_destroy(ctx); return rc; };
#define CUB(call)                                                                                              \
  do {                                                                                                         \
    cudaError_t e_ = (call);                                                                                   \
    if (e_ != cudaSuccess) {                                                                                   \
      fail(ctx, e_ == cudaErrorMemoryAllocation ? MADGPU_ENOMEM : MADGPU_ECUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
      return bail(e_ == cudaErrorMemoryAllocation ? MADGPU_ENOMEM : MADGPU_ECUDA);                             \
    }                                                                                                          \
  } while (0)
  CUB(cudaSetDevice(p->device));
  if (shared_stream) ctx->stream = shared_stream;
  else CUB(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  CUB(cudaEventCreate(&ctx->ev_a));
  CUB(cudaEventCreate(&ctx->ev_b));
  CUB(cudaEventCreate(&ctx->ev_c));

  int sizes[MADGPU_MAX_LEVELS][3], cent[MADGPU_MAX_LEVELS][3];
  ctx->nlevels = level_schedule(p->dim, p->size, sizes, cent);
  if (ctx->nlevels < 1) { fail(ctx, MADGPU_EINVAL, "too many levels"); return bail(MADGPU_EINVAL); }
  ctx->total_levels = ctx->nlevels;
  memcpy(ctx->gsize, sizes, sizeof sizes);
  memcpy(ctx->gcent, cent, sizeof cent);
  int La = -1;
  if (world > 1) {
    std::string why;
    La = plan_slabs(ctx->nlevels, sizes, world, why);
    if (La < 1) { fail(ctx, MADGPU_EINVAL, "cannot cut %dx%dx%d into %d z-slabs: %s", p->size[0], p->size[1], p->size[2], world, why.c_str()); return bail(MADGPU_EINVAL); }
    const char* ne = nccl_load();
    if (ne) { fail(ctx, MADGPU_ECUDA, "%s", ne); return bail(MADGPU_ECUDA); }
    Nccl::UniqueId id;
    memcpy(&id, nccl_id, sizeof id);
    const int r = g_nccl.CommInitRank(&ctx->comm, world, id, rank);
    if (r != 0) { fail(ctx, MADGPU_ECUDA, "ncclCommInitRank: %s", g_nccl.GetErrorString(r)); ctx->comm = nullptr; return bail(MADGPU_ECUDA); }
    ctx->nlevels = La + 1;  // this context: the distributed levels and the agglomeration level
  }
  size_t max_blocks = 0;
  for (int l = 0; l < ctx->nlevels; ++l) {
    Level& L = ctx->lv[l];
    for (int d = 0; d < 3; ++d) {
      L.n[d] = sizes[l][d];
      L.cent[d] = cent[l][d];
      L.h[d] = d < p->dim ? p->spacing[d] * (double)(1u << l) : 1.0;  // mad/itkGridsHierarchy.hxx:80
    }
    L.gnz = L.n[2];
    L.zb = 0;
    if (world > 1) {  // this rank's slab of the level
      L.n[2] = L.gnz / world;
      L.zb = rank * L.n[2];
    }
    fill_geom(L, p->dim, p->time_step);
    if (world > 1) {
      L.g.zlo_phys = rank == 0;
      L.g.zhi_phys = rank == world - 1;
      L.g.z0 = L.zb;
    }
    float** fields[3] = {&L.u, &L.f, &L.tmp};
    for (auto f : fields) {
      int rc = dalloc(ctx, L.allocs, f, L.elems);
      if (rc) return bail(rc);
      *f += L.g.plane;
    }
    L.coef16 = nullptr;
    L.coef16_valid = false;
    L.coef16_off = false;
    L.tb_flip = 0;
    for (int c = 0; c < 6; ++c) L.D[c] = nullptr;
    for (int c = 0; c < ctx->ncomp; ++c) {
      int rc = dalloc(ctx, L.allocs, &L.D[c], L.elems);
      if (rc) return bail(rc);
      L.D[c] += L.g.plane;
    }
    const dim3 b = block3(p->dim), g = grid3(L.g, b);
    max_blocks = std::max(max_blocks, (size_t)g.x * g.y * g.z);
  }
  {
    Level& L = ctx->lv[0];
    int rc = dalloc(ctx, ctx->allocs, &ctx->u64, L.elems);
    if (rc) return bail(rc);
    rc = dalloc(ctx, ctx->allocs, &ctx->f64, L.elems);
    if (rc) return bail(rc);
    ctx->u64 += L.g.plane;
    ctx->f64 += L.g.plane;
  }
  ctx->npartials = max_blocks;
  CUB(cudaMalloc((void**)&ctx->partials, max_blocks * sizeof(double)));
  CUB(cudaMalloc((void**)&ctx->d_scalar, 8 * sizeof(double)));
  CUB(cudaMemsetAsync(ctx->d_scalar, 0, 8 * sizeof(double), ctx->stream));
  CUB(cudaMallocHost((void**)&ctx->h_scalar, 8 * sizeof(double)));
  if (world > 1) {
    // agglomeration level: dense staging of the local slab, and on rank 0 the gathered level plus the serial
    // sub-hierarchy below it (a context of its own on the same stream)
    const Level& A = ctx->lv[La];
    const size_t slab = (size_t)A.n[0] * A.n[1] * A.n[2];
    CUB(cudaMalloc((void**)&ctx->slab_buf, slab * sizeof(float)));
    if (rank == 0) {
      CUB(cudaMalloc((void**)&ctx->gather_buf, slab * world * sizeof(float)));
      madgpu_params sp = *p;
      sp.world_size = 1;
      sp.rank = 0;
      for (int d = 0; d < 3; ++d) { sp.size[d] = sizes[La][d]; sp.spacing[d] = p->spacing[d] * (double)(1u << La); }
      const int rc = create_ctx(&sp, nullptr, ctx->stream, &ctx->sub);
      if (rc) { ctx->err = g_create_error; return bail(rc); }
      if (ctx->sub->nlevels != ctx->total_levels - La) { fail(ctx, MADGPU_ESTATE, "internal: sub-hierarchy depth mismatch"); return bail(MADGPU_ESTATE); }
    }
  }
  CUB(cudaStreamSynchronize(ctx->stream));
#undef CUB
  *out = ctx;
  return MADGPU_OK;
}

void madgpu_destroy(madgpu_ctx* ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->p.device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (auto& g : ctx->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  for (int l = 0; l < MADGPU_MAX_LEVELS; ++l)
    for (void* p : ctx->lv[l].allocs) cudaFree(p);
  for (void* p : ctx->allocs) cudaFree(p);
  if (ctx->gj_exec) cudaGraphExecDestroy(ctx->gj_exec);
  if (ctx->gjM) cudaFree(ctx->gjM);
  if (ctx->Ainv) cudaFree(ctx->Ainv);
  if (ctx->partials) cudaFree(ctx->partials);
  if (ctx->d_scalar) cudaFree(ctx->d_scalar);
  if (ctx->h_scalar) cudaFreeHost(ctx->h_scalar);
  for (auto& pe : ctx->prof) { cudaEventDestroy(pe.e0); cudaEventDestroy(pe.e1); }
  for (auto e : ctx->ev_pool) cudaEventDestroy(e);
  if (ctx->ev_a) cudaEventDestroy(ctx->ev_a);
  if (ctx->ev_b) cudaEventDestroy(ctx->ev_b);
  if (ctx->ev_c) cudaEventDestroy(ctx->ev_c);
  for (auto& sh : ctx->shared) {
    if (sh.lo) cudaIpcCloseMemHandle(sh.lo);
    if (sh.hi) cudaIpcCloseMemHandle(sh.hi);
  }
  if (ctx->flags) cudaFree(ctx->flags);
  for (int i = 0; i < 2; ++i) if (ctx->stage[i]) cudaFree(ctx->stage[i]);
  if (ctx->sub) madgpu_destroy(ctx->sub);
  if (ctx->gather_buf) cudaFree(ctx->gather_buf);
  if (ctx->slab_buf) cudaFree(ctx->slab_buf);
  if (ctx->comm) g_nccl.CommDestroy(ctx->comm);
  if (ctx->stream && !ctx->borrowed_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int madgpu_set_solver(madgpu_ctx* ctx, int32_t smoother, double omega, int32_t iterations_per_grid, int32_t cycle, double tolerance,
                      int32_t max_cycles, int32_t number_of_steps, int32_t verbose)
{
  if (!ctx) return MADGPU_EINVAL;
  if (smoother != MADGPU_SMOOTHER_GS && smoother != MADGPU_SMOOTHER_WJ) return fail(ctx, MADGPU_EINVAL, "unknown smoother %d", smoother);
  if (cycle < 0 || cycle > 2) return fail(ctx, MADGPU_EINVAL, "unknown cycle %d", cycle);
  if (max_cycles < 1 || number_of_steps < 1 || iterations_per_grid < 0) return fail(ctx, MADGPU_EINVAL, "bad iteration counts");
  ctx->p.smoother = smoother; ctx->p.omega = omega; ctx->p.iterations_per_grid = iterations_per_grid; ctx->p.cycle = cycle;
  ctx->p.tolerance = tolerance; ctx->p.max_cycles = max_cycles; ctx->p.number_of_steps = number_of_steps; ctx->p.verbose = verbose;
  if (ctx->sub) return madgpu_set_solver(ctx->sub, smoother, omega, iterations_per_grid, MADGPU_CYCLE_V, tolerance, max_cycles, number_of_steps, 0);
  return 0;
}

int madgpu_set_tensor_f32(madgpu_ctx* ctx, const float* aos) { return set_tensor_host<float>(ctx, aos); }
int madgpu_set_tensor_f64(madgpu_ctx* ctx, const double* aos) { return set_tensor_host<double>(ctx, aos); }

int madgpu_set_tensor_device_f32(madgpu_ctx* ctx, const float* const* planes)
{
  if (!ctx || !planes) return fail(ctx, MADGPU_EINVAL, "null argument");
  const auto t0 = std::chrono::steady_clock::now();
  CU(cudaSetDevice(ctx->p.device));
  Level& L = ctx->lv[0];
  const dim3 b = block3(ctx->dim), g = grid3(L.g, b);
  for (int c = 0; c < ctx->ncomp; ++c) {
    if (!planes[c]) return fail(ctx, MADGPU_EINVAL, "null tensor plane %d", c);
    MAD_LAUNCH((k_dense_to_pitched<float, float>), g, b, 0, ctx->stream, L.g, planes[c], L.D[c]);
  }
  const int rc = finish_tensor(ctx);
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->st.setup_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return rc;
}

int madgpu_solve_cast(madgpu_ctx* ctx, int32_t in_type, const void* in, int32_t out_type, void* out, madgpu_stats* stats)
{
  int rc = check_ready(ctx);
  if (rc) return rc;
  if (!in || !out) return fail(ctx, MADGPU_EINVAL, "null image pointer");
  if (in_type < 0 || in_type > 3 || out_type < 0 || out_type > 3) return fail(ctx, MADGPU_EINVAL, "bad pixel type");
  CU(cudaSetDevice(ctx->p.device));
  begin_stats(ctx);
  Level& L = ctx->lv[0];
  const size_t nvox = (size_t)L.n[0] * L.n[1] * L.n[2];
  const size_t sb = nvox * std::max(pix_size(in_type), pix_size(out_type));
  rc = ensure_stage(ctx, 0, sb);
  if (rc) return rc;
  void* stage = ctx->stage[0];
  auto t0 = std::chrono::steady_clock::now();
  cudaError_t e = cudaMemcpyAsync(stage, in, nvox * pix_size(in_type), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) { rc = stage_input(ctx, in_type, stage); e = cudaStreamSynchronize(ctx->stream); }
  ctx->st.h2d_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (e != cudaSuccess || rc) return rc ? rc : fail(ctx, MADGPU_ECUDA, "input upload: %s", cudaGetErrorString(e));
  rc = run_steps(ctx);
  if (rc) return rc;
  t0 = std::chrono::steady_clock::now();
  rc = stage_output(ctx, out_type, stage);
  if (!rc) {
    e = cudaMemcpyAsync(out, stage, nvox * pix_size(out_type), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  }
  ctx->st.d2h_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (rc) return rc;
  if (e != cudaSuccess) return fail(ctx, MADGPU_ECUDA, "output download: %s", cudaGetErrorString(e));
  end_stats(ctx, stats);
  return 0;
}

int madgpu_solve_u8(madgpu_ctx* ctx, const uint8_t* in, uint8_t* out, madgpu_stats* s) { return madgpu_solve_cast(ctx, MADGPU_PIX_U8, in, MADGPU_PIX_U8, out, s); }
int madgpu_solve_i16(madgpu_ctx* ctx, const int16_t* in, int16_t* out, madgpu_stats* s) { return madgpu_solve_cast(ctx, MADGPU_PIX_I16, in, MADGPU_PIX_I16, out, s); }
int madgpu_solve_f32(madgpu_ctx* ctx, const float* in, float* out, madgpu_stats* s) { return madgpu_solve_cast(ctx, MADGPU_PIX_F32, in, MADGPU_PIX_F32, out, s); }
int madgpu_solve_f64(madgpu_ctx* ctx, const double* in, double* out, madgpu_stats* s) { return madgpu_solve_cast(ctx, MADGPU_PIX_F64, in, MADGPU_PIX_F64, out, s); }

int madgpu_solve_device_f32(madgpu_ctx* ctx, const float* d_in, float* d_out, madgpu_stats* stats)
{
  int rc = check_ready(ctx);
  if (rc) return rc;
  if (!d_out) return fail(ctx, MADGPU_EINVAL, "null image pointer");
  if (!d_in && !ctx->have_result) return fail(ctx, MADGPU_ESTATE, "d_in == NULL continues from the previous solve of this context, and there is none");
  CU(cudaSetDevice(ctx->p.device));
  begin_stats(ctx);
  // d_in == NULL: the right-hand side is the fp64 result of the previous solve, which run_steps left in f64 (…Filter.hxx:248-261) --
  // the carrier between the outer iterations of the VED filter, which the reference keeps in double (VED.h:64)
  if (d_in) {
    rc = stage_input(ctx, MADGPU_PIX_F32, d_in);
    if (rc) return rc;
  }
  rc = run_steps(ctx);
  if (rc) return rc;
  rc = stage_output(ctx, MADGPU_PIX_F32, d_out);
  if (rc) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  end_stats(ctx, stats);
  return 0;
}

// ---- cycle-level driving --------------------------------------------------------------------
static int cycles_begin_common(madgpu_ctx* ctx)
{
  ctx->have_result = false;  // cycle-level driving leaves the right-hand side, not the iterate, in f64
  Level& L = ctx->lv[0];
  const size_t bytes64 = (size_t)L.g.plane * L.g.nz * sizeof(double);
  const dim3 b = block3(ctx->dim), g = grid3(L.g, b);
  halo_dirty(ctx, ctx->u64);
  CU(cudaMemcpyAsync(ctx->u64, ctx->f64, bytes64, cudaMemcpyDeviceToDevice, ctx->stream));
  MAD_LAUNCH((k_sumsq<double>), g, b, 0, ctx->stream, L.g, ctx->f64, ctx->partials);
  reduce_partials(ctx, (size_t)g.x * g.y * g.z);
  ctx->rhs_norm = std::sqrt(read_scalar(ctx));
  op_residual64(ctx, L.f, nullptr);
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaGetLastError());
  return 0;
}

int madgpu_cycles_begin_device_f32(madgpu_ctx* ctx, const float* d_in)
{
  int rc = check_ready(ctx);
  if (rc) return rc;
  if (!d_in) return fail(ctx, MADGPU_EINVAL, "null image pointer");
  CU(cudaSetDevice(ctx->p.device));
  rc = stage_input(ctx, MADGPU_PIX_F32, d_in);
  if (rc) return rc;
  return cycles_begin_common(ctx);
}

int madgpu_cycles_begin_f32(madgpu_ctx* ctx, const float* in)
{
  int rc = check_ready(ctx);
  if (rc) return rc;
  if (!in) return fail(ctx, MADGPU_EINVAL, "null image pointer");
  CU(cudaSetDevice(ctx->p.device));
  Level& L = ctx->lv[0];
  const size_t nvox = (size_t)L.n[0] * L.n[1] * L.n[2];
  float* stage = nullptr;
  CU(cudaMalloc((void**)&stage, nvox * sizeof(float)));
  cudaError_t e = cudaMemcpyAsync(stage, in, nvox * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) { rc = stage_input(ctx, MADGPU_PIX_F32, stage); e = cudaStreamSynchronize(ctx->stream); }
  cudaFree(stage);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(ctx, MADGPU_ECUDA, "input upload: %s", cudaGetErrorString(e));
  return cycles_begin_common(ctx);
}

int madgpu_cycles_run(madgpu_ctx* ctx, int32_t n, double* relres, float* device_ms, madgpu_stats* stats)
{
  int rc = check_ready(ctx);
  if (rc) return rc;
  if (n < 0) return fail(ctx, MADGPU_EINVAL, "negative cycle count");
  CU(cudaSetDevice(ctx->p.device));
  begin_stats(ctx);
  CU(cudaEventRecord(ctx->ev_a, ctx->stream));
  for (int i = 0; i < n; ++i) {
    outer_iteration(ctx, ctx->p.cycle == MADGPU_CYCLE_SMOOTHER);
    const double r = std::sqrt(read_scalar(ctx)) / ctx->rhs_norm;
    if (!ctx->sticky.empty()) return fail(ctx, MADGPU_ECUDA, "%s", ctx->sticky.c_str());
    if (relres) relres[i] = r;
    if (!std::isfinite(r) && ctx->rhs_norm > 0.0) return fail(ctx, MADGPU_ENUMERIC, "cycle %d: the relative residual is not finite (%g)", i + 1, r);
  }
  CU(cudaEventRecord(ctx->ev_b, ctx->stream));
  CU(cudaEventSynchronize(ctx->ev_b));
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b));
  if (device_ms) *device_ms = ms;
  ctx->st.solve_ms = ms;
  ctx->st.total_cycles = n;
  CU(cudaGetLastError());
  end_stats(ctx, stats);
  return 0;
}

int madgpu_cycles_end_device_f32(madgpu_ctx* ctx, float* d_out)
{
  int rc = check_ready(ctx);
  if (rc) return rc;
  if (!d_out) return fail(ctx, MADGPU_EINVAL, "null image pointer");
  CU(cudaSetDevice(ctx->p.device));
  rc = stage_output(ctx, MADGPU_PIX_F32, d_out);
  if (rc) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int madgpu_cycles_end_f64(madgpu_ctx* ctx, double* out)
{
  int rc = check_ready(ctx);
  if (rc) return rc;
  if (!out) return fail(ctx, MADGPU_EINVAL, "null image pointer");
  CU(cudaSetDevice(ctx->p.device));
  Level& L = ctx->lv[0];
  const size_t w = (size_t)L.g.nx * sizeof(double), dp = (size_t)L.g.pitch * sizeof(double), hrows = (size_t)L.g.ny * L.g.nz;
  CU(cudaMemcpy2DAsync(out, w, ctx->u64, dp, w, hrows, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int madgpu_fetch_output(madgpu_ctx* ctx, int32_t out_type, void* out)
{
  int rc = check_ready(ctx);
  if (rc) return rc;
  if (!out) return fail(ctx, MADGPU_EINVAL, "null image pointer");
  if (out_type < 0 || out_type > 3) return fail(ctx, MADGPU_EINVAL, "bad pixel type");
  CU(cudaSetDevice(ctx->p.device));
  Level& L = ctx->lv[0];
  const size_t nvox = (size_t)L.n[0] * L.n[1] * L.n[2];
  rc = ensure_stage(ctx, 0, nvox * pix_size(out_type));
  if (rc) return rc;
  rc = stage_output(ctx, out_type, ctx->stage[0]);
  if (rc) return rc;
  CU(cudaMemcpyAsync(out, ctx->stage[0], nvox * pix_size(out_type), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaGetLastError());
  return 0;
}

int madgpu_get_relres_history(const madgpu_ctx* ctx, double* hist, int32_t capacity)
{
  if (!ctx || !hist) return MADGPU_EINVAL;
  const size_t n = std::min((size_t)std::max(capacity, 0), ctx->relres_hist.size());
  memcpy(hist, ctx->relres_hist.data(), n * sizeof(double));
  return (int)n;
}

int madgpu_set_profiling(madgpu_ctx* ctx, int32_t on)
{
  if (!ctx) return MADGPU_EINVAL;
  ctx->profiling = on;
  return 0;
}

int madgpu_gs_leg_plan(const madgpu_ctx* ctx, int32_t level, int32_t n_iter, int32_t* passes, int32_t capacity)
{
  if (!ctx || !passes || level < 0 || level >= ctx->nlevels || n_iter < 0) return MADGPU_EINVAL;
  const Level& L = ctx->lv[level];
  int np = 0, flip = L.tb_flip;
  for (int it = 0; it < n_iter; ) {
    if (np >= capacity) return MADGPU_EINVAL;
    int32_t* q = passes + 6 * np++;
    int32_t tile[3];
    madgpu_gs_tile(ctx, level, tile);
    const bool tb_ok = ctx->p.smoother == MADGPU_SMOOTHER_GS && use_tb(ctx, L);
    const int fuse = tb_ok ? tb_fuse(ctx, n_iter - it) : 1;
    if (tb_ok) {
      const int wp = tb_wp(ctx), zc = tb_zc(L.g, wp);
      q[0] = fuse; q[1] = fast::TX; q[2] = 2 * wp; q[3] = zc; q[4] = flip ? wp : 0; q[5] = flip ? zc / 2 : 0;
      flip ^= 1;
    } else {
      const bool alt = ctx->gs_private == 2 && tile[1] == 2;  // warp-private pairs on an alternating grid
      q[0] = 1; q[1] = tile[0]; q[2] = tile[1]; q[3] = tile[2]; q[4] = alt ? flip : 0; q[5] = 0;
      if (alt) flip ^= 1;
    }
    it += fuse;
  }
  return np;
}

int madgpu_gs_tile(const madgpu_ctx* ctx, int32_t level, int32_t tile[3])
{
  if (!ctx || !tile || level < 0 || level >= ctx->nlevels) return MADGPU_EINVAL;
  const Level& L = ctx->lv[level];
  if (ctx->p.smoother == MADGPU_SMOOTHER_GS && use_fast(ctx, L) && ctx->gs_fused) {
    const int wy = (ctx->gs_coef16 && !L.coef16_off) ? (gs_pairs(ctx, L) ? (ctx->gs_private ? 2 : 8) : 4) : (ctx->fast_cfg == 1 || ctx->fast_cfg == 5 || ctx->fast_cfg == 8) ? 8 : ctx->fast_cfg == 6 ? 2 : 4;
    tile[0] = fast::TX; tile[1] = wy; tile[2] = fast_zc(L.g, wy == 2 ? 8 : wy);  // the warp-private variant keeps the launch geometry of the row-pair kernel
  } else if (ctx->p.smoother == MADGPU_SMOOTHER_GS && use_fast2_gs(ctx, L)) {
    tile[0] = fast::TX; tile[1] = fast2_yc(L.g); tile[2] = 1;  // 2-D ordering: rows in y order, even then odd columns (mad_fast2d.cuh)
  } else {
    tile[0] = tile[1] = tile[2] = 0;  // one pass per colour over the whole level
  }
  return 0;
}

// ---- peer-memory halo set-up: every rank exports the CUDA IPC handles of its level fields and arrival counters, the
// ---- host layer hands each rank its two neighbours' blobs (any transport), import maps them.
static int build_shared_list(madgpu_ctx* ctx)
{
  if (!ctx->shared.empty()) return 0;
  for (int l = 0; l < ctx->nlevels; ++l) {
    Level& L = ctx->lv[l];
    for (int i = 0; i < 3; ++i) ctx->shared.push_back({L.allocs[i], L.elems * sizeof(float), nullptr, nullptr, 0u});  // u, f, tmp
  }
  ctx->shared.push_back({ctx->allocs[0], ctx->lv[0].elems * sizeof(double), nullptr, nullptr, 0u});  // u64
  if (!ctx->flags) {
    CU(cudaMalloc((void**)&ctx->flags, 256));
    CU(cudaMemsetAsync(ctx->flags, 0, 256, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  ctx->shared.push_back({ctx->flags, 256, nullptr, nullptr, 0u});
  return 0;
}

int madgpu_ipc_export(madgpu_ctx* ctx, void* blob, size_t capacity, size_t* needed)
{
  if (!ctx) return MADGPU_EINVAL;
  if (ctx->world < 2) return fail(ctx, MADGPU_ESTATE, "not a z-slab context");
  CU(cudaSetDevice(ctx->p.device));
  int rc = build_shared_list(ctx);
  if (rc) return rc;
  const size_t bytes = ctx->shared.size() * sizeof(cudaIpcMemHandle_t);
  if (needed) *needed = bytes;
  if (!blob) return 0;
  if (capacity < bytes) return fail(ctx, MADGPU_EINVAL, "ipc blob needs %zu bytes", bytes);
  cudaIpcMemHandle_t* h = (cudaIpcMemHandle_t*)blob;
  for (size_t i = 0; i < ctx->shared.size(); ++i) CU(cudaIpcGetMemHandle(&h[i], ctx->shared[i].local));
  return 0;
}

int madgpu_ipc_import(madgpu_ctx* ctx, const void* blob_lower, const void* blob_upper)
{
  if (!ctx) return MADGPU_EINVAL;
  if (ctx->world < 2) return fail(ctx, MADGPU_ESTATE, "not a z-slab context");
  if ((ctx->rank > 0) != (blob_lower != nullptr) || (ctx->rank < ctx->world - 1) != (blob_upper != nullptr))
    return fail(ctx, MADGPU_EINVAL, "rank %d of %d: pass the blob of rank-1 (NULL on rank 0) and of rank+1 (NULL on the last rank)", ctx->rank, ctx->world);
  CU(cudaSetDevice(ctx->p.device));
  int rc = build_shared_list(ctx);
  if (rc) return rc;
  const char* e = stream_memops_load();
  if (e) return fail(ctx, MADGPU_ECUDA, "%s", e);
  for (int l = 0; l + 1 < ctx->nlevels; ++l)
    if (!use_fast(ctx, ctx->lv[l])) return fail(ctx, MADGPU_ESTATE, "level %d of this slab is relaxed by the generic kernels (nx < %d): peer-memory halo not available, the NCCL exchange stays in use", l, ctx->fast_min_nx);
  const cudaIpcMemHandle_t* lo = (const cudaIpcMemHandle_t*)blob_lower;
  const cudaIpcMemHandle_t* hi = (const cudaIpcMemHandle_t*)blob_upper;
  for (size_t i = 0; i < ctx->shared.size(); ++i) {
    if (lo && !ctx->shared[i].lo) CU(cudaIpcOpenMemHandle(&ctx->shared[i].lo, lo[i], cudaIpcMemLazyEnablePeerAccess));
    if (hi && !ctx->shared[i].hi) CU(cudaIpcOpenMemHandle(&ctx->shared[i].hi, hi[i], cudaIpcMemLazyEnablePeerAccess));
  }
  ctx->flags_lo = (uint32_t*)ctx->shared.back().lo;
  ctx->flags_hi = (uint32_t*)ctx->shared.back().hi;
  // Handshake before anything relies on the mechanism: a stream write into each neighbour's counters block (slots 2 / 3,
  // apart from the arrival counters) and a bounded host-side poll for theirs.  A driver that refuses stream memory
  // operations on a peer mapping, or a pair of GPUs without a working peer path, must make this call fail -- a missing
  // signal later would leave the neighbour's stream waiting for ever.
  const bool dbg = getenv("MADGPU_P2P_DEBUG") != nullptr;  // one line per step on stderr, to locate a rank that stalls or fails
  if (dbg) fprintf(stderr, "[madgpu p2p] rank %d/%d: %zu allocations mapped from %s%s, starting handshake\n", ctx->rank, ctx->world, ctx->shared.size(),
                   lo ? "rank-1 " : "", hi ? "rank+1" : "");
  const uint32_t magic = 0xA5A50000u + (uint32_t)ctx->world;
  int wr = 0;
  if (ctx->flags_lo) wr |= g_write_value(ctx->stream, (unsigned long long)(uintptr_t)(ctx->flags_lo + 3), magic, 0);
  if (ctx->flags_hi) wr |= g_write_value(ctx->stream, (unsigned long long)(uintptr_t)(ctx->flags_hi + 2), magic, 0);
  cudaError_t se = cudaStreamSynchronize(ctx->stream);
  bool ok = wr == 0 && se == cudaSuccess;
  const auto t0 = std::chrono::steady_clock::now();
  while (ok) {
    uint32_t got[2] = {0, 0};
    if (cudaMemcpy(got, ctx->flags + 2, sizeof got, cudaMemcpyDeviceToHost) != cudaSuccess) { ok = false; break; }
    if ((!ctx->flags_lo || got[0] == magic) && (!ctx->flags_hi || got[1] == magic)) break;
    if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 30.0) ok = false;  // neighbours open ~20 IPC handles first: allow for the skew
  }
  if (dbg) fprintf(stderr, "[madgpu p2p] rank %d: handshake %s after %.3f s (stream write rc %d, sync %s)\n", ctx->rank, ok ? "ok" : "FAILED",
                   std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(), wr, cudaGetErrorString(se));
  if (!ok) {
    cudaGetLastError();
    return fail(ctx, MADGPU_ECUDA, "peer-memory handshake with the neighbouring ranks failed (stream write rc %d, %s): the NCCL exchange stays in use", wr,
                cudaGetErrorString(se));
  }
  ctx->p2p = true;
  return 0;
}

int madgpu_ipc_disable(madgpu_ctx* ctx)
{
  if (!ctx) return MADGPU_EINVAL;
  ctx->p2p = false;
  for (auto& sh : ctx->shared) sh.produced = 0;
  return 0;
}

int madgpu_slab(const madgpu_ctx* ctx, int32_t level, int32_t* z_begin, int32_t* z_count, int32_t* global_nz)
{
  if (!ctx || level < 0 || level >= ctx->nlevels) return MADGPU_EINVAL;
  if (z_begin) *z_begin = ctx->lv[level].zb;
  if (z_count) *z_count = ctx->lv[level].n[2];
  if (global_nz) *global_nz = ctx->lv[level].gnz;
  return 0;
}

int madgpu_num_levels(const madgpu_ctx* ctx) { return ctx ? ctx->total_levels : MADGPU_EINVAL; }

int madgpu_level_info(const madgpu_ctx* ctx, int32_t level, int32_t size[3], double spacing[3], int32_t centering[3])
{
  if (!ctx || level < 0 || level >= ctx->total_levels) return MADGPU_EINVAL;
  for (int d = 0; d < 3; ++d) {  // global extents (a z-slab context reports its own planes through madgpu_slab)
    if (size) size[d] = ctx->gsize[level][d];
    if (spacing) spacing[d] = d < ctx->dim ? ctx->p.spacing[d] * (double)(1u << level) : 1.0;
    if (centering) centering[d] = ctx->gcent[level][d];
  }
  return 0;
}

// ---- per-operator entry points -----------------------------------------------------------------
#define CHECK_LEVEL(l)                                                                     \
  if (!ctx) return MADGPU_EINVAL;                                                          \
  if ((l) < 0 || (l) >= ctx->nlevels) return fail(ctx, MADGPU_EINVAL, "level %d out of range", (int)(l)); \
  CU(cudaSetDevice(ctx->p.device));

int madgpu_op_get_tensor(madgpu_ctx* ctx, int32_t level, float* planes)
{
  CHECK_LEVEL(level);
  if (!ctx->tensor_set) return fail(ctx, MADGPU_ESTATE, "tensor not set");
  Level& L = ctx->lv[level];
  const size_t nv = (size_t)L.n[0] * L.n[1] * L.n[2];
  for (int c = 0; c < ctx->ncomp; ++c) {
    int rc = download_field(ctx, L, L.D[c], planes + c * nv);
    if (rc) return rc;
  }
  return 0;
}

int madgpu_op_assemble(madgpu_ctx* ctx, int32_t level, float* stencil)
{
  CHECK_LEVEL(level);
  if (!ctx->tensor_set) return fail(ctx, MADGPU_ESTATE, "tensor not set");
  Level& L = ctx->lv[level];
  const size_t nv = (size_t)L.n[0] * L.n[1] * L.n[2];
  const int ns = ctx->dim == 2 ? 9 : 27;
  float* d = nullptr;
  CU(cudaMalloc((void**)&d, nv * ns * sizeof(float)));
  const dim3 b = block3(ctx->dim), g = grid3(L.g, b);
  if (ctx->dim == 3) MAD_LAUNCH((k_assemble<3>), g, b, 0, ctx->stream, L.g, tensor_of(L), d);
  else MAD_LAUNCH((k_assemble<2>), g, b, 0, ctx->stream, L.g, tensor_of(L), d);
  cudaError_t e = cudaMemcpyAsync(stencil, d, nv * ns * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d);
  if (e != cudaSuccess) return fail(ctx, MADGPU_ECUDA, "assemble: %s", cudaGetErrorString(e));
  return 0;
}

int madgpu_op_smooth(madgpu_ctx* ctx, int32_t level, int32_t smoother, int32_t n_iter, const float* u, const float* f, float* out)
{
  CHECK_LEVEL(level);
  if (!ctx->tensor_set) return fail(ctx, MADGPU_ESTATE, "tensor not set");
  Level& L = ctx->lv[level];
  int rc = upload_field(ctx, L, u, L.u);
  if (!rc) rc = upload_field(ctx, L, f, L.f);
  if (rc) return rc;
  op_smooth(ctx, level, smoother, n_iter);
  CU(cudaGetLastError());
  return download_field(ctx, L, L.u, out);
}

int madgpu_op_residual(madgpu_ctx* ctx, int32_t level, const float* u, const float* f, float* r, double* norm)
{
  CHECK_LEVEL(level);
  if (!ctx->tensor_set) return fail(ctx, MADGPU_ESTATE, "tensor not set");
  Level& L = ctx->lv[level];
  int rc = upload_field(ctx, L, u, L.u);
  if (!rc) rc = upload_field(ctx, L, f, L.f);
  if (rc) return rc;
  op_residual32(ctx, level, L.tmp, true);
  CU(cudaGetLastError());
  if (norm) *norm = std::sqrt(read_scalar(ctx));
  if (r) return download_field(ctx, L, L.tmp, r);
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int madgpu_op_residual_f64(madgpu_ctx* ctx, const double* u, const double* f, double* r, double* norm)
{
  CHECK_LEVEL(0);
  if (!ctx->tensor_set) return fail(ctx, MADGPU_ESTATE, "tensor not set");
  Level& L = ctx->lv[0];
  const size_t w = (size_t)L.g.nx * sizeof(double), dp = (size_t)L.g.pitch * sizeof(double), hrows = (size_t)L.g.ny * L.g.nz;
  halo_dirty(ctx, ctx->u64);
  CU(cudaMemcpy2DAsync(ctx->u64, dp, u, w, w, hrows, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpy2DAsync(ctx->f64, dp, f, w, w, hrows, cudaMemcpyHostToDevice, ctx->stream));
  if (!r) {  // norm only: the kernel of the solve loop (fp32 residual into lv[0].tmp + fp64 norm)
    op_residual64(ctx, L.tmp, nullptr);
    if (norm) *norm = std::sqrt(read_scalar(ctx));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    return 0;
  }
  double* d_r = nullptr;
  CU(cudaMalloc((void**)&d_r, L.elems * sizeof(double)));
  op_residual64(ctx, nullptr, d_r + L.g.plane);
  if (norm) *norm = std::sqrt(read_scalar(ctx));
  cudaError_t e = cudaSuccess;
  if (r) e = cudaMemcpy2DAsync(r, w, d_r + L.g.plane, dp, w, hrows, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_r);
  if (e != cudaSuccess) return fail(ctx, MADGPU_ECUDA, "residual_f64: %s", cudaGetErrorString(e));
  return 0;
}

int madgpu_op_restrict(madgpu_ctx* ctx, int32_t fine_level, const float* fine, float* coarse)
{
  CHECK_LEVEL(fine_level);
  if (fine_level + 1 >= ctx->nlevels) return fail(ctx, MADGPU_EINVAL, "level %d has no coarser level", fine_level);
  Level& F = ctx->lv[fine_level];
  Level& C = ctx->lv[fine_level + 1];
  int rc = upload_field(ctx, F, fine, F.tmp);
  if (rc) return rc;
  op_restrict<float>(ctx, fine_level, F.tmp, C.tmp);
  CU(cudaGetLastError());
  return download_field(ctx, C, C.tmp, coarse);
}

int madgpu_op_prolong(madgpu_ctx* ctx, int32_t fine_level, const float* coarse, float* fine)
{
  CHECK_LEVEL(fine_level);
  if (fine_level + 1 >= ctx->nlevels) return fail(ctx, MADGPU_EINVAL, "level %d has no coarser level", fine_level);
  Level& F = ctx->lv[fine_level];
  Level& C = ctx->lv[fine_level + 1];
  int rc = upload_field(ctx, C, coarse, C.tmp);
  if (rc) return rc;
  op_prolong<float, false>(ctx, fine_level, C.tmp, F.tmp);
  CU(cudaGetLastError());
  return download_field(ctx, F, F.tmp, fine);
}

int madgpu_op_coarse_solve(madgpu_ctx* ctx, const float* f, float* e)
{
  CHECK_LEVEL(0);
  if (!ctx->tensor_set) return fail(ctx, MADGPU_ESTATE, "tensor not set");
  Level& L = ctx->lv[ctx->nlevels - 1];
  int rc = upload_field(ctx, L, f, L.f);
  if (rc) return rc;
  op_coarse_solve(ctx);
  CU(cudaGetLastError());
  return download_field(ctx, L, L.u, e);
}

int madgpu_op_vcycle(madgpu_ctx* ctx, int32_t level, const float* u, const float* f, float* out)
{
  CHECK_LEVEL(level);
  if (!ctx->tensor_set) return fail(ctx, MADGPU_ESTATE, "tensor not set");
  Level& L = ctx->lv[level];
  int rc = upload_field(ctx, L, u, L.u);
  if (!rc) rc = upload_field(ctx, L, f, L.f);
  if (rc) return rc;
  vcycle(ctx, level);
  CU(cudaGetLastError());
  return download_field(ctx, L, L.u, out);
}

}  // extern "C"
