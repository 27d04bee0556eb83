// fake_cuda/cuda_runtime.h -- TEST INFRASTRUCTURE.  The subset of the CUDA runtime API that multigridanisotropicdiffusion_b200/
// csrc/ved.cu calls, implemented on host memory (cudaMalloc = malloc, copies = memcpy, one synchronous "stream", events = wall
// clock), so that tests/ved_cabi_host.cpp can compile ved.cu UNMODIFIED for the host and the CPU suite can drive its C-ABI
// (context, staging, chunking, call-sequence checks) without a GPU.  "Device memory" is poisoned on allocation so that reads of
// never-written memory show up.  The product is never built against this header.
#ifndef FAKE_CUDA_RUNTIME_H
#define FAKE_CUDA_RUNTIME_H

#include <sched.h>
#include <sys/mman.h>

#include <chrono>
#include <functional>
#include <vector>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

// ---- vector types and dim3 (what the kernels of csrc/ use of vector_types.h) ----
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) double2 { double x, y; };
struct uint3 { unsigned x, y, z; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline double2 make_double2(double x, double y) { return double2{x, y}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }

enum cudaError_t { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum { cudaEventDisableTiming = 2, cudaEnableDefault = 0, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0 };
struct cudaIpcMemHandle_t { char reserved[64]; };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
// in-order and synchronous; while a capture is open (cudaStreamBeginCapture) launches, memsets and copies are recorded as closures
// with their arguments by value -- the semantics of a CUDA graph -- instead of being run
struct FakeStream { std::vector<std::function<void()>>* rec = nullptr; };
typedef FakeStream* cudaStream_t;
struct FakeGraph { std::vector<std::function<void()>> ops; };
typedef FakeGraph* cudaGraph_t;
typedef FakeGraph* cudaGraphExec_t;
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal = 0, cudaStreamCaptureModeThreadLocal = 1, cudaStreamCaptureModeRelaxed = 2 };
struct FakeEvent { std::chrono::steady_clock::time_point t; };
typedef FakeEvent* cudaEvent_t;
enum { cudaStreamNonBlocking = 1 };
struct cudaDeviceProp { int major, minor; char name[64]; };

namespace fake_cuda { inline int g_devices = 1; inline long long g_live_allocs = 0; }

inline cudaError_t cudaGetDeviceCount(int* n) { *n = fake_cuda::g_devices; return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { p->major = 10; p->minor = 0; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "fake CUDA error"; }
// "Device" allocations are poisoned (0xFF: NaN as float and double) and fenced by guard zones that are checked when the block is
// freed: a kernel that writes outside its buffers aborts the test, one that reads outside or reads memory nobody wrote gets NaNs.
namespace fake_cuda
{
constexpr size_t GUARD = 4096;
constexpr unsigned char GUARD_BYTE = 0xA5;
struct Header { size_t bytes; size_t magic; };
}  // namespace fake_cuda
namespace fake_cuda
{
// FAKE_CUDA_GUARD_PAGES=1 (2): every allocation ends (starts) at an inaccessible page, so that an out-of-bounds READ faults too
// (electric-fence style; slow and memory-hungry, for one-off checks of the kernels' addressing)
struct PageAlloc { char* base; size_t total; };
inline std::mutex g_pages_m;
inline std::map<void*, PageAlloc> g_pages;
inline int guard_pages_mode()
{
  static const int m = std::getenv("FAKE_CUDA_GUARD_PAGES") ? std::atoi(std::getenv("FAKE_CUDA_GUARD_PAGES")) : 0;
  return m;
}
}  // namespace fake_cuda
inline cudaError_t cudaMalloc(void** p, size_t bytes)
{
  using namespace fake_cuda;
  if (guard_pages_mode()) {
    const size_t pg = 4096, body = (bytes + 15) & ~(size_t)15, inner = (body + pg - 1) / pg * pg, total = inner + 2 * pg;
    char* base = static_cast<char*>(mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0));
    if (base == MAP_FAILED) { *p = nullptr; return cudaErrorMemoryAllocation; }
    mprotect(base, pg, PROT_NONE);
    mprotect(base + total - pg, pg, PROT_NONE);
    char* q = guard_pages_mode() == 2 ? base + pg : base + total - pg - body;
    std::memset(q, 0xFF, body);
    { std::lock_guard<std::mutex> g(g_pages_m); g_pages[q] = PageAlloc{base, total}; }
    *p = q;
    __atomic_add_fetch(&g_live_allocs, 1, __ATOMIC_RELAXED);
    return cudaSuccess;
  }
  char* raw = static_cast<char*>(std::malloc(bytes + 2 * GUARD));
  if (!raw) { *p = nullptr; return cudaErrorMemoryAllocation; }
  std::memset(raw, GUARD_BYTE, GUARD);
  std::memset(raw + GUARD, 0xFF, bytes);
  std::memset(raw + GUARD + bytes, GUARD_BYTE, GUARD);
  Header h = {bytes, 0xC0DAC0DAu};
  std::memcpy(raw, &h, sizeof h);  // the first bytes of the lower guard hold the size
  *p = raw + GUARD;
  __atomic_add_fetch(&g_live_allocs, 1, __ATOMIC_RELAXED);
  return cudaSuccess;
}
inline cudaError_t cudaFree(void* p)
{
  using namespace fake_cuda;
  if (!p) return cudaSuccess;
  if (guard_pages_mode()) {
    PageAlloc a;
    {
      std::lock_guard<std::mutex> g(g_pages_m);
      auto it = g_pages.find(p);
      if (it == g_pages.end()) { std::fprintf(stderr, "fake_cuda: cudaFree of an unknown pointer %p\n", p); std::abort(); }
      a = it->second;
      g_pages.erase(it);
    }
    munmap(a.base, a.total);
    __atomic_sub_fetch(&g_live_allocs, 1, __ATOMIC_RELAXED);
    return cudaSuccess;
  }
  char* raw = static_cast<char*>(p) - GUARD;
  Header h;
  std::memcpy(&h, raw, sizeof h);
  bool ok = h.magic == 0xC0DAC0DAu;
  for (size_t i = sizeof h; ok && i < GUARD; ++i) ok = (unsigned char)raw[i] == GUARD_BYTE;
  for (size_t i = 0; ok && i < GUARD; ++i) ok = (unsigned char)raw[GUARD + h.bytes + i] == GUARD_BYTE;
  if (!ok) {
    std::fprintf(stderr, "fake_cuda: a guard zone of the %zu-byte allocation %p was overwritten (out-of-bounds write)\n", h.magic == 0xC0DAC0DAu ? h.bytes : 0, p);
    std::abort();
  }
  std::free(raw);
  __atomic_sub_fetch(&g_live_allocs, 1, __ATOMIC_RELAXED);
  return cudaSuccess;
}
inline cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind, cudaStream_t s)
{
  if (s && s->rec) s->rec->push_back([=] { std::memmove(dst, src, bytes); });
  else std::memmove(dst, src, bytes);
  return cudaSuccess;
}
inline cudaError_t cudaStreamBeginCapture(cudaStream_t s, cudaStreamCaptureMode)
{
  if (!s || s->rec) return cudaErrorInvalidValue;
  s->rec = new std::vector<std::function<void()>>();
  return cudaSuccess;
}
inline cudaError_t cudaStreamEndCapture(cudaStream_t s, cudaGraph_t* g)
{
  if (!s || !s->rec) return cudaErrorInvalidValue;
  *g = new FakeGraph{std::move(*s->rec)};
  delete s->rec;
  s->rec = nullptr;
  return cudaSuccess;
}
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t g, unsigned long long) { *e = new FakeGraph(*g); return cudaSuccess; }
inline cudaError_t cudaGraphDestroy(cudaGraph_t g) { delete g; return cudaSuccess; }
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t)
{
  for (auto& op : e->ops) op();
  return cudaSuccess;
}
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = new FakeStream(); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new FakeEvent(); return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b)
{
  *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
  return cudaSuccess;
}

inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind) { std::memmove(dst, src, bytes); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t bytes, cudaStream_t s)
{
  if (s && s->rec) s->rec->push_back([=] { std::memset(p, v, bytes); });
  else std::memset(p, v, bytes);
  return cudaSuccess;
}
inline cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, cudaMemcpyKind, cudaStream_t s)
{
  auto op = [=] { for (size_t r = 0; r < height; ++r) std::memmove((char*)dst + r * dpitch, (const char*)src + r * spitch, width); };
  if (s && s->rec) s->rec->push_back(op);
  else op();
  return cudaSuccess;
}
inline cudaError_t cudaMallocHost(void** p, size_t bytes) { *p = std::malloc(bytes ? bytes : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
template <typename F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
// "device memory" is this process's heap and every rank of a multi-rank test is a thread of it: an IPC handle is the pointer
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { std::memset(h, 0, sizeof *h); std::memcpy(h->reserved, &p, sizeof p); return cudaSuccess; }
inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) { std::memcpy(p, h.reserved, sizeof *p); return *p ? cudaSuccess : cudaErrorInvalidValue; }
inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }

// stream memory operations (the driver entry points the peer-memory halo binds): the stream is synchronous, so a write is a
// release store and a wait blocks the calling rank (thread) until the value arrives.  A wait that lasts FAKE_CUDA_WAIT_TIMEOUT_S
// seconds (default 60) reports what it was waiting for and aborts -- on a GPU that would be a hung stream.
namespace fake_cuda
{
inline int stream_write_value32(cudaStream_t, unsigned long long addr, uint32_t value, unsigned)
{
  __atomic_store_n(reinterpret_cast<uint32_t*>(addr), value, __ATOMIC_RELEASE);
  return 0;
}
inline int stream_wait_value32(cudaStream_t, unsigned long long addr, uint32_t value, unsigned flags)
{
  const auto t0 = std::chrono::steady_clock::now();
  const char* e = std::getenv("FAKE_CUDA_WAIT_TIMEOUT_S");
  const double limit = e ? std::atof(e) : 60.0;
  for (unsigned long spin = 0;; ++spin) {
    const uint32_t cur = __atomic_load_n(reinterpret_cast<uint32_t*>(addr), __ATOMIC_ACQUIRE);
    if (flags == 1 ? (int32_t)(cur - value) >= 0 : cur == value) return 0;  // 1 = CU_STREAM_WAIT_VALUE_GEQ, 0 = EQ
    if ((spin & 1023) == 1023) {
      sched_yield();
      if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > limit) {
        std::fprintf(stderr, "fake_cuda: stream wait on %p for value %u timed out (current %u) -- a hung stream on a GPU\n", (void*)addr, value, cur);
        std::abort();
      }
    }
  }
}
}  // namespace fake_cuda
inline cudaError_t cudaGetDriverEntryPoint(const char* name, void** f, unsigned long long, cudaDriverEntryPointQueryResult* q = nullptr)
{
  *f = nullptr;
  if (!std::strcmp(name, "cuStreamWriteValue32")) *f = (void*)&fake_cuda::stream_write_value32;
  if (!std::strcmp(name, "cuStreamWaitValue32")) *f = (void*)&fake_cuda::stream_wait_value32;
  if (q) *q = cudaDriverEntryPointSuccess;
  return *f ? cudaSuccess : cudaErrorInvalidValue;
}

#endif  // FAKE_CUDA_RUNTIME_H
