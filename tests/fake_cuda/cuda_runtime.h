// fake_cuda/cuda_runtime.h -- TEST INFRASTRUCTURE.  The subset of the CUDA runtime API that multigridanisotropicdiffusion_b200/
// csrc/ved.cu calls, implemented on host memory (cudaMalloc = malloc, copies = memcpy, one synchronous "stream", events = wall
// clock), so that tests/ved_cabi_host.cpp can compile ved.cu UNMODIFIED for the host and the CPU suite can drive its C-ABI
// (context, staging, chunking, call-sequence checks) without a GPU.  "Device memory" is poisoned on allocation so that reads of
// never-written memory show up.  The product is never built against this header.
#ifndef FAKE_CUDA_RUNTIME_H
#define FAKE_CUDA_RUNTIME_H

#include <chrono>
#include <cstdlib>
#include <cstring>

enum cudaError_t { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
struct FakeStream { int unused; };
typedef FakeStream* cudaStream_t;  // in-order and synchronous
struct FakeEvent { std::chrono::steady_clock::time_point t; };
typedef FakeEvent* cudaEvent_t;
enum { cudaStreamNonBlocking = 1 };
struct cudaDeviceProp { int major, minor; };

namespace fake_cuda { inline int g_devices = 1; inline long long g_live_allocs = 0; }

inline cudaError_t cudaGetDeviceCount(int* n) { *n = fake_cuda::g_devices; return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { p->major = 10; p->minor = 0; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "fake CUDA error"; }
inline cudaError_t cudaMalloc(void** p, size_t bytes)
{
  *p = std::malloc(bytes ? bytes : 1);
  if (!*p) return cudaErrorMemoryAllocation;
  std::memset(*p, 0xFF, bytes);  // NaN pattern for floats and doubles
  ++fake_cuda::g_live_allocs;
  return cudaSuccess;
}
inline cudaError_t cudaFree(void* p) { if (p) { std::free(p); --fake_cuda::g_live_allocs; } return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind, cudaStream_t) { std::memmove(dst, src, bytes); return cudaSuccess; }
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

#endif  // FAKE_CUDA_RUNTIME_H
