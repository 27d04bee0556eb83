// fiber_shim.h -- TEST INFRASTRUCTURE.  Runs the CUDA kernels of multigridanisotropicdiffusion_b200/csrc (mad_kernels.cuh,
// mad_fast.cuh, launched from madgpu.cu through MAD_LAUNCH) on the host: every CUDA thread of a block is a fibre (ucontext) of the
// calling OS thread, blocks run one after the other, and the fibres are switched only where CUDA threads interact --
// __syncthreads(), __syncwarp(), the warp shuffles and the named barriers -- so lock-step semantics are reproduced exactly and
// cheaply.  `__shared__` is a static array of the (single) running block.  A barrier that not every live thread reaches is
// reported as the dead-lock it would be on the GPU.  Each OS thread (= one rank of a multi-rank test) has its own scheduler.
// Nothing in the product includes this header; the CUDA path never runs on the CPU.
#ifndef MAD_HOST_FIBER_SHIM_H
#define MAD_HOST_FIBER_SHIM_H

#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#include "../fake_cuda/cuda_runtime.h"

// Context switch between fibres.  glibc's swapcontext saves and restores the signal mask with two system calls per switch, which
// dominates the run time of the warp shuffles; on x86-64 a dozen instructions do (callee-saved registers + stack pointer).
#if defined(__x86_64__) && !defined(MAD_HOST_USE_UCONTEXT)
#define MAD_HOST_ASM_SWITCH 1
extern "C" void mad_host_switch(void** save_sp, void* load_sp);
asm(R"(
    .text
    .p2align 4
    .globl mad_host_switch
    .hidden mad_host_switch
    .type mad_host_switch, @function
mad_host_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
    .size mad_host_switch, .-mad_host_switch
)");
#endif

namespace mad_host
{
enum State { RUN, WAIT_BLOCK, WAIT_WARP, WAIT_NAMED, DONE };
constexpr size_t STACK_BYTES = 192 * 1024;
constexpr int MAX_NAMED = 16;

struct Fiber {
#ifdef MAD_HOST_ASM_SWITCH
  void* sp;
#else
  ucontext_t ctx;
#endif
  State state;
  uint3 tid;
  unsigned linear, warp, lane;
  int named_id;
};

struct Block {
  std::vector<Fiber> fibers;
  char* stacks = nullptr;  // uninitialised on purpose (192 KB per CUDA thread of the largest block seen)
  size_t stacks_bytes = 0;
#ifdef MAD_HOST_ASM_SWITCH
  void* sched_sp;
#else
  ucontext_t sched;
#endif
  unsigned current;
  dim3 bdim, gdim;
  uint3 bidx;
  unsigned live, block_waiting;
  std::vector<unsigned> warp_live, warp_waiting;
  std::vector<uint64_t> xchg;  // 32 slots per warp
  std::vector<char> dyn_smem;
  unsigned named_count[MAX_NAMED], named_need[MAX_NAMED];
  const std::function<void()>* body;
  long long launches, switches;
};

inline Block& blk()
{
  static thread_local Block b;
  return b;
}
inline Fiber& cur() { Block& b = blk(); return b.fibers[b.current]; }

inline void yield_to_scheduler()
{
  Block& b = blk();
  ++b.switches;
#ifdef MAD_HOST_ASM_SWITCH
  mad_host_switch(&b.fibers[b.current].sp, b.sched_sp);
#else
  swapcontext(&b.fibers[b.current].ctx, &b.sched);
#endif
}

inline void fiber_entry()
{
  Block& b = blk();
  (*b.body)();
  Fiber& f = b.fibers[b.current];
  f.state = DONE;
  --b.live;
  --b.warp_live[f.warp];
#ifdef MAD_HOST_ASM_SWITCH
  mad_host_switch(&f.sp, b.sched_sp);
#else
  swapcontext(&f.ctx, &b.sched);
#endif
  std::abort();  // a finished fibre is never resumed
}

inline void run_block(Block& b, unsigned nthreads)
{
  if (b.fibers.size() < nthreads) b.fibers.resize(nthreads);
  if (b.stacks_bytes < (size_t)nthreads * STACK_BYTES) {
    std::free(b.stacks);
    b.stacks_bytes = (size_t)nthreads * STACK_BYTES;
    b.stacks = static_cast<char*>(std::malloc(b.stacks_bytes));
  }
  const unsigned nwarps = (nthreads + 31) / 32;
  b.warp_live.assign(nwarps, 0);
  b.warp_waiting.assign(nwarps, 0);
  b.xchg.assign((size_t)nwarps * 32, 0);
  for (int i = 0; i < MAX_NAMED; ++i) b.named_count[i] = b.named_need[i] = 0;
  b.live = nthreads;
  b.block_waiting = 0;
  for (unsigned t = 0; t < nthreads; ++t) {
    Fiber& f = b.fibers[t];
    f.linear = t;
    f.warp = t / 32;
    f.lane = t % 32;
    f.tid.x = t % b.bdim.x;
    f.tid.y = (t / b.bdim.x) % b.bdim.y;
    f.tid.z = t / (b.bdim.x * b.bdim.y);
    f.state = RUN;
    f.named_id = -1;
    ++b.warp_live[f.warp];
#ifdef MAD_HOST_ASM_SWITCH
    {  // initial frame: six callee-saved registers, then fiber_entry as the return address; entered with rsp = 8 mod 16
      uintptr_t top = (reinterpret_cast<uintptr_t>(b.stacks + (size_t)(t + 1) * STACK_BYTES)) & ~(uintptr_t)15;
      void** q = reinterpret_cast<void**>(top);
      *--q = nullptr;                                      // where fiber_entry would return to (it never does)
      *--q = reinterpret_cast<void*>(&fiber_entry);        // popped by `ret`
      for (int i = 0; i < 6; ++i) *--q = nullptr;          // rbp rbx r12 r13 r14 r15
      f.sp = q;
    }
#else
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = b.stacks + (size_t)t * STACK_BYTES;
    f.ctx.uc_stack.ss_size = STACK_BYTES;
    f.ctx.uc_link = nullptr;
    makecontext(&f.ctx, fiber_entry, 0);
#endif
  }
  while (b.live > 0) {
    bool progressed = false;
    for (unsigned t = 0; t < nthreads; ++t) {
      if (b.fibers[t].state != RUN) continue;
      b.current = t;
#ifdef MAD_HOST_ASM_SWITCH
      mad_host_switch(&b.sched_sp, b.fibers[t].sp);
#else
      swapcontext(&b.sched, &b.fibers[t].ctx);
#endif
      progressed = true;
    }
    if (b.live > 0 && b.block_waiting == b.live) {  // __syncthreads: every thread that has not exited
      for (unsigned t = 0; t < nthreads; ++t)
        if (b.fibers[t].state == WAIT_BLOCK) b.fibers[t].state = RUN;
      b.block_waiting = 0;
      progressed = true;
    }
    for (unsigned w = 0; w < nwarps; ++w)
      if (b.warp_live[w] > 0 && b.warp_waiting[w] == b.warp_live[w]) {
        for (unsigned t = w * 32; t < std::min(nthreads, (w + 1) * 32); ++t)
          if (b.fibers[t].state == WAIT_WARP) b.fibers[t].state = RUN;
        b.warp_waiting[w] = 0;
        progressed = true;
      }
    for (int id = 0; id < MAX_NAMED; ++id)
      if (b.named_need[id] && b.named_count[id] >= b.named_need[id]) {  // bar.sync / bar.arrive with a thread count
        for (unsigned t = 0; t < nthreads; ++t)
          if (b.fibers[t].state == WAIT_NAMED && b.fibers[t].named_id == id) b.fibers[t].state = RUN;
        b.named_count[id] -= b.named_need[id];
        if (b.named_count[id] == 0) b.named_need[id] = 0;
        progressed = true;
      }
    if (!progressed) {
      std::fprintf(stderr, "mad_host: dead-lock in block (%u,%u,%u): %u live threads, %u at __syncthreads\n", b.bidx.x, b.bidx.y, b.bidx.z, b.live,
                   b.block_waiting);
      std::abort();
    }
  }
}

inline bool trace()
{
  static const bool on = std::getenv("MAD_HOST_TRACE") != nullptr;
  return on;
}

template <typename F>
void launch(dim3 grid, dim3 block, size_t smem, F body)
{
  if (trace()) std::fprintf(stderr, "mad_host:   grid (%u,%u,%u) block (%u,%u,%u) smem %zu\n", grid.x, grid.y, grid.z, block.x, block.y, block.z, smem);
  Block& b = blk();
  const std::function<void()> fn = body;
  b.body = &fn;
  b.bdim = block;
  b.gdim = grid;
  ++b.launches;
  if (b.dyn_smem.size() < smem) b.dyn_smem.resize(smem);
  const unsigned nthreads = block.x * block.y * block.z;
  for (unsigned z = 0; z < grid.z; ++z)
    for (unsigned y = 0; y < grid.y; ++y)
      for (unsigned x = 0; x < grid.x; ++x) {
        b.bidx = uint3{x, y, z};
        run_block(b, nthreads);
      }
}

inline void sync_block()
{
  Block& b = blk();
  cur().state = WAIT_BLOCK;
  ++b.block_waiting;
  yield_to_scheduler();
}
inline void sync_warp()
{
  Block& b = blk();
  Fiber& f = cur();
  f.state = WAIT_WARP;
  ++b.warp_waiting[f.warp];
  yield_to_scheduler();
}
inline void named_sync(int id, int nthreads)
{
  Block& b = blk();
  Fiber& f = cur();
  b.named_need[id] = (unsigned)nthreads;
  ++b.named_count[id];
  f.state = WAIT_NAMED;
  f.named_id = id;
  yield_to_scheduler();
}
inline void named_arrive(int id, int nthreads)
{
  Block& b = blk();
  b.named_need[id] = (unsigned)nthreads;
  ++b.named_count[id];
}

// src < 0 or > 31: the lane keeps its own value (what the hardware does for out-of-range sources)
template <typename T>
T shuffle(T v, int src)
{
  static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
  Block& b = blk();
  Fiber& f = cur();
  uint64_t raw = 0;
  std::memcpy(&raw, &v, sizeof v);
  b.xchg[(size_t)f.warp * 32 + f.lane] = raw;
  sync_warp();
  if (src >= 0 && src < 32) raw = b.xchg[(size_t)f.warp * 32 + src];
  T r;
  std::memcpy(&r, &raw, sizeof r);
  sync_warp();  // nobody overwrites a slot before everybody has read
  return r;
}
}  // namespace mad_host

// ---- the CUDA surface the kernels use ---------------------------------------------------------------------------------------
#define MAD_HOST_EMULATION 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static thread_local  // one running block per OS thread (= rank)
#define threadIdx (mad_host::cur().tid)
#define blockIdx (mad_host::blk().bidx)
#define blockDim (mad_host::blk().bdim)
#define gridDim (mad_host::blk().gdim)
#define MAD_DYNAMIC_SHARED(type, name) type* name = reinterpret_cast<type*>(mad_host::blk().dyn_smem.data())
#define MAD_UNPAREN(...) __VA_ARGS__
#define MAD_LAUNCH(kernel, grid, block, smem, stream, ...)                                                                   \
  do {                                                                                                                      \
    if (mad_host::trace()) std::fprintf(stderr, "mad_host: launch %s\n", #kernel);                                          \
    if ((stream) && (stream)->rec) /* stream capture: record the launch with its arguments by value, run it at cudaGraphLaunch */ \
      (stream)->rec->push_back([=] { mad_host::launch(dim3(grid), dim3(block), (size_t)(smem), [&] { MAD_UNPAREN kernel(__VA_ARGS__); }); }); \
    else                                                                                                                    \
      mad_host::launch(dim3(grid), dim3(block), (size_t)(smem), [&] { MAD_UNPAREN kernel(__VA_ARGS__); });                   \
  } while (0)

// csrc/ved_kernels.cuh launches through VED_LAUNCH (no dynamic shared memory)
#define VED_LAUNCH(kernel, grid, block, stream, ...)                                                                        \
  do {                                                                                                                      \
    if (mad_host::trace()) std::fprintf(stderr, "mad_host: launch %s\n", #kernel);                                          \
    mad_host::launch(dim3(grid), dim3(block), 0, [&] { kernel(__VA_ARGS__); });                                             \
  } while (0)

inline void __syncthreads() { mad_host::sync_block(); }
inline void __syncwarp(unsigned = 0xffffffffu) { mad_host::sync_warp(); }
inline void __threadfence_block() {}
template <typename T> T __shfl_sync(unsigned, T v, int src) { return mad_host::shuffle(v, src & 31); }
template <typename T> T __shfl_up_sync(unsigned, T v, unsigned d) { return mad_host::shuffle(v, (int)mad_host::cur().lane - (int)d); }
template <typename T> T __shfl_down_sync(unsigned, T v, unsigned d) { return mad_host::shuffle(v, (int)mad_host::cur().lane + (int)d); }
template <typename T> T __shfl_xor_sync(unsigned, T v, int m) { return mad_host::shuffle(v, (int)mad_host::cur().lane ^ m); }
template <typename T> T __ldg(const T* p) { return *p; }
inline float __fdividef(float a, float b) { return a / b; }
inline long long clock64() { return 2 * std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }  // ~2 GHz
using std::max;
using std::min;
namespace mad { namespace fast {
inline void bar_sync(int id, int nthreads) { mad_host::named_sync(id, nthreads); }
inline void bar_arrive(int id, int nthreads) { mad_host::named_arrive(id, nthreads); }
} }

#endif  // MAD_HOST_FIBER_SHIM_H
