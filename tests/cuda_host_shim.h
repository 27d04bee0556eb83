// cuda_host_shim.h -- TEST INFRASTRUCTURE.  Just enough of the CUDA execution model for tests/ved_host_harness.cpp to run the
// kernels of multigridanisotropicdiffusion_b200/csrc/ved_kernels.cuh -- the unmodified kernel source, with its launch geometry --
// on the host: one std::thread per CUDA thread of a block (1-D blocks), blocks one after the other, `__shared__` as a static
// array of the (single) running block, `__syncwarp()` as a barrier over the 32 threads of a warp.  Nothing in the product
// includes this header; the CUDA path never runs on the CPU.
#ifndef CUDA_HOST_SHIM_H
#define CUDA_HOST_SHIM_H

#include <algorithm>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

struct shim_dim3 {
  unsigned x, y, z;
  shim_dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

namespace cuda_host
{
class WarpBarrier  // reusable barrier for the lanes of one warp that are still running
{
public:
  explicit WarpBarrier(int n) : m_n(n), m_waiting(0), m_gen(0) {}
  void arrive_and_wait()
  {
    std::unique_lock<std::mutex> lk(m_m);
    const unsigned gen = m_gen;
    if (++m_waiting == m_n) {
      m_waiting = 0;
      ++m_gen;
      m_cv.notify_all();
    } else {
      m_cv.wait(lk, [&] { return gen != m_gen; });
    }
  }
private:
  std::mutex m_m;
  std::condition_variable m_cv;
  int m_n, m_waiting;
  unsigned m_gen;
};

inline thread_local shim_dim3 t_threadIdx, t_blockIdx;
inline shim_dim3 g_blockDim, g_gridDim;
inline thread_local WarpBarrier* t_warp = nullptr;
inline long long g_launches = 0;

template <typename F>
void launch(shim_dim3 grid, shim_dim3 block, F body)
{
  g_blockDim = block;
  g_gridDim = grid;
  ++g_launches;
  const unsigned nthreads = block.x, nwarps = (nthreads + 31) / 32;
  for (unsigned b = 0; b < grid.x; ++b) {
    std::vector<std::unique_ptr<WarpBarrier> > bars;
    for (unsigned w = 0; w < nwarps; ++w) bars.emplace_back(new WarpBarrier((int)std::min(32u, nthreads - 32 * w)));
    std::vector<std::thread> th;
    th.reserve(nthreads);
    for (unsigned t = 0; t < nthreads; ++t)
      th.emplace_back([&, t, b] {
        t_threadIdx = shim_dim3(t, 0, 0);
        t_blockIdx = shim_dim3(b, 0, 0);
        t_warp = bars[t / 32].get();
        body();
      });
    for (auto& x : th) x.join();
  }
}
}  // namespace cuda_host

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define threadIdx (cuda_host::t_threadIdx)
#define blockIdx (cuda_host::t_blockIdx)
#define blockDim (cuda_host::g_blockDim)
#define gridDim (cuda_host::g_gridDim)
// all lanes of a warp either return before the first barrier or take part in every one (true for k_rg_rows: the early exit is
// warp-uniform); a lane that returned early while others wait would dead-lock here, which is the bug it would be on the GPU
inline void __syncwarp() { cuda_host::t_warp->arrive_and_wait(); }
using std::min;
using std::max;

#define VED_LAUNCH(kernel, grid, block, stream, ...) cuda_host::launch(shim_dim3(grid), shim_dim3(block), [&] { kernel(__VA_ARGS__); })

#endif  // CUDA_HOST_SHIM_H
