// madgpu_host.cpp -- TEST INFRASTRUCTURE.  multigridanisotropicdiffusion_b200/csrc/madgpu.cu (with mad_kernels.cuh and
// mad_fast.cuh) and csrc/ved.cu (with ved_kernels.cuh) compiled UNMODIFIED for the host: kernels on fibres (fiber_shim.h), CUDA runtime calls on host memory
// (tests/fake_cuda/), NCCL through tests/mad_host/fake_nccl.cpp when MADGPU_NCCL_LIB points at it.  The result,
// tests/_build/libmadgpu_host.so, exports the same madgpu_* C-ABI as libmadgpu.so, so the CPU suite can drive the real solver
// source -- hierarchy, kernels, V-cycle / FMG drivers, z-slab decomposition with both halo mechanisms (every rank a thread) --
// against the oracle.  This is a check of the code, not a CPU path of the product: libmadgpu.so refuses to run without a GPU.
#include "fiber_shim.h"

#include "../../multigridanisotropicdiffusion_b200/csrc/madgpu.cu"
// the VED tensor front-end on the same shim, so that madved_run drives the REAL solver here (tests/test_cpu_ved_cabi.py runs
// ved.cu against an oracle-backed stand-in solver instead)
#include "../../multigridanisotropicdiffusion_b200/csrc/ved.cu"

extern "C" {
long long mad_host_launches() { return mad_host::blk().launches; }
long long mad_host_switches() { return mad_host::blk().switches; }
long long mad_host_live_allocs() { return fake_cuda::g_live_allocs; }
}
