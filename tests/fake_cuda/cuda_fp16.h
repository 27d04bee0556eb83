// fake_cuda/cuda_fp16.h -- TEST INFRASTRUCTURE: the few half-precision intrinsics csrc/mad_fast.cuh uses, on the host compiler's
// _Float16 (IEEE binary16, conversions round to nearest even like the _rn intrinsics).
#ifndef FAKE_CUDA_FP16_H
#define FAKE_CUDA_FP16_H
struct __half2 {
  _Float16 x, y;  // x = low half
};
static_assert(sizeof(__half2) == 4, "__half2 must be one 32-bit word");
inline __half2 __floats2half2_rn(float a, float b) { return __half2{(_Float16)a, (_Float16)b}; }
inline float __low2float(__half2 h) { return (float)h.x; }
inline float __high2float(__half2 h) { return (float)h.y; }
#endif
