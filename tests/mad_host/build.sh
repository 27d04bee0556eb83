#!/bin/sh
# Builds the CPU test builds of the solver (TEST INFRASTRUCTURE, see the headers of the sources):
#   tests/_build/libmadgpu_host.so  multigridanisotropicdiffusion_b200/csrc/madgpu.cu compiled unmodified for the host
#                                   (kernels on fibres, CUDA runtime on host memory)
#   tests/_build/libfakenccl.so     the NCCL entry points libmadgpu binds, for ranks that are threads of one process
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
OUT=$ROOT/tests/_build
mkdir -p "$OUT"
# -fno-gnu-unique: static locals of inline / template functions (the kernels' __shared__ arrays) must not be unified with those
# of the other test builds loaded into the same pytest process
g++ -O2 -std=c++17 -fPIC -shared -w -fno-strict-aliasing -fno-gnu-unique -x c++ -I"$ROOT/tests/fake_cuda" -I"$HERE" -o "$OUT/libmadgpu_host.so" "$HERE/madgpu_host.cpp" -ldl
g++ -O2 -std=c++17 -fPIC -shared -pthread -Wall -o "$OUT/libfakenccl.so" "$HERE/fake_nccl.cpp"
echo "built $OUT/libmadgpu_host.so $OUT/libfakenccl.so"
