#!/bin/bash
# round 2, GPU call o (1 GPU): the row-pair sweep capped at 128 registers (16 instead of 12 warps per SM), shared and warp-private tiles
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export PYTHONPATH="$PWD:$PYTHONPATH"
O=gpurun_out
B="--steps 20 --warmup 5 --e2e-reps 2 --no-cpu-baseline --no-ved"
timeout 200 python bench.py $B > $O/r02o_bench_default.json 2> $O/r02o_bench_default.err
MADGPU_GS_PRIVATE=3 timeout 200 python bench.py $B > $O/r02o_bench_regs128.json 2> $O/r02o_bench_regs128.err
MADGPU_GS_PRIVATE=2 timeout 200 python bench.py $B > $O/r02o_bench_private_regs128.json 2> $O/r02o_bench_private_regs128.err
echo done
