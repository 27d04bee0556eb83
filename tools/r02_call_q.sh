#!/bin/bash
# round 2, GPU call q (1 GPU): warp-private row pairs on a grid that alternates between even and odd first rows (MADGPU_GS_PRIVATE=2)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export PYTHONPATH="$PWD:$PYTHONPATH"
O=gpurun_out
B="--steps 20 --warmup 5 --e2e-reps 2 --no-cpu-baseline --no-ved"
(timeout 200 python -m pytest tests/test_gpu_fast.py -m gpu -x -q -k "fused_gs and priv" 2>&1 | tail -3) > $O/r02q_pytest_gpu.log
timeout 200 python bench.py $B > $O/r02q_bench_default.json 2> $O/r02q_bench_default.err
MADGPU_GS_PRIVATE=2 timeout 200 python bench.py $B > $O/r02q_bench_privalt.json 2> $O/r02q_bench_privalt.err
MADGPU_GS_PRIVATE=1 timeout 200 python bench.py $B > $O/r02q_bench_private.json 2> $O/r02q_bench_private.err
echo done
