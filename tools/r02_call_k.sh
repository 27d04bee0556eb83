#!/bin/bash
# round 2, GPU call k (1 GPU): A/B of the warp-private row-pair sweep (no CTA barriers) and of the staged shared-memory sweeps
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export PYTHONPATH="$PWD:$PYTHONPATH"
O=gpurun_out
B="--steps 20 --warmup 5 --e2e-reps 2 --no-cpu-baseline --no-ved"
(timeout 300 python -m pytest tests/test_gpu_fast.py -m gpu -x -q -k "fused_gs and (private or staged)" 2>&1 | tail -4) > $O/r02k_pytest_gpu.log
timeout 200 python bench.py $B > $O/r02k_bench_default.json 2> $O/r02k_bench_default.err
MADGPU_GS_PRIVATE=1 timeout 200 python bench.py $B > $O/r02k_bench_private.json 2> $O/r02k_bench_private.err
MADGPU_GS_TB_SINGLE=2 timeout 200 python bench.py $B > $O/r02k_bench_staged2.json 2> $O/r02k_bench_staged2.err
MADGPU_GS_TB_SINGLE=3 timeout 200 python bench.py $B > $O/r02k_bench_staged3.json 2> $O/r02k_bench_staged3.err
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_coef_gs2" --launch-skip 2 -c 1 -o $O/r02k_full_private -f \
   env MADGPU_GS_PRIVATE=1 python bench.py --steps 1 --warmup 1 --e2e-reps 0 --no-cpu-baseline --no-ved > $O/r02k_ncu_full_private.log 2>&1
python tools/ncu_summary.py full $O/r02k_full_private.ncu-rep > $O/r02k_full_private.txt 2>&1; rm -f $O/r02k_full_private.ncu-rep
echo done
