#!/bin/bash
# round 2, GPU call h (1 GPU): A/B of the shared-memory fp64 residual (k_fast_res64), the prefetching marching restriction, the
# single-sweep shared-memory Gauss-Seidel (k_coef_gs_tb<1>), the batched recursive-Gaussian line filter; ncu of the new kernels
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export PYTHONPATH="$PWD:$PYTHONPATH"
O=gpurun_out
(timeout 600 python -m pytest tests/test_gpu_fast.py tests/test_gpu_ved.py tests/test_gpu_solve.py -m gpu -x -q 2>&1 | tail -8) > $O/r02h_pytest_gpu.log
B="--steps 20 --warmup 5 --e2e-reps 2 --no-cpu-baseline --no-ved"
timeout 200 python bench.py $B > $O/r02h_bench_default.json 2> $O/r02h_bench_default.err
MADGPU_RES64_SMEM=0 timeout 200 python bench.py $B > $O/r02h_bench_res64reg.json 2> $O/r02h_bench_res64reg.err
MADGPU_RES64_SMEM=4 timeout 200 python bench.py $B > $O/r02h_bench_res64smem4.json 2> $O/r02h_bench_res64smem4.err
MADGPU_GS_TB_SINGLE=1 timeout 200 python bench.py $B > $O/r02h_bench_tbsingle.json 2> $O/r02h_bench_tbsingle.err
MADGPU_RESTRICT_CELL=0 timeout 200 python bench.py $B > $O/r02h_bench_oldrestrict.json 2> $O/r02h_bench_oldrestrict.err
timeout 300 python tools/ved_probe.py 512 > $O/r02h_ved_probe.log 2>&1
cap() { # name regex skip
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" --launch-skip $3 -c 1 -o $O/r02h_full_$1 -f \
     python bench.py --steps 1 --warmup 1 --e2e-reps 0 --no-cpu-baseline --no-ved > $O/r02h_ncu_full_$1.log 2>&1
  python tools/ncu_summary.py full $O/r02h_full_$1.ncu-rep > $O/r02h_full_$1.txt 2>&1; rm -f $O/r02h_full_$1.ncu-rep
}
cap res64 'k_fast_res64' 1
MADGPU_GS_TB_SINGLE=1 cap tbsingle 'k_coef_gs_tb' 0
du -sh $O
echo done
