#!/bin/bash
# round 2, GPU call e (1 GPU): temporal blocking of the Gauss-Seidel sweeps (k_coef_gs_tb) -- ordering tests, solves, A/B bench, ncu
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export PYTHONPATH="$PWD:$PYTHONPATH"
O=gpurun_out
(timeout 900 python -m pytest tests/test_gpu_fast.py tests/test_gpu_solve.py tests/test_gpu_golden.py tests/test_gpu_large.py -m gpu -x -q 2>&1 | tail -30) > $O/r02e_pytest_gpu.log
B="--steps 20 --warmup 5 --e2e-reps 1 --no-cpu-baseline --no-ved"
for tb in 3 2 1; do MADGPU_GS_TB=$tb timeout 200 python bench.py $B > $O/r02e_bench_tb$tb.json 2> $O/r02e_bench_tb$tb.err; done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02e_gs_launches.csv \
   python bench.py --steps 2 --warmup 1 --e2e-reps 0 --no-cpu-baseline --no-ved > $O/r02e_ncu_launches.log 2>&1
python tools/ncu_summary.py launches $O/r02e_gs_launches.csv > $O/r02e_gs_launches.txt 2>&1; rm -f $O/r02e_gs_launches.csv
cap() { # name regex skip [keep]
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" --launch-skip $3 -c 1 -o $O/r02e_full_$1 -f \
     python bench.py --steps 1 --warmup 1 --e2e-reps 0 --no-cpu-baseline --no-ved > $O/r02e_ncu_full_$1.log 2>&1
  python tools/ncu_summary.py full $O/r02e_full_$1.ncu-rep > $O/r02e_full_$1.txt 2>&1
  [ "$4" = keep ] || rm -f $O/r02e_full_$1.ncu-rep
}
cap coef_gs_tb 'k_coef_gs_tb' 2 keep
du -sh $O
echo done
