#!/bin/bash
# round 2, GPU call r (1 GPU): the final state with warp-private Gauss-Seidel pairs as the default -- smoke, full GPU suite, default bench, reference arm, launch list, ncu of the shipped kernels
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export PYTHONPATH="$PWD:$PYTHONPATH"
O=gpurun_out
(timeout 300 python __graft_entry__.py smoke 2>&1 | tail -12) > $O/r02r_smoke.log
(timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -20) > $O/r02r_pytest_gpu.log
timeout 500 python bench.py --steps 20 --warmup 5 > $O/r02r_bench.json 2> $O/r02r_bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02r_gs_launches.csv \
   python bench.py --steps 2 --warmup 1 --e2e-reps 0 --no-cpu-baseline --no-ved > $O/r02r_ncu_launches.log 2>&1
python tools/ncu_summary.py launches $O/r02r_gs_launches.csv > $O/r02r_gs_launches.txt 2>&1; rm -f $O/r02r_gs_launches.csv
cap() { # name regex skip
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" --launch-skip $3 -c 1 -o $O/r02r_full_$1 -f \
     python bench.py --steps 1 --warmup 1 --e2e-reps 0 --no-cpu-baseline --no-ved > $O/r02r_ncu_full_$1.log 2>&1
  python tools/ncu_summary.py full $O/r02r_full_$1.ncu-rep > $O/r02r_full_$1.txt 2>&1; rm -f $O/r02r_full_$1.ncu-rep
}
cap coef_gs2 'k_coef_gs2' 2
du -sh $O
echo done
