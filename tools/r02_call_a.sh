#!/bin/bash
# round 2, first GPU call: everything written after round 1's budget ran out + profiles of the shipped kernels
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > $O/r02a_pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 --ved > $O/r02a_bench_ved.json 2> $O/r02a_bench_ved.err
MADGPU_RES64_COEF32=1 timeout 200 python bench.py --steps 20 --warmup 5 --e2e-reps 1 --no-cpu-baseline > $O/r02a_bench_res64c32.json 2> $O/r02a_bench_res64c32.err
timeout 200 python bench.py --steps 20 --warmup 5 --smoother wj --e2e-reps 1 --no-cpu-baseline > $O/r02a_bench_wj.json 2> $O/r02a_bench_wj.err
# launch list of the default cycle
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02a_gs_launches.csv \
   python bench.py --steps 2 --warmup 1 --e2e-reps 0 --no-cpu-baseline > $O/r02a_ncu_launches.log 2>&1
python tools/ncu_summary.py launches $O/r02a_gs_launches.csv > $O/r02a_gs_launches.txt 2>&1; rm -f $O/r02a_gs_launches.csv
# full capture: one level-0 launch of each kernel of the cycle
cap() { # name regex skip
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" --launch-skip $3 -c 1 -o $O/r02a_full_$1 -f \
     python bench.py --steps 1 --warmup 1 --e2e-reps 0 --no-cpu-baseline > $O/r02a_ncu_full_$1.log 2>&1
  python tools/ncu_summary.py full $O/r02a_full_$1.ncu-rep > $O/r02a_full_$1.txt 2>&1
  [ "$4" = keep ] || rm -f $O/r02a_full_$1.ncu-rep
}
cap coef_gs2 'k_coef_gs2' 1 keep
cap coef_residual 'k_coef_residual' 0
cap res64 'k_fast_sweep<\(int\)1, double|k_fast_sweep<1, double' 0 keep
cap restrict 'k_fast_restrict' 0
cap prolong 'k_fast_prolong' 3
cap axpy 'k_axpy_f64_f32' 0
# VED front-end launch list + full capture at 256^3
cat > /tmp/ved256.py <<'P'
import numpy as np, multigridanisotropicdiffusion_b200 as M
from multigridanisotropicdiffusion_b200 import phantom
img = phantom.vessel_phantom((256, 256, 256))[0].numpy()
f = M.VEDMultigridImageFilter("gs"); f.SetInput(img, phantom.VED_SPACING); f.SetOmega(1.5); f.SetDiffusionIterations(1); f.Update(); print(f.ved_stats, f.stats)
P
timeout 200 python /tmp/ved256.py > $O/r02a_ved256_plain.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r02a_ved_launches.csv python /tmp/ved256.py > $O/r02a_ved_ncu.log 2>&1
python tools/ncu_summary.py launches $O/r02a_ved_launches.csv > $O/r02a_ved_launches.txt 2>&1; rm -f $O/r02a_ved_launches.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_rg_|k_ved_update' -c 12 -o $O/r02a_full_ved -f python /tmp/ved256.py > $O/r02a_ncu_full_ved.log 2>&1
python tools/ncu_summary.py full $O/r02a_full_ved.ncu-rep > $O/r02a_full_ved.txt 2>&1; rm -f $O/r02a_full_ved.ncu-rep
nvidia-smi topo -m > $O/r02a_topo.txt 2>&1
du -sh $O
echo done
