#!/bin/bash
# round 2, GPU call b (1 GPU): new tests, bench with graphs / blocked prolongation / fp32-row stop-test residual, A/B runs, VED profile
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export PYTHONPATH="$PWD:$PYTHONPATH"
O=gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > $O/r02b_pytest_gpu.log
timeout 400 python bench.py --steps 20 --warmup 5 > $O/r02b_bench.json 2> $O/r02b_bench.err
B="--steps 20 --warmup 5 --e2e-reps 1 --no-cpu-baseline --no-ved"
MADGPU_GRAPH_VOXELS=0 timeout 200 python bench.py $B > $O/r02b_bench_nograph.json 2> $O/r02b_bench_nograph.err
MADGPU_PROLONG_CELL=0 timeout 200 python bench.py $B > $O/r02b_bench_oldprolong.json 2> $O/r02b_bench_oldprolong.err
for pf in 0 1 3 4; do MADGPU_PF_DIST=$pf timeout 200 python bench.py --steps 10 --warmup 3 --e2e-reps 0 --no-cpu-baseline --no-ved > $O/r02b_bench_pf$pf.json 2>/dev/null; done
timeout 200 python bench.py $B --smoother wj > $O/r02b_bench_wj.json 2> $O/r02b_bench_wj.err
timeout 200 python bench.py $B --smoother wj > $O/r02b_bench_wj2.json 2> $O/r02b_bench_wj2.err
timeout 200 python bench.py $B --size 256 > $O/r02b_bench_256.json 2> $O/r02b_bench_256.err
# launch list of the default cycle
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02b_gs_launches.csv \
   python bench.py --steps 2 --warmup 1 --e2e-reps 0 --no-cpu-baseline --no-ved > $O/r02b_ncu_launches.log 2>&1
python tools/ncu_summary.py launches $O/r02b_gs_launches.csv > $O/r02b_gs_launches.txt 2>&1; rm -f $O/r02b_gs_launches.csv
cap() { # name regex skip [keep]
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" --launch-skip $3 -c 1 -o $O/r02b_full_$1 -f \
     python bench.py --steps 1 --warmup 1 --e2e-reps 0 --no-cpu-baseline --no-ved > $O/r02b_ncu_full_$1.log 2>&1
  python tools/ncu_summary.py full $O/r02b_full_$1.ncu-rep > $O/r02b_full_$1.txt 2>&1
  [ "$4" = keep ] || rm -f $O/r02b_full_$1.ncu-rep
}
cap prolong_cell 'k_fast_prolong_cell' 3
cap res64c32 'k_fast_sweep<\(int\)3, double|k_fast_sweep<3, double' 0
# VED front-end launch list + full capture at 256^3
cat > $O/ved256.py <<'P'
import numpy as np, multigridanisotropicdiffusion_b200 as M
from multigridanisotropicdiffusion_b200 import phantom
img = phantom.vessel_phantom((256, 256, 256))[0].numpy()
f = M.VEDMultigridImageFilter("gs"); f.SetInput(img, phantom.VED_SPACING); f.SetOmega(1.5); f.SetDiffusionIterations(1); f.Update(); print(f.ved_stats, f.stats)
P
timeout 200 python $O/ved256.py > $O/r02b_ved256_plain.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r02b_ved_launches.csv python $O/ved256.py > $O/r02b_ved_ncu.log 2>&1
python tools/ncu_summary.py launches $O/r02b_ved_launches.csv > $O/r02b_ved_launches.txt 2>&1; rm -f $O/r02b_ved_launches.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_rg_|k_ved_update' -c 12 -o $O/r02b_full_ved -f python $O/ved256.py > $O/r02b_ncu_full_ved.log 2>&1
python tools/ncu_summary.py full $O/r02b_full_ved.ncu-rep > $O/r02b_full_ved.txt 2>&1; rm -f $O/r02b_full_ved.ncu-rep
du -sh $O
echo done
