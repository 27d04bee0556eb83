#!/bin/bash
# round 2, GPU call g (1 GPU): full GPU suite with durations (new: 512^3 properties, 512-wide planes, 2-D streaming kernels, captured
# Gauss-Jordan), default bench, A/B: marching restriction off, fp64-residual register cap, set-up trace, VED probe, 2-D bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export PYTHONPATH="$PWD:$PYTHONPATH"
O=gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q --durations=15 2>&1 | tail -40) > $O/r02g_pytest_gpu.log
timeout 500 python bench.py --steps 20 --warmup 5 > $O/r02g_bench.json 2> $O/r02g_bench.err
B="--steps 20 --warmup 5 --e2e-reps 2 --no-cpu-baseline --no-ved"
MADGPU_RESTRICT_CELL=0 timeout 200 python bench.py $B > $O/r02g_bench_oldrestrict.json 2> $O/r02g_bench_oldrestrict.err
MADGPU_RES64_MINB=3 timeout 200 python bench.py $B > $O/r02g_bench_res64minb3.json 2> $O/r02g_bench_res64minb3.err
MADGPU_SETUP_TRACE=1 timeout 200 python tools/e2e_probe.py > $O/r02g_e2e_probe.log 2>&1
timeout 300 python tools/ved_probe.py 512 > $O/r02g_ved_probe.log 2>&1
timeout 200 python tools/bench2d.py > $O/r02g_bench2d.jsonl 2> $O/r02g_bench2d.err
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_fast_restrict_cell" --launch-skip 6 -c 1 -o $O/r02g_full_restrict_cell -f \
   python bench.py --steps 1 --warmup 1 --e2e-reps 0 --no-cpu-baseline --no-ved > $O/r02g_ncu_full_restrict_cell.log 2>&1
python tools/ncu_summary.py full $O/r02g_full_restrict_cell.ncu-rep > $O/r02g_full_restrict_cell.txt 2>&1; rm -f $O/r02g_full_restrict_cell.ncu-rep
du -sh $O
echo done
