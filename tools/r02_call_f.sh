#!/bin/bash
# round 2, GPU call f (1 GPU): full GPU suite incl. the 2-D streaming kernels, default bench, A/B of the streaming threshold, set-up trace,
# VED filter breakdown, 2-D bench + ncu of k2_sweep
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export PYTHONPATH="$PWD:$PYTHONPATH"
O=gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > $O/r02f_pytest_gpu.log
timeout 500 python bench.py --steps 20 --warmup 5 > $O/r02f_bench.json 2> $O/r02f_bench.err
B="--steps 20 --warmup 5 --e2e-reps 1 --no-cpu-baseline --no-ved"
for nx in 16 32; do MADGPU_FAST_MIN_NX=$nx timeout 200 python bench.py $B > $O/r02f_bench_minnx$nx.json 2> $O/r02f_bench_minnx$nx.err; done
MADGPU_SETUP_TRACE=1 timeout 200 python tools/e2e_probe.py > $O/r02f_e2e_probe.log 2>&1
timeout 300 python tools/ved_probe.py 512 > $O/r02f_ved_probe.log 2>&1
timeout 200 python tools/bench2d.py > $O/r02f_bench2d_fast.jsonl 2> $O/r02f_bench2d_fast.err
MADGPU_FAST2D=0 timeout 200 python tools/bench2d.py > $O/r02f_bench2d_generic.jsonl 2> $O/r02f_bench2d_generic.err
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k2_sweep" -c 6 -o $O/r02f_full_k2 -f \
   python tools/bench2d.py --sizes 8192 --steps 1 --warmup 0 > $O/r02f_ncu_full_k2.log 2>&1
python tools/ncu_summary.py full $O/r02f_full_k2.ncu-rep > $O/r02f_full_k2.txt 2>&1; rm -f $O/r02f_full_k2.ncu-rep
du -sh $O
echo done
