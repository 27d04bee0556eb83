#!/bin/bash
# round 2, GPU call m (1 GPU): BASELINE.json configs[4]'s volume, 1024^3, on ONE GPU (the strong-scaling denominator of the 8-GPU figure)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export PYTHONPATH="$PWD:$PYTHONPATH"
O=gpurun_out
timeout 500 python bench.py --size 1024 --steps 5 --warmup 3 --e2e-reps 0 --no-cpu-baseline --no-ved > $O/r02m_bench_1024_n1.json 2> $O/r02m_bench_1024_n1.err; echo rc=$? >> $O/r02m_bench_1024_n1.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv > $O/r02m_smi.txt 2>&1
echo done
