#!/bin/bash
# round 2, GPU call p (4 GPUs): the multi-GPU test files (their 4-GPU cases run here), bench.py --gpus 4, and the newest single-GPU test
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export PYTHONPATH="$PWD:$PYTHONPATH"
O=gpurun_out
(timeout 300 python -m pytest tests/test_gpu_solve.py -m gpu -x -q -k "survives" 2>&1 | tail -4) > $O/r02p_pytest_new.log
(timeout 500 python -m pytest tests/test_gpu_multi.py tests/test_zz_gpu_multi_fmg.py -m gpu -x -q 2>&1 | tail -8) > $O/r02p_multi4.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29641 bench.py --gpus 4 --steps 10 --warmup 3 --e2e-reps 2 > $O/r02p_bench_n4.json 2> $O/r02p_bench_n4.err; echo rc=$? >> $O/r02p_bench_n4.err
echo done
