#!/bin/bash
# round 2, GPU call s (2 GPUs): the multi-GPU test files and bench.py --gpus 2 on the final kernels (marching restriction on slabs)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export PYTHONPATH="$PWD:$PYTHONPATH"
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29661 bench.py --gpus 2 --steps 10 --warmup 3 --e2e-reps 2 > $O/r02s_bench_n2.json 2> $O/r02s_bench_n2.err; echo rc=$? >> $O/r02s_bench_n2.err
echo done
