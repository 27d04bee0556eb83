#!/bin/bash
# round 2, GPU call t (8 GPUs, short leash): bench.py --gpus 8 on the final kernels: 512^3 on 8 slabs, slab parity, 1024^3
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export PYTHONPATH="$PWD:$PYTHONPATH"
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29731 bench.py --gpus 8 --steps 10 --warmup 3 --e2e-reps 1 --no-n1-1024 > $O/r02t_bench_n8.json 2> $O/r02t_bench_n8.err; echo rc=$? >> $O/r02t_bench_n8.err
echo done
