#!/bin/bash
# round 2, GPU call d (8 GPUs, short leash): the peer-store halo at 8 ranks with the bounded wait + debug log, slab parity, 1024^3
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export PYTHONPATH="$PWD:$PYTHONPATH"
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
MADGPU_BENCH_PEER8=1 MADGPU_P2P_DEBUG=1 timeout 330 $TR --master-port 29701 bench.py --gpus 8 --steps 10 --warmup 3 --e2e-reps 1 > $O/r02d_bench_n8_peer.json 2> $O/r02d_bench_n8_peer.err; echo rc=$? >> $O/r02d_bench_n8_peer.err
nvidia-smi --query-gpu=index,name,memory.used --format=csv > $O/r02d_smi.txt 2>&1
echo done
