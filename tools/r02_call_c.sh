#!/bin/bash
# round 2, GPU call c (2 GPUs): z-slab tests incl. FMG on slabs (never run on hardware before), bench at N = 2 with the slab parity
# solves and the weak-scaling volume, NCCL-halo A/B
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export PYTHONPATH="$PWD:$PYTHONPATH"
O=gpurun_out
(MADGPU_P2P_DEBUG=1 timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_zz_gpu_multi_fmg.py -m gpu -x -q -s 2>&1 | tail -60) > $O/r02c_multi2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29601 bench.py --gpus 2 --steps 10 --warmup 3 --weak > $O/r02c_bench_n2.json 2> $O/r02c_bench_n2.err; echo rc=$? >> $O/r02c_bench_n2.err
timeout 300 $TR --master-port 29602 bench.py --gpus 2 --steps 10 --warmup 3 --nccl-halo --e2e-reps 1 > $O/r02c_bench_n2_nccl.json 2> $O/r02c_bench_n2_nccl.err; echo rc=$? >> $O/r02c_bench_n2_nccl.err
MADGPU_P2P_WAIT=kernel timeout 300 $TR --master-port 29603 bench.py --gpus 2 --steps 10 --warmup 3 --e2e-reps 1 --no-slab-parity > $O/r02c_bench_n2_kwait.json 2> $O/r02c_bench_n2_kwait.err; echo rc=$? >> $O/r02c_bench_n2_kwait.err
nvidia-smi topo -m > $O/r02c_topo.txt 2>&1
echo done
