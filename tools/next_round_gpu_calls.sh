#!/bin/sh
# The GPU calls to make first in the next round, in this order (each line is one `gpurun` call from the repo root).  Everything
# below was written or changed after round 1's GPU budget was spent and has only run on the emulated device.
#
#   sh tools/next_round_gpu_calls.sh 1      # new GPU tests (VED front-end, drop-ins, the reference's own test programs)
#   sh tools/next_round_gpu_calls.sh 2      # bench incl. the whole VED filter, launch list, ncu of the front-end kernels
#   sh tools/next_round_gpu_calls.sh 2b     # A/B of the opt-in fp64 residual with fp32-evaluated rows (MADGPU_RES64_COEF32=1)
#   sh tools/next_round_gpu_calls.sh 3      # 2 GPUs: slab tests incl. FMG on slabs, peer-halo handshake
#   sh tools/next_round_gpu_calls.sh 4      # 8 GPUs: the peer-store halo that hung in round 1, with logging and a short leash
set -e
G=/usr/local/graft/bin/gpurun
case "$1" in
1) $G --timeout 1500 -- 'python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02_pytest_gpu.log; python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1' ;;
2) $G --timeout 1500 -- 'python bench.py --steps 10 --warmup 3 --ved > gpurun_out/r02_bench_ved.json 2> gpurun_out/r02_bench_ved.err &&
     ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_ved_launches.csv \
         python -c "
import numpy as np, multigridanisotropicdiffusion_b200 as M
from multigridanisotropicdiffusion_b200 import phantom
img = phantom.vessel_phantom((256, 256, 256))[0].numpy()
f = M.VEDMultigridImageFilter(\"gs\"); f.SetInput(img, phantom.VED_SPACING); f.SetOmega(1.5); f.SetDiffusionIterations(1); f.Update(); print(f.ved_stats)
" > gpurun_out/r02_ved_ncu.log 2>&1' ;;
2b) $G --timeout 900 -- 'python bench.py --steps 10 --warmup 3 --e2e-reps 1 --no-cpu-baseline > gpurun_out/r02_bench_res64_default.json 2>/dev/null; MADGPU_RES64_COEF32=1 python bench.py --steps 10 --warmup 3 --e2e-reps 1 --no-cpu-baseline > gpurun_out/r02_bench_res64_coef32.json 2>/dev/null' ;;
3) $G --gpus 2 --timeout 1200 -- 'MADGPU_P2P_DEBUG=1 python -m pytest tests/test_gpu_multi.py tests/test_zz_gpu_multi_fmg.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r02_multi2.log' ;;
4) $G --gpus 8 --timeout 600 -- 'MADGPU_P2P_DEBUG=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29577 \
     bench.py --gpus 8 --steps 5 --warmup 3 --peer-halo --e2e-reps 0 > gpurun_out/r02_scale8_peer.json 2> gpurun_out/r02_scale8_peer.err; echo rc=$? >> gpurun_out/r02_scale8_peer.err' ;;
*) echo "usage: $0 {1|2|2b|3|4}"; exit 1 ;;
esac
