#!/usr/bin/env python
"""Where the time of one VEDMultigridImageFilter.Update() at 512^3 goes (wall clock, stream-synchronised): context creation, the
front-end, the diffusion step, the copies.  Diagnostic for bench.py's `ved_filter` figure."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import multigridanisotropicdiffusion_b200 as M
from multigridanisotropicdiffusion_b200 import phantom

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
img = phantom.vessel_phantom((n, n, n), device=torch.device("cuda", 0))[0].cpu().numpy()
torch.cuda.empty_cache()
for rep in range(3):
    f = M.VEDMultigridImageFilter("gs", 0)
    f.SetInput(img, phantom.VED_SPACING)
    f.SetOmega(1.5)
    f.SetDiffusionIterationsPerGrid(3)
    f.SetDiffusionIterations(4)
    f.SetTolerance(1e-10)
    t0 = time.perf_counter()
    f.Update()
    t1 = time.perf_counter()
    print(f"rep {rep}: Update {t1 - t0:.3f} s  ved {f.ved_stats}  solver setup_ms {f.stats['setup_ms']:.1f} solve_ms {f.stats['solve_ms']:.1f} "
          f"cycles {f.stats['cycles_per_step']} timing {getattr(f, 'timing', None)}", flush=True)
    f.close()
