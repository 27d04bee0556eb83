import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from multigridanisotropicdiffusion_b200 import MadSolver, phantom
n=512; shape=(n,n,n); dev=torch.device('cuda',0)
img, D = phantom.vessel_phantom(shape, device=dev)
T_h = torch.empty(shape+(6,), dtype=torch.float32, pin_memory=True); T_h.copy_(phantom.planes_to_aos(D))
img_h = torch.empty(shape, dtype=torch.float32, pin_memory=True); img_h.copy_(img)
out_h = torch.empty(shape, dtype=torch.float32, pin_memory=True)
del D; torch.cuda.synchronize()
s = MadSolver(shape, phantom.VED_SPACING, time_step=0.1, smoother=0, iterations_per_grid=3, tolerance=1e-10, max_cycles=100, number_of_steps=4)
s.set_profiling(True)
for rep in range(4):
    t0=time.perf_counter(); s.set_tensor(T_h.numpy()); torch.cuda.synchronize(); t1=time.perf_counter()
    s.solve(img_h.numpy(), out=out_h.numpy()); torch.cuda.synchronize(); t2=time.perf_counter()
    st=s.last_stats
    print(rep, f"set_tensor {t1-t0:.3f}s solve {t2-t1:.3f}s", "solve_ms", round(st['solve_ms'],1), "setup", round(st['setup_ms'],1), "h2d", round(st['h2d_ms'],1), "d2h", round(st['d2h_ms'],1), st['cycles_per_step'], {k:round(v,1) for k,v in st['prof_ms'].items() if v>0})
