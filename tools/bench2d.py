#!/usr/bin/env python
"""2-D path (BASELINE.json configs[0..1]: itk2DDiffusionTest_{WJ,GS}): V(nu,nu) cycles incl. the fp64 stop-test residual on a
synthetic square image with the tensor of test/itk2DDiffusionTest_WJ.cxx:66-73 scaled by a smooth random field, device-resident,
timed with CUDA events inside the library (madgpu_cycles_run).  One JSON line per (size, smoother); roofline of the level-0 sweep
on the canonical 24 B per pixel (u, f, u', three tensor planes).  MADGPU_FAST2D=0 selects the generic one-pixel-per-thread
kernels for an A/B.  Not the headline bench (bench.py); its lines go to profiles/."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="512,8192")
    ap.add_argument("--smoothers", default="gs,wj")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--nu", type=int, default=3)
    a = ap.parse_args()
    import numpy as np
    import torch

    from multigridanisotropicdiffusion_b200 import MadSolver

    peak = 6459.9
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:  # noqa: BLE001
        pass
    dev = torch.device("cuda", 0)
    for n in [int(x) for x in a.sizes.split(",")]:
        g = torch.Generator(device=dev).manual_seed(7)
        yy, xx = torch.meshgrid(torch.linspace(0, 6.28, n, device=dev), torch.linspace(0, 6.28, n, device=dev), indexing="ij")
        img = (128 + 100 * torch.sin(3 * xx) * torch.cos(2 * yy) + 5 * torch.randn((n, n), device=dev, generator=g)).float().contiguous()
        ang = 0.7 * torch.sin(xx) + 0.5 * torch.cos(yy)
        l1, l2 = 50.0 * (1.0 + 0.5 * torch.sin(2 * yy)), 1.0 + 0.5 * torch.cos(xx)
        c, s_ = torch.cos(ang), torch.sin(ang)
        D = [(l1 * c * c + l2 * s_ * s_).float().contiguous(), ((l1 - l2) * c * s_).float().contiguous(), (l1 * s_ * s_ + l2 * c * c).float().contiguous()]
        torch.cuda.synchronize()
        for sm in a.smoothers.split(","):
            s = MadSolver((n, n), (1.0, 1.0), time_step=0.1, smoother=MadSolver.GS if sm == "gs" else MadSolver.WJ, iterations_per_grid=a.nu,
                          tolerance=0.0, max_cycles=1 << 20)
            s.set_tensor_device([d.data_ptr() for d in D])
            s.cycles_begin(d_in=img.data_ptr())
            s.cycles_run(a.warmup)
            s.set_profiling(True)
            relres, dev_ms, st = s.cycles_run(a.steps)
            s.set_profiling(False)
            ms = dev_ms / a.steps
            sw_ms, sw_n = st["prof_ms"]["smooth0"], st["prof_launches"]["smooth0"]
            tile = s.gs_tile(0)
            line = {"metric": "2-D V-cycle Mpixel/s", "size": [n, n], "smoother": sm, "nu": a.nu, "ms_per_cycle": ms, "value": n * n / (ms * 1e-3) / 1e6,
                    "unit": "Mpixel/s", "steps": a.steps, "warmup": a.warmup, "levels": s.nlevels, "fast2d": os.environ.get("MADGPU_FAST2D", "1"),
                    "gs_tile": tile, "relres_last": float(relres[-1]),
                    "class_ms_per_cycle": {k: v / a.steps for k, v in st["prof_ms"].items() if v > 0},
                    "class_launches_per_cycle": {k: v / a.steps for k, v in st["prof_launches"].items() if v > 0},
                    "roofline": {"kernel": "level-0 sweep", "alg_bytes_per_launch": 24.0 * n * n, "ms_per_launch": sw_ms / max(sw_n, 1),
                                 "achieved_GBps": 24.0 * n * n / (sw_ms / max(sw_n, 1) * 1e-3) / 1e9 if sw_ms > 0 else None, "peak_GBps": peak,
                                 "note": "the 4-colour generic path needs 4 launches per sweep; launches counted as the library counts them"}}
            if line["roofline"]["achieved_GBps"]:
                line["roofline"]["frac"] = line["roofline"]["achieved_GBps"] / peak
            print(json.dumps(line), flush=True)
            s.close()
        del img, D
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
