"""Golden vectors produced BY THE REFERENCE'S OWN CODE: the unmodified headers under /root/reference/include
compiled against the stand-in ITK of oracle/shim (oracle/_ref/libmadref.so, `make -C oracle ref`).

    python tests/golden/make_golden.py            # authoring container only (needs /root/reference to build _ref)

The reference's test programs (test/itk2DDiffusionTest_{WJ,GS}.cxx, test/itkVEDTest_GS.cxx) assert nothing, so
these files are what pins parity: tests/test_cpu_golden.py checks the oracle against them, tests/test_gpu_golden.py
the CUDA path.  To keep the fixtures small the big outputs are stored sub-sampled together with full-image
statistics in double precision; small synthetic cases are stored whole.

Cases (parameters of the reference's tests):
  lena_{wj,gs}_{v,fmg,s}   512x512, D = diag(50, 30), dt .1, nu 2, tol 1e-10, MaxCycles 100, float pixels
  ved_gs_v                 ved_test.mhd (69x77x69 short, spacing .3125/.3125/.5), analytic VED-form tensor,
                           GS, nu 3, dt .1, 4 time steps, tol 1e-10 (DiffusionStep of itkVEDTest_GS.cxx)
  small2d_*, small3d_*     49x33 and 23x25x27 random SPD tensor fields (cross terms, mixed centring), whole outputs
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref as R  # noqa: E402
from util import load_lena, load_ved_test, random_image, random_spd_tensor  # noqa: E402


def stats(a):
    a = np.asarray(a, dtype=np.float64)
    return np.array([np.linalg.norm(a), a.mean(), a.min(), a.max(), np.abs(np.diff(a, axis=-1)).sum()])


def save(name, out, cycles, log, smoother_mode, sub, **meta):
    rr = R.relres_per_cycle(log, smoother_mode)
    hist = np.full((len(rr), max(len(r) for r in rr)), np.nan)
    for i, r in enumerate(rr):
        hist[i, :len(r)] = r
    sl = tuple(slice(None, None, sub) for _ in out.shape)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), sample=out[sl], sub=sub, stats=stats(out), cycles=np.array(cycles),
                        relres=hist, **meta)
    print(f"{name}: cycles {cycles} final relres {[r[-1] for r in rr]} norm {stats(out)[0]:.6f}", flush=True)


def main():
    if not R.available():
        R.build()
    t0 = time.time()
    # ---- reference 2-D tests ----
    lena = load_lena().astype(np.float64)
    T = np.zeros(lena.shape + (3,))
    T[..., 0] = 50.0
    T[..., 2] = 30.0
    for sm, sname in ((1, "wj"), (0, "gs")):
        for cyc, cname in ((0, "v"), (1, "fmg"), (2, "s")):
            out, cycles, log = R.run_filter(lena, (1.0, 1.0), T, smoother=sm, cycle=cyc, nu=2, time_step=0.1, tolerance=1e-10,
                                            max_cycles=100, number_of_steps=1, pixel="float")
            save(f"ref_lena_{sname}_{cname}", out, cycles, log, cyc == 2, 4)
            print(f"   {time.time() - t0:.0f}s", flush=True)
    # ---- small synthetic cases, whole outputs ----
    for shape, sp, tag in (((49, 33), (0.7, 1.3), "small2d"), ((23, 25, 27), (0.3125, 0.3125, 0.5), "small3d")):
        Ts = random_spd_tensor(shape, seed=2).astype(np.float64)
        img = random_image(shape, seed=5).astype(np.float64)
        for sm, sname in ((1, "wj"), (0, "gs")):
            for cyc, cname in ((0, "v"), (1, "fmg")):
                out, cycles, log = R.run_filter(img, sp, Ts, smoother=sm, cycle=cyc, nu=2, time_step=0.1, tolerance=1e-10, max_cycles=100,
                                                number_of_steps=2, pixel="double")
                save(f"ref_{tag}_{sname}_{cname}", out, cycles, log, False, 1)
    # ---- reference 3-D test: the diffusion step of itkVEDTest_GS.cxx ----
    from multigridanisotropicdiffusion_b200 import phantom
    vol, sp = load_ved_test()
    _, D = phantom.vessel_phantom(vol.shape, spacing=sp)
    Tv = phantom.planes_to_aos(D).numpy().astype(np.float64)
    out, cycles, log = R.run_filter(vol.astype(np.float64), sp, Tv, smoother=0, cycle=0, nu=3, time_step=0.1, tolerance=1e-10, max_cycles=100,
                                    number_of_steps=4, pixel="double")
    save("ref_ved_gs_v", out, cycles, log, False, 3)
    print(f"done in {time.time() - t0:.0f}s")


if __name__ == "__main__":
    main()
