"""Golden vectors of the WHOLE VED filter produced by the reference's own code: /root/reference/include/
itkVEDMultigridImageFilter.{h,hxx} compiled unmodified into oracle/_ref/libmadref.so (`make -C oracle ref`), driven with the
parameters of test/itkVEDTest_GS.cxx (:64-101) on the reference's own volume test/test_data/ved_test.mhd.

    python tests/golden/make_golden_ved.py        # authoring container only (needs /root/reference to build _ref)

The Hessian filter and the eigen-solver underneath are third-party (ITK / VXL, absent): in that build they are the stand-ins
of oracle/shim/mini_itk_ved.h, so these vectors pin the reference's own code around them (vesselness, arg-max over scales,
tensor synthesis, DiffusionStep, casts), not ITK's Hessian -- see the header of oracle/ved_oracle.c.

  ref_vedfilter_gs_v   ved_test (69x77x69, spacing .3125/.3125/.5), 5 scales .3 .482 .775 1.245 2.0, alpha = beta = .5, gamma 5,
                       epsilon .01, sensitivity 10, omega 1.5, Iterations 1, DiffusionIterations 4, dt .1, tol 1e-10, GS, nu 3,
                       V-cycles; run with double pixels (sample every 3rd voxel + statistics) and with short pixels (as the
                       test does: the truncated output, whole, as int16); tensor: statistics + sample
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
from oracle import ved as V  # noqa: E402
from util import load_ved_test  # noqa: E402

VED_TEST = dict(alpha=0.5, beta=0.5, gamma=5.0, epsilon=0.01, omega=1.5, sensitivity=10.0, iterations=1, diffusion_iterations=4,
                smoother=0, cycle=0, time_step=0.1, tolerance=1e-10, iterations_per_grid=3)


def stats(a):
    a = np.asarray(a, dtype=np.float64)
    return np.array([np.linalg.norm(a), a.mean(), a.min(), a.max(), np.abs(np.diff(a, axis=-1)).sum()])


def main():
    if not R.available():
        R.build()
    vol, sp = load_ved_test()
    t0 = time.time()
    out_d, T = R.run_ved_filter(vol.astype(np.float64), sp, V.DEFAULT_SCALES, pixel="double", **VED_TEST)
    print(f"double pixels: {time.time() - t0:.0f}s  norm {np.linalg.norm(out_d):.6f}", flush=True)
    out_s, _ = R.run_ved_filter(vol.astype(np.float64), sp, V.DEFAULT_SCALES, pixel="short", **VED_TEST)
    print(f"short pixels: {time.time() - t0:.0f}s", flush=True)
    assert np.array_equal(out_s, np.trunc(out_d))  # a short input is exact in double: the two runs differ by the final cast only
    sub = 3
    sl = (slice(None, None, sub),) * 3
    np.savez_compressed(os.path.join(HERE, "ref_vedfilter_gs_v.npz"), sample=out_d[sl], sub=sub, stats=stats(out_d),
                        out_short=out_s.astype(np.int16), tensor_sample=T[sl], tensor_stats=np.stack([stats(T[..., k]) for k in range(6)]))
    print("tensor stats", np.stack([stats(T[..., k]) for k in range(6)])[:, 0])


if __name__ == "__main__":
    main()
