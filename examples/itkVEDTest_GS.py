#!/usr/bin/env python
"""The reference's test program test/itkVEDTest_GS.cxx on the B200 path, line for line through the mirror filter:

    python examples/itkVEDTest_GS.py {v|fmg|s} [input.mhd [output.mhd]]

Defaults to the reference's own volume test/test_data/ved_test.mhd when /root/reference is present, else to the copy of it
under tests/golden.  Like the reference's test it writes the enhanced volume (short pixels, geometry of the input) and
asserts nothing; parity is what tests/test_gpu_ved.py checks."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import multigridanisotropicdiffusion_b200 as M  # noqa: E402
from multigridanisotropicdiffusion_b200 import metaimage  # noqa: E402


def main(argv):
    mode = argv[1] if len(argv) > 1 else "v"
    ref = "/root/reference/test/test_data/ved_test.mhd"
    if len(argv) > 2:
        image, meta = metaimage.read(argv[2])
    elif os.path.exists(ref):
        image, meta = metaimage.read(ref)  # :27-38
    else:
        z = np.load(os.path.join(ROOT, "tests", "golden", "ved_test_i16.npz"))
        image, meta = z["image"], {"spacing": tuple(float(s) for s in z["spacing"])}
    out_path = argv[3] if len(argv) > 3 else "ved_test_out.mhd"
    print(f"input {image.shape[::-1]} {image.dtype} spacing {meta['spacing']}")

    f = M.VEDMultigridImageFilter(M.MultigridGaussSeidelSmoother)  # :47-48
    f.SetCycle({"fmg": f.FMG, "s": f.SMOOTHER}.get(mode, f.VCYCLE))  # :55-63
    f.SetDiffusionIterationsPerGrid(3)  # :64
    f.SetInput(image, meta["spacing"])  # :66
    f.SetVerbose(True)
    f.SetScales([0.300, 0.482, 0.775, 1.245, 2.000])  # :71-80
    f.SetAlpha(0.5)
    f.SetBeta(0.5)
    f.SetGamma(5.0)
    f.SetEpsilon(0.01)
    f.SetSensitivity(10.0)
    f.SetIterations(1)
    f.SetTolerance(1e-10)
    f.SetTimeStep(0.1)
    f.SetDiffusionIterations(4)
    f.SetOmega(1.5)  # :95
    f.Update()  # :102
    out = f.GetOutput()
    print(f"cycles per diffusion step {f.stats['cycles_per_step']}  front-end {f.ved_stats}")
    metaimage.write(out_path, out, meta)  # :119-124 (the input's geometry is kept, :108-117)
    print("wrote", out_path)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
