#!/usr/bin/env python
"""The reference's test programs test/itk2DDiffusionTest_{WJ,GS}.cxx on the B200 path through the mirror filter:

    python examples/itk2DDiffusionTest.py {wj|gs} {v|fmg|s} [input image [output.npy]]

Constant tensor Dxx = 50, Dyy = 30, Dxy = 0, two iterations per grid, dt 0.1, one step, tolerance 1e-10 (:61-97).  The input
defaults to test/test_data/lena.jpg (decoded with PIL) or the committed decode of it under tests/golden."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import multigridanisotropicdiffusion_b200 as M  # noqa: E402


def main(argv):
    smoother = argv[1] if len(argv) > 1 else "gs"
    mode = argv[2] if len(argv) > 2 else "v"
    ref = "/root/reference/test/test_data/lena.jpg"
    if len(argv) > 3 or os.path.exists(ref):
        from PIL import Image
        image = np.asarray(Image.open(argv[3] if len(argv) > 3 else ref).convert("L"), dtype=np.uint8)
    else:
        image = np.load(os.path.join(ROOT, "tests", "golden", "lena_512_u8.npz"))["image"]
    image = image.astype(np.float32)  # the test casts the unsigned char image to float (:37-45)
    tensor = np.zeros(image.shape + (3,), dtype=np.float32)  # :61-75
    tensor[..., 0] = 50.0
    tensor[..., 2] = 30.0
    f = M.MultigridAnisotropicDiffusionImageFilter(smoother)
    f.SetInput(image, (1.0, 1.0))
    f.SetDiffusionTensor(tensor)
    f.SetIterationsPerGrid(2)  # :91-97
    f.SetTimeStep(0.1)
    f.SetNumberOfSteps(1)
    f.SetMaxCycles(100)
    f.SetTolerance(1e-10)
    f.SetVerbose(True)
    f.SetCycle({"fmg": f.FMG, "s": f.SMOOTHER}.get(mode, f.VCYCLE))
    f.Update()
    out = f.GetOutput()
    print(f"cycles {f.stats['cycles_per_step']} final relative residual {f.stats['final_relres']}")
    np.save(argv[4] if len(argv) > 4 else "lena_out.npy", out)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
