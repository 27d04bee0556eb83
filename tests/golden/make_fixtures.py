"""Decode the reference's own input fixtures into small committed files (run in the authoring
container only; /root/reference does not exist on the GPU box).

    python tests/golden/make_fixtures.py

lena.jpg is decoded with PIL (the reference uses ITK's JPEG reader; decoder differences are
irrelevant because the oracle and the GPU path both consume THIS decode), ved_test.zraw with zlib
per its MetaImage header (/root/reference/test/test_data/ved_test.mhd: MET_SHORT, 69 77 69,
spacing .3125 .3125 .5, little endian, zlib).
"""
import os
import zlib

import numpy as np
from PIL import Image

REF = "/root/reference/test/test_data"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    img = np.asarray(Image.open(os.path.join(REF, "lena.jpg")).convert("L"), dtype=np.uint8)
    assert img.shape == (512, 512)
    np.savez_compressed(os.path.join(HERE, "lena_512_u8.npz"), image=img)
    raw = zlib.decompress(open(os.path.join(REF, "ved_test.zraw"), "rb").read())
    vol = np.frombuffer(raw, dtype="<i2").reshape(69, 77, 69)  # z, y, x (DimSize = 69 77 69 is x y z)
    np.savez_compressed(os.path.join(HERE, "ved_test_i16.npz"), image=vol, spacing=np.array([0.3125, 0.3125, 0.5]))
    print("lena", img.shape, img.min(), img.max(), img.mean())
    print("ved_test", vol.shape, vol.min(), vol.max(), vol.mean())


if __name__ == "__main__":
    main()
