"""MetaImage reader / writer (multigridanisotropicdiffusion_b200/metaimage.py): round trips and, where /root/reference exists, the
reference's own fixtures against the committed decodes the rest of the suite uses."""
import os

import numpy as np
import pytest

from multigridanisotropicdiffusion_b200 import metaimage
from util import load_ved_test

REF = "/root/reference/test/test_data"


@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.float32, np.float64])
@pytest.mark.parametrize("compressed", [True, False])
def test_round_trip(tmp_path, dtype, compressed):
    rng = np.random.default_rng(0)
    a = (rng.normal(size=(5, 6, 7)) * 50).astype(dtype)
    meta = {"spacing": (0.3125, 0.3125, 0.5), "Offset": "-13.9881 -27.1641 -52.1181", "TransformMatrix": "-1 0 0 0 -1 0 0 0 1",
            "AnatomicalOrientation": "LPI"}
    p = str(tmp_path / "vol.mhd")
    metaimage.write(p, a, meta, compressed=compressed)
    b, m = metaimage.read(p)
    assert b.dtype == a.dtype and b.shape == a.shape
    np.testing.assert_array_equal(a, b)
    assert m["spacing"] == (0.3125, 0.3125, 0.5) and m["Offset"] == meta["Offset"] and m["TransformMatrix"] == meta["TransformMatrix"]
    assert m["DimSize"] == "7 6 5"  # x y z
    assert os.path.exists(str(tmp_path / ("vol.zraw" if compressed else "vol.raw")))


def test_two_dimensional_and_errors(tmp_path):
    a = np.arange(12, dtype=np.int16).reshape(3, 4)
    p = str(tmp_path / "img.mhd")
    metaimage.write(p, a)
    b, m = metaimage.read(p)
    np.testing.assert_array_equal(a, b)
    assert m["spacing"] == (1.0, 1.0)
    with pytest.raises(ValueError):
        metaimage.write(p, a.astype(np.complex64))
    open(str(tmp_path / "bad.mhd"), "w").write("ObjectType = Image\nNDims = 2\nDimSize = 4 3\nElementType = MET_LONG_LONG\nElementDataFile = img.zraw\n")
    with pytest.raises(ValueError):
        metaimage.read(str(tmp_path / "bad.mhd"))


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "ved_test.mhd")), reason="/root/reference not present")
def test_reference_fixtures():
    vol, meta = metaimage.read(os.path.join(REF, "ved_test.mhd"))
    want, sp = load_ved_test()
    np.testing.assert_array_equal(vol, want)
    assert meta["spacing"] == sp and vol.dtype == np.int16
    vol2, meta2 = metaimage.read(os.path.join(REF, "ved_test_2.mhd"))
    assert vol2.shape == (119, 140, 134) and vol2.dtype == np.int16 and (vol2.min(), vol2.max()) == (-1024, 744)
