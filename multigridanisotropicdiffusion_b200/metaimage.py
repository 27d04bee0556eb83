"""MetaImage (.mhd + .raw / .zraw) reader and writer for the volumes the reference's tests use
(/root/reference/test/test_data/ved_test.mhd: MET_SHORT, zlib-compressed .zraw).  The reference reads and writes them with
itk::ImageFileReader / ImageFileWriter (test/itkVEDTest_GS.cxx:35-38, 119-124); this module stands in for those two classes
on the Python side so that the reference's test programs can be run through the mirror filters (examples/).  Host-side file
I/O only -- nothing here touches the GPU path.

Arrays are (nz, ny, nx) / (ny, nx), x fastest, i.e. DimSize reversed.  The header fields that describe geometry
(ElementSpacing, Offset, TransformMatrix, ...) are returned in `meta` and written back unchanged, which is what
test/itkVEDTest_GS.cxx:108-117 achieves with ChangeInformationImageFilter (the filter itself drops the direction).
"""
from __future__ import annotations

import os
import zlib

import numpy as np

_TYPES = {"MET_UCHAR": np.uint8, "MET_CHAR": np.int8, "MET_USHORT": np.uint16, "MET_SHORT": np.int16, "MET_UINT": np.uint32,
          "MET_INT": np.int32, "MET_FLOAT": np.float32, "MET_DOUBLE": np.float64}
_NAMES = {np.dtype(v): k for k, v in _TYPES.items()}


def read(path):
    """-> (array, meta).  meta: dict of the header (strings), plus 'spacing' (sx, sy[, sz]) as floats."""
    meta = {}
    with open(path, "r") as f:
        for line in f:
            if "=" not in line:
                continue
            k, v = line.split("=", 1)
            meta[k.strip()] = v.strip()
    if meta.get("ObjectType", "Image") != "Image":
        raise ValueError(f"{path}: ObjectType {meta.get('ObjectType')} is not an image")
    ndims = int(meta["NDims"])
    size = [int(s) for s in meta["DimSize"].split()]
    if len(size) != ndims:
        raise ValueError(f"{path}: DimSize has {len(size)} entries, NDims = {ndims}")
    if int(meta.get("ElementNumberOfChannels", "1")) != 1:
        raise ValueError(f"{path}: multi-channel images are not supported")
    try:
        dt = np.dtype(_TYPES[meta["ElementType"]])
    except KeyError:
        raise ValueError(f"{path}: unsupported ElementType {meta.get('ElementType')}") from None
    dt = dt.newbyteorder(">" if meta.get("BinaryDataByteOrderMSB", meta.get("ElementByteOrderMSB", "False")) == "True" else "<")
    data_file = meta["ElementDataFile"]
    if data_file == "LOCAL":
        raise ValueError(f"{path}: data embedded in the header (.mha) is not supported")
    raw = open(os.path.join(os.path.dirname(os.path.abspath(path)), data_file), "rb").read()
    if meta.get("CompressedData", "False") == "True":
        raw = zlib.decompress(raw)
    n = int(np.prod(size))
    if len(raw) < n * dt.itemsize:
        raise ValueError(f"{path}: {len(raw)} bytes of data, {n * dt.itemsize} expected")
    arr = np.frombuffer(raw, dtype=dt, count=n).reshape(size[::-1]).astype(dt.newbyteorder("="))
    meta["spacing"] = tuple(float(s) for s in meta.get("ElementSpacing", " ".join(["1"] * ndims)).split())
    return arr, meta


def write(path, array, meta=None, compressed=True):
    """Write `array` as <path> (.mhd) + a .zraw / .raw file next to it.  Geometry fields of `meta` (as returned by read) are kept."""
    a = np.ascontiguousarray(array)
    if a.dtype not in _NAMES:
        raise ValueError(f"unsupported pixel type {a.dtype}")
    meta = dict(meta or {})
    base = os.path.splitext(os.path.basename(path))[0]
    data_file = base + (".zraw" if compressed else ".raw")
    raw = a.astype(a.dtype.newbyteorder("<")).tobytes()
    payload = zlib.compress(raw) if compressed else raw
    spacing = meta.pop("spacing", None)
    hdr = [("ObjectType", "Image"), ("NDims", str(a.ndim)), ("BinaryData", "True"), ("BinaryDataByteOrderMSB", "False"),
           ("CompressedData", "True" if compressed else "False")]
    if compressed:
        hdr.append(("CompressedDataSize", str(len(payload))))
    for k in ("TransformMatrix", "Offset", "CenterOfRotation", "AnatomicalOrientation"):
        if k in meta:
            hdr.append((k, meta[k]))
    if spacing is not None:
        hdr.append(("ElementSpacing", " ".join(repr(float(s)) for s in spacing)))
    elif "ElementSpacing" in meta:
        hdr.append(("ElementSpacing", meta["ElementSpacing"]))
    hdr += [("DimSize", " ".join(str(s) for s in a.shape[::-1])), ("ElementType", _NAMES[a.dtype]), ("ElementDataFile", data_file)]
    d = os.path.dirname(os.path.abspath(path))
    with open(os.path.join(d, data_file), "wb") as f:
        f.write(payload)
    with open(path, "w") as f:
        for k, v in hdr:
            f.write(f"{k} = {v}\n")
