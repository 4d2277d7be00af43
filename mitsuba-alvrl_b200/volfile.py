"""Grid volume files ("VOL", version 3) as the reference's `gridvolume` plugin maps them (src/volume/gridvolume.cpp:217-287):
a writer for synthetic density grids and a numpy reader that mirrors what the library's own reader (csrc/hostio.h) does, for
scripts that want the array."""
import struct

import numpy as np

VOL_FLOAT32, VOL_FLOAT16, VOL_UINT8, VOL_QUANTIZED_DIRECTIONS = 1, 2, 3, 4          # gridvolume.cpp:101-106


def write_vol(path, density, bbox_min, bbox_max, voxel_type=VOL_FLOAT32):
    """density: [z][y][x] (x fastest in the file).  uint8 output quantises to round(255 * d), the inverse of the reference's
    density map i / 255 (gridvolume.cpp:212-215)."""
    d = np.asarray(density)
    if d.ndim != 3:
        raise ValueError("one-channel density grid [z][y][x] expected")
    with open(path, "wb") as f:
        f.write(b"VOL\x03")
        f.write(struct.pack("<iiiii", voxel_type, d.shape[2], d.shape[1], d.shape[0], 1))
        f.write(struct.pack("<6f", *[float(v) for v in bbox_min], *[float(v) for v in bbox_max]))
        if voxel_type == VOL_FLOAT32:
            f.write(np.ascontiguousarray(d, dtype="<f4").tobytes())
        elif voxel_type == VOL_UINT8:
            q = d if d.dtype == np.uint8 else np.clip(np.rint(np.asarray(d, np.float64) * 255.0), 0, 255).astype(np.uint8)
            f.write(np.ascontiguousarray(q).tobytes())
        else:
            raise ValueError("only float32 and uint8 density grids are written")


def read_vol(path):
    """-> (density float32 [z][y][x], bbox_min, bbox_max, voxel_type)"""
    with open(path, "rb") as f:
        head = f.read(48)
        if len(head) < 48 or head[:3] != b"VOL":
            raise ValueError("invalid volume data file (incorrect header identifier)")
        if head[3] != 3:
            raise ValueError("invalid volume data file (incorrect file version)")
        vtype, xres, yres, zres, channels = struct.unpack("<iiiii", head[4:24])
        box = struct.unpack("<6f", head[24:48])
        if channels != 1 or vtype not in (VOL_FLOAT32, VOL_UINT8):
            raise ValueError(f"unsupported volume data file (type={vtype}, channels={channels})")
        n = xres * yres * zres
        if vtype == VOL_FLOAT32:
            d = np.frombuffer(f.read(4 * n), dtype="<f4").astype(np.float32)
        else:
            d = np.frombuffer(f.read(n), dtype=np.uint8).astype(np.float32) / np.float32(255.0)
        if d.size != n:
            raise ValueError("invalid volume data file (fewer voxels than the header announces)")
    return d.reshape(zres, yres, xres), np.array(box[:3], np.float32), np.array(box[3:], np.float32), vtype
