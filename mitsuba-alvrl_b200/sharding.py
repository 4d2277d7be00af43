"""Slice sharding for the multi-GPU path: the Clustering object, the R rows and the pixels of a slice depend only on that
slice, so rank r owns a contiguous range of slice ids (alvrl_set_slice_range).  Slices differ a lot in size --
their ids come from the array order of the slice builder's heap (Preprocessor.cpp:1400-1417) -- so the ranges are cut by
PIXELS, not by slice count: rows of R, refinement work and render work of a slice all grow with its pixel count."""
import numpy as np

NO_SLICE = 0xFFFFFFFF


def slice_sizes(pixel_to_slice, num_slices):
    """pixels per slice from the pixel -> slice map (misses carry NO_SLICE)"""
    p2s = np.asarray(pixel_to_slice).reshape(-1)
    return np.bincount(p2s[p2s != NO_SLICE].astype(np.int64), minlength=num_slices)[:num_slices]


def balanced_ranges(sizes, world):
    """[(begin, end)] per rank: contiguous, covering, boundaries at the slice whose cumulative size is closest to r / world"""
    sizes = np.asarray(sizes, dtype=np.int64)
    S = len(sizes)
    cum = np.concatenate([[0], np.cumsum(sizes)])
    bounds = [0]
    for r in range(1, world):
        target = cum[-1] * r / world
        b = int(np.searchsorted(cum, target))
        b = min(b, S)
        if b > 0 and abs(cum[b - 1] - target) <= abs(cum[b] - target):
            b -= 1
        bounds.append(min(S, max(b, bounds[-1])))
    bounds.append(S)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]
