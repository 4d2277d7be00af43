"""The reference's image metric: `mtsutil rms` (src/utils/rms.cpp:25-113).

    rms(image, reference, gamma=1.0, robust_fraction=0.0, relative=False)

follows the utility's arithmetic: every channel value of both images goes through pow(v, 1 / gamma); the deviation of an
entry is `sample - reference`, or `(sample - reference) / reference` in the relative variant with zero-reference entries
masked to 0 (88-91); `robust_fraction` drops that fraction of the entries at both extremes of the sorted deviations (43-51,
95-98); the squared deviations are sorted before they are summed in double (99-108); the result is sqrt(sum / count).
Host-side tooling next to the ground-truth generator (alvrl_volpath_render): it compares images, it is not on the hot path."""
import numpy as np


def rms(image, reference, gamma=1.0, robust_fraction=0.0, relative=False):
    a = np.asarray(image, dtype=np.float32).reshape(-1).astype(np.float64)
    b = np.asarray(reference, dtype=np.float32).reshape(-1).astype(np.float64)
    if a.shape != b.shape:
        raise ValueError("Images must have the same size and number of channels!")            # rms.cpp:33-36
    n = a.size
    drop = int(0.5 + n * robust_fraction) if robust_fraction else 0
    if 2 * drop >= n and drop:
        raise ValueError("robustFraction: dropping more elements than there are available!")   # rms.cpp:48-50
    with np.errstate(invalid="ignore", divide="ignore"):
        a = np.power(a, 1.0 / gamma)
        b = np.power(b, 1.0 / gamma)
        diffs = np.where(b == 0, 0.0, (a - b) / np.where(b == 0, 1.0, b)) if relative else a - b
    if drop:
        diffs = np.sort(diffs)[drop:n - drop]
    sq = np.sort(diffs * diffs)
    acc = 0.0
    for chunk in np.array_split(sq, max(1, len(sq) // 65536)):        # ascending order, double accumulation (rms.cpp:103-107)
        acc += float(np.sum(chunk))
    return float(np.sqrt(acc / (n - 2 * drop)))


def main(argv=None):
    """rms <image.npy> <reference.npy> [gamma] [robustFraction] [relative]  -- the argument order of the utility (25-56)"""
    import sys
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) < 2:
        print("rms <image.npy> <reference.npy> [gamma] [robustFraction] [relative]")
        return 1
    gamma = float(argv[2]) if len(argv) > 2 else 1.0
    frac = float(argv[3]) if len(argv) > 3 else 0.0
    print(rms(np.load(argv[0]), np.load(argv[1]), gamma, frac, relative=len(argv) > 4))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
