"""Development aid (no GPU needed): runs bench.py's product arm -- run_ours(), the code path of the headline measurement --
with the CUDA calls faked and the CPU oracle standing in for the device handle, on a tiny workload.  The numbers it prints mean
nothing; what it catches is Python-level breakage of bench.py (names, keys, struct fields, the JSON line's shape) in a container
without a GPU.  The oracle is test infrastructure: this script lives in tools/ and is never part of a measurement.

    python tools/dryrun_bench.py            # prints the keys of the JSON line and checks the contract's required ones
"""
import io
import json
import os
import sys
import time
from contextlib import redirect_stdout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import torch  # noqa: E402

import alvrl_loader  # noqa: E402
import orc  # noqa: E402

pkg = alvrl_loader.load()


class FakeEvent:
    def __init__(self, enable_timing=False):
        self.t = 0.0

    def record(self):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


class StandIn(orc.Oracle):
    """the oracle behind the product binding's method names; calls the oracle does not mirror are no-ops"""

    def __init__(self, device=0, **params):
        super().__init__(**params)

    def _call(self, name, *args):
        if not self.api.has(name):
            return
        super()._call(name, *args)

    def stats(self):
        s = super().stats()
        s.msTransportKernelR = max(s.msTransportKernelR, 1.0)
        s.msTransportKernelRender = max(s.msTransportKernelRender, 1.0)
        return s


class FakeGroup:
    def __init__(self, g):
        self.g = g

    def comm_size(self):
        return 1

    def frame(self, want_image=False):
        g = self.g
        g.build_slices(); g.sample_slice_mapping(); g.build_R(); g.build_clusters()
        img = g.render()
        return img if want_image else None

    def close(self):
        pass


def main():
    real_device = torch.device
    torch.cuda.is_available = lambda: True
    torch.cuda.set_device = lambda *_: None
    torch.cuda.synchronize = lambda *_: None
    torch.cuda.Event = FakeEvent
    torch.device = lambda *a, **k: real_device("cpu")
    pkg.integrator = lambda device=0, **params: StandIn(device, **params)
    pkg.binding.Group.rank = staticmethod(lambda g, rank, world, uid: FakeGroup(g))
    sys.argv = ["bench.py", "--width", "48", "--height", "48", "--vrls", "300", "--steps", "2", "--warmup", "1",
                "--cpu-seconds", "1", "--parity-seconds", "1"] + sys.argv[1:]
    import bench
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.run_ours(bench.parse())
    line = json.loads(buf.getvalue().strip().splitlines()[-1])
    need = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
            "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"]
    missing = [k for k in need if k not in line]
    print("keys:", sorted(line))
    print("roofline:", {k: line["roofline"][k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic")})
    print("e2e:", line["e2e"]); print("cpu_baseline:", {k: line["cpu_baseline"][k] for k in ("value", "unit", "cores", "kind")})
    print("parity:", {k: (v if not isinstance(v, dict) else "...") for k, v in line.get("parity", {}).items()})
    assert not missing, missing
    assert line["config"].get("workload") and "model" not in line["config"]
    print("dry run ok")


if __name__ == "__main__":
    main()
