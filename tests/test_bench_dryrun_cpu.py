"""CPU suite: bench.py stays runnable.  tools/dryrun_bench.py drives bench.py's product arm (run_ours) on a tiny workload with the
CUDA calls faked and the CPU oracle standing in for the device handle -- the numbers mean nothing, the point is that the headline
path has no Python-level breakage and that its JSON line carries every key of the measurement contract.  The reference arm
(--impl reference) runs for real on the CPU."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_product_arm_dry_run():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "dryrun_bench.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "dry run ok" in out.stdout


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-seconds", "2",
                          "--width", "128", "--height", "128", "--vrls", "2000"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "parallelism" in line["config"]
    # under torchrun only rank 0 works: the other ranks exit 0 without output
    other = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"], capture_output=True,
                           text=True, timeout=120, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert other.returncode == 0 and other.stdout.strip() == ""
