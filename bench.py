"""bench.py -- headline benchmark of the VRL hot path (BASELINE.json: VRL-segment contributions/s + 1024^2 frame time).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port) on the host cores

One "step" = one frame of the hot path on synthetic input (config C2 of BASELINE.json: Cornell box, homogeneous
isotropic medium, 1024x1024, 100k VRLs, volVolSamples = volSurfSamples = 4):
    Preprocessor::buildSlices -> sampleSliceMapping -> "Building R" -> buildClusters -> clustered render of every pixel
`value`  : integrateVRL evaluations (R entries + clustered render terms, the reference's own StatsCounter unit) per second,
           whole job, VRL set / BVH / medium resident in HBM when the timed region starts, framebuffer left on the device.
`e2e`    : the same metric through the C ABI with HOST buffers: every step uploads the VRL set (alvrl_set_vrls from host
           arrays) and reads the W x H x 3 float image back (alvrl_render into a host buffer).
Multi-GPU: slices are sharded over the ranks (VRLs, BVH, medium replicated) by the library's own group entry point
(alvrl_group_frame, csrc/group.cu); the only exchanges are an N-byte all-reduce of the zero / non-zero column flags and the
framebuffer reduce over NCCL.  Total work is fixed -> "scaling": "strong".  torch.distributed carries the NCCL unique id and
the max-over-ranks timing only.
`parity` : the benchmarked (fast) and the strict flavour against the strict CPU oracle on a bounded sample of this workload.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def flops_per_pair(nvv, nvs, hg):
    """algorithmic flops of one integrateVRL call, traversal excluded."""
    return 160 + (158 if hg else 136) * nvv + (102 if hg else 91) * nvs


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(self.rows)}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2")
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--vrls", type=int, default=None)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the bounded cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=1, help="steps of the end-to-end leg for the configs other than C2")
    ap.add_argument("--slice-range", type=int, nargs=2, default=None, metavar=("BEGIN", "END"),
                    help="1 GPU, configs other than C2: a frame covers only these slice ids (R rows, clusters and pixels of those "
                         "slices, full-size scene / grid / BVH / VRL set) -- the bounded form of a config whose whole frame takes minutes")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle parity sample (rank 0, N=1 only)")
    ap.add_argument("--no-strict", action="store_true", help="skip the extra frame in the strict math flavour")
    ap.add_argument("--parity-seconds", type=float, default=8.0, help="target CPU time of the parity sample")
    return ap.parse_args()


def workload(pkg, a):
    scene, vrls, params = pkg.scenes.make_config(a.config, width=a.width, height=a.height, n_vrls=a.vrls)
    cfg = pkg.scenes.CONFIGS[a.config]
    W, H = scene["camera"]["width"], scene["camera"]["height"]
    hg = cfg.get("hg") is not None
    med = scene["medium"]
    if med["type"] == "grid":
        what = f"Cornell box, heterogeneous grid-volume medium ({'x'.join(str(d) for d in med['density'].shape)} procedural density, method=simpson)"
    elif a.config == "C4":
        what = f"{len(scene['tris'])}-triangle icosphere occluder scene in homogeneous fog inside closed Cornell walls (BVH traversal path)"
    else:
        what = "Cornell box, homogeneous medium, " + (f"HG phase g={cfg['hg']}" if hg else "isotropic phase")
    desc = {"workload": f"{a.config}: {what}, {W}x{H}, {len(vrls[0])} VRLs, "
                        f"volVolSamples={params['volVolSamples']}, volSurfSamples={params['volSurfSamples']}",
            "slices": params.get("targetNumSlices", 100), "pixel_undersampling": 64,
            "l2_note": f"inputs larger than L2: R alone is rows x VRLs x 8 B ({W * H / 64 * len(vrls[0]) * 8 / 1e9:.1f} GB) and is rewritten every step"}
    if a.slice_range:
        desc["slice_range"] = list(a.slice_range)
        desc["workload"] += f", slices [{a.slice_range[0]}, {a.slice_range[1]}) of {desc['slices']} only"
    return scene, vrls, params, desc, hg


# ---------------------------------------------------------------------------------------------------------------------
def cpu_baseline(pkg, scene, vrls, params, hg, seconds, full_desc):
    """The reference's CPU algorithm (oracle port, -O3 + the reference's -funsafe-math-optimizations) on a bounded sample
    of the same workload with all host threads: the R rows of the first slices x a prefix of the VRL set, sized from a
    short calibration so that the timed part takes about `seconds`."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc  # the CPU oracle: allowed here (cpu_baseline / --impl reference legs only)
    cores = os.cpu_count() or 1
    start, end, power, pc = vrls

    def run(n_vrls, n_slices):
        o = orc.Oracle(fast=True, threads=cores, **params)
        o.set_scene(scene)
        o.set_vrls(start[:n_vrls], end[:n_vrls], power[:n_vrls], pc)
        o.build_slices()
        o.sample_slice_mapping()
        off, _ = o.rep_pixels()
        n_slices = min(n_slices, len(off) - 1)
        o.set_slice_range(0, n_slices)
        t0 = time.perf_counter()
        o.build_R()
        dt = time.perf_counter() - t0
        return int(off[n_slices]) * o.N, dt, int(off[n_slices]), o.N

    pairs, dt, _, _ = run(min(len(start), 2000), 1)                    # calibration (a second or so)
    rate = pairs / max(dt, 1e-3)
    rows_per_slice = max(1, pairs // min(len(start), 2000))
    want = seconds * rate
    n_vrls = int(min(len(start), max(2000, want / rows_per_slice)))
    n_slices = int(max(1, want / (rows_per_slice * n_vrls)))
    pairs, dt, rows, n_used = run(n_vrls, n_slices)
    return {"value": pairs / dt, "unit": "VRL-segment contributions/s", "cores": cores, "kind": "port",
            "sample": f"R rows of the first {n_slices} slice(s) = {rows} rows x the first {n_used} VRLs ({pairs:.3e} integrateVRL calls) "
                      f"of {full_desc}; oracle port built -O3 -march=x86-64-v3 -funsafe-math-optimizations, {cores} threads, {dt:.1f} s"}, dt


def parity_sample(pkg, g, scene, vrls, params, seconds):
    """The benchmarked kernels against the STRICT oracle (liborc.so: IEEE fp32, the parity definition -- not the
    -funsafe-math speed build cpu_baseline times) on a bounded sample of this very workload: the R rows of slice 0 x a
    prefix of the VRL set, same counter stream (an entry depends on (seed, row, vrl) only, so a VRL prefix reproduces the
    corresponding columns of the full R).  Entries whose shadow rays are oracle-flagged grazing ties (occlusion decision
    within 1e-5 of flipping) are counted separately; tests/test_c2_parity_gpu.py asserts the same quantities."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc  # the checker; runs after the timed region
    cores = os.cpu_count() or 1
    start, end, power, pc = vrls
    n = int(min(len(start), max(2000, seconds * 0.6e6 * min(cores, 32) / 8 / 160)))       # ~0.6 M pairs/s per 8 threads, ~160 rows
    o = orc.Oracle(fast=False, threads=cores, **params)
    o.set_scene(scene); o.set_vrls(start[:n], end[:n], power[:n], pc)
    o.set_graze_tolerance(1e-5)
    o.build_slices(); o.set_slice_range(0, 1); o.sample_slice_mapping()
    off, px = o.rep_pixels()
    rows = int(off[1])
    goff, gpx = g.rep_pixels()
    if not np.array_equal(gpx[:rows], px[:rows]):
        return {"error": "representative pixels of slice 0 differ from the oracle's"}
    t0 = time.perf_counter(); o.build_R(); dt = time.perf_counter() - t0
    Ro, graze = o.get_R(0, rows), o.R_graze(0, rows).astype(bool)
    out = {"rows": rows, "vrls": n, "oracle": "liborc.so (IEEE fp32, -ffp-contract=off)", "oracle_seconds": round(dt, 1), "graze_tol": 1e-5,
           "frac_graze_flagged": float(graze.mean())}
    for name, strict in (("fast", 0), ("strict", 1)):
        g._call("set_math_mode", pkg.binding.C.c_int(strict))
        g.set_slice_range(0, 1); g.sample_slice_mapping(); g.build_R()
        Rg = g.get_R(0, rows)[:, :n]
        mo, mg = Ro[..., 0], Rg[..., 0]
        floor = 1e-12 * np.abs(mo).max()
        em = np.abs(mg - mo) / (np.abs(mo) + floor)
        ev = np.abs(Rg[..., 1] - Ro[..., 1]) / (Ro[..., 1] + mo * mo + floor * floor)
        bad = (em > 1e-4) | (ev > 1e-4)
        out[name] = {"max_rel": float(em.max()), "max_rel_unflagged": float(em[~graze].max()), "p9999": float(np.quantile(em, 0.9999)),
                     "median": float(np.median(em)), "frac_gt_1e-4": float(bad.mean()),
                     "frac_gt_1e-4_unflagged": float((bad & ~graze).sum() / max(1, (~graze).sum()))}
    g._call("set_math_mode", pkg.binding.C.c_int(0))
    g.set_slice_range(0, 0xFFFFFFFF)
    return out


def run_reference(a):
    """--impl reference: the reference's own CPU implementation of the path.  The reference cannot be built in this image
    (Boost/Xerces-C/OpenEXR/SCons missing), so this is the oracle port, with all host threads, on bounded
    samples of the same workload.  Under torchrun only rank 0 works."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import alvrl_loader
    pkg = alvrl_loader.load()
    scene, vrls, params, desc, hg = workload(pkg, a)
    vals, times = [], []
    base = None
    budget = max(2.0, min(a.cpu_seconds, 150.0 / max(1, a.steps + a.warmup)))
    for i in range(a.warmup + a.steps):
        base, dt = cpu_baseline(pkg, scene, vrls, params, hg, budget, desc["workload"])
        if i >= a.warmup:
            vals.append(base["value"]); times.append(dt)
    v = float(np.mean(vals))
    base["value"] = v
    line = {"impl": "reference", "metric": "vrl_segment_contributions_per_s", "value": v, "unit": "VRL-segment contributions/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": float(np.mean(times)) * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            # the same config object as the product arm's line (the driver compares them key by key)
            "config": dict(desc, parallelism=f"slice-sharded x{int(os.environ.get('WORLD_SIZE', a.gpus))}"), "cpu_baseline": base,
            "e2e": {"value": v, "unit": "VRL-segment contributions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist
    import alvrl_loader
    pkg = alvrl_loader.load()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    multi = world > 1
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    saved_stdout = None
    if multi:
        # the caller's NCCL_DEBUG stays as it is (the driver reads the rank count from NCCL's INFO lines).  NCCL logs to
        # stdout: while the job runs, file descriptor 1 points at stderr, and it is put back for the one JSON line
        # (NCCL_DEBUG_FILE=/dev/stderr would truncate a redirected log file)
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)

    scene, vrls, params, desc, hg = workload(pkg, a)
    g = pkg.integrator(local, **params)
    g.set_scene(scene)
    g.set_vrls(*vrls)
    W, H, N = g.W, g.H, g.N
    host_img = None
    # the library's multi-GPU entry point (alvrl_group_*, csrc/group.cu): this rank's handle + an NCCL communicator built
    # from a unique id that rank 0 creates and torch.distributed carries to the other ranks (plumbing only)
    uid = [pkg.binding.Group.unique_id(pkg.api()) if (multi and rank == 0) else None]
    if multi:
        dist.broadcast_object_list(uid, src=0)
    group = pkg.binding.Group.rank(g, rank, world, uid[0])
    comm_ranks = group.comm_size()

    def barrier():
        torch.cuda.synchronize()
        if multi:
            dist.barrier()
        torch.cuda.synchronize()

    def frame(e2e):
        """one step = alvrl_group_frame: slices -> slice range of this rank -> slice mapping -> R -> column-flag all-reduce ->
        clusters -> render -> framebuffer reduce; returns (#integrateVRL evaluations of this rank, kernel times, ...)"""
        nonlocal host_img
        s0 = g.stats()
        if e2e:
            g.set_vrls(*vrls)                                    # host -> device: the step's input (VRL set)
        if a.slice_range:                                        # bounded form: the same calls alvrl_group_frame makes, on a sub-range
            g.build_slices(); g.set_slice_range(*a.slice_range); g.sample_slice_mapping(); g.build_R(); g.set_column_nonzero(None)
            g.build_clusters()
            img = g.render()                                     # the image comes back to the host in both legs (W x H x 3 floats)
            if not e2e:
                img = None
        else:
            img = group.frame(want_image=e2e and rank == 0)      # e2e: device -> host read of the step's result on rank 0
        if img is not None:
            host_img = img
        s1 = g.stats()
        pairs = (s1.pairsPreprocess - s0.pairsPreprocess) + (s1.pairsRender - s0.pairsRender)
        return pairs, s1.msTransportKernelR, s1.msTransportKernelRender, s1.pairsPreprocess - s0.pairsPreprocess, \
            s1.kernelLaunches - s0.kernelLaunches, s1

    def timed(e2e, steps):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        tot_pairs = tot_launch = 0
        kr, kp, kpairs = [], [], []
        last = None
        for _ in range(steps):
            p, msr, msp, rp, nl, last = frame(e2e)
            tot_pairs += p; tot_launch += nl
            kr.append(msr); kp.append(msp); kpairs.append(rp)
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        t = torch.tensor([ms, float(tot_pairs), float(tot_launch)], dtype=torch.float64, device=dev)
        if multi:
            tm = t.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ts = t.clone(); dist.all_reduce(ts, op=dist.ReduceOp.SUM)
            ms, tot_pairs, tot_launch = float(tm[0]), float(ts[1]), float(ts[2])
        return ms, tot_pairs, tot_launch, kr, kp, kpairs, last

    # W >= 3 warm-up frames for the headline configuration; the larger configs (minutes per frame) may be run with fewer
    warm = max(3, a.warmup) if a.config == "C2" else a.warmup
    for _ in range(warm):
        frame(False)
    clocks = ClockSampler(local)
    clocks.start()
    ms, pairs, launches, kr, kp, kpairs, st = timed(False, a.steps)
    clk = clocks.finish()
    e_steps = a.steps if a.config == "C2" else max(0, min(a.steps, a.e2e_steps))
    if e_steps:
        ms_e, pairs_e, _, _, _, _, st_e = timed(True, e_steps)
    else:                                                    # (large configs only) no end-to-end leg: reported as null
        ms_e, pairs_e, st_e, e_steps = float("nan"), float("nan"), st, 1

    # roofline of the dominant kernel (k_build_R_fast): algorithmic flops per launch / CUDA-event duration of that launch,
    # measured inside the library on the launching stream (alvrl_stats.msTransportKernelR)
    F = flops_per_pair(params["volVolSamples"], params["volSurfSamples"], hg)
    k_ms = float(np.mean(kr))
    k_pairs = float(np.mean(kpairs))
    peak = pkg.binding.C.c_float()
    pkg.api().lib.alvrl_measure_fp32_peak(pkg.binding.C.c_int(local), pkg.binding.C.byref(peak))
    achieved = k_pairs * F / (k_ms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    vis_mode = {0: "BVH traversal", 1: "flat leaf sweep", 2: "occluder set"}.get(int(st.visMode), "?")
    # what SURVEY 8(d) names as the bound of the transport kernel per config; the FP32 fraction is reported for every config so
    # that the lines are comparable, with the note saying what actually limits the kernel there
    bound_note = {"C3": "memory system (SURVEY 8d): the Simpson marches stream ~172 KB of the 512 MiB grid per contribution, 3.0 TB/s of DRAM at "
                        "16 % L2 hit rate (ncu, profiles/r2_c3_c4_transport.txt); the FP32 fraction is not the limiter here",
                  "C4": "BVH traversal (SURVEY 8d): latency of dependent node fetches, 20.8 node visits per shadow ray in the 4-wide tree, L2 hit "
                        "rate 87 % (ncu, profiles/r2_c3_c4_transport.txt); see shadow_rays_per_s_R_kernel; the FP32 fraction is not the limiter here"}.get(a.config)
    roof = {"bound": "fp32", "kernel": f"k_build_R_fast<{'grid' if a.config == 'C3' else 'homogeneous'} medium, {vis_mode}>", "achieved": achieved,
            "peak": float(peak.value), "unit": "TFLOP/s",
            "frac": achieved / float(peak.value) if peak.value else None,
            # dram__bytes_read.sum + dram__bytes_write.sum of this very launch (C2: 15 759 rows x 100 000 VRLs) from one
            # `ncu --set full` capture of the round-2 kernel, profiles/r2_transport_v11.txt: 0.2349 GB read + 12.6687 GB written
            "traffic": 12903596944 if (N == 100000 and W == 1024 and H == 1024 and world == 1 and a.config == "C2" and not a.slice_range) else None,
            "peak_source": "FP32 FFMA microbenchmark measured live on this device (alvrl_measure_fp32_peak); north_star names the "
                           "non-tensor FP32 roofline for this kernel (no dense contraction, tensor cores unused); nominal 148 SM x 128 x 2 x 1.965 GHz = 74.5",
            "flops_per_contribution": F, "contributions_per_launch": k_pairs, "launch_ms": k_ms,
            "contributions_per_s_kernel": k_pairs / (k_ms * 1e-3),
            "hbm": {"algorithmic_bytes_per_launch": k_pairs * 8, "achieved_gbs": k_pairs * 8 / (k_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}}
    if bound_note:
        roof["bound_note"] = bound_note

    if rank == 0:
        line = {"metric": "vrl_segment_contributions_per_s", "value": pairs / (ms * 1e-3), "unit": "VRL-segment contributions/s",
                "n_gpus": world, "steps": a.steps, "warmup": warm, "ms_per_step": ms / a.steps, "frame_time_ms": ms / a.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(desc, parallelism=f"slice-sharded x{world}"),
                "e2e": {"value": (pairs_e / (ms_e * 1e-3)) if ms_e == ms_e else None, "unit": "VRL-segment contributions/s", "ms_per_step": (ms_e / e_steps) if ms_e == ms_e else None,
                        "h2d_bytes_per_step": int(N * 9 * 4), "d2h_bytes_per_step": int(W * H * 3 * 4)},
                "gpu_launches": int(launches), "clocks": clk, "roofline": roof,
                "phases_ms": {"slices": st.msSlices, "slice_mapping": st.msSliceMapping, "build_R": st.msBuildR, "clusters": st.msClusters,
                              "render_kernel": st.msTransportKernelRender},
                "e2e_phases_ms": {"slices": st_e.msSlices, "build_R": st_e.msBuildR, "clusters": st_e.msClusters, "render_total": st_e.msRender},
                "contributions_per_step": pairs / a.steps, "rows": st.numRows, "vrls": st.numVrls, "slices": st.numSlices,
                "shadow_rays_per_s_R_kernel": k_pairs * (params["volVolSamples"] + params["volSurfSamples"]) / (k_ms * 1e-3),
                "visibility_mode": {0: "stackless BVH traversal per lane", 1: "flat leaf sweep", 2: "compiled occluder set + pair-level culling"}.get(int(st.visMode), "?"),
                "bvh_nodes": int(st.bvhNodes),
                "scene_build_ms": float(st.msSceneBuild),     # host BVH build + triangle records + upload: outside the frame, reported separately
                "comm": {"library": "NCCL via alvrl_group_* (csrc/group.cu)" if multi else None, "nranks": comm_ranks},
                "frame_definition": "buildSlices + sampleSliceMapping + Building R + buildClusters (per-slice refinement; the global/"
                                    "fallback lists are only computed when a slice cannot be refined, as no slice of this workload "
                                    "does) + clustered render of every pixel; k_primary (camera segments) is cached across steps "
                                    "because camera and mesh do not change (65 us)"}
        if world == 1 and not a.no_strict:
            # the parity flavour of the same kernels (reference operation order, no FMA contraction, exp through double): one frame
            g._call("set_math_mode", pkg.binding.C.c_int(1))
            frame(False)
            ms_s, pairs_s, _, kr_s, _, kpairs_s, st_s = timed(False, 1)
            g._call("set_math_mode", pkg.binding.C.c_int(0))
            line["strict_flavour"] = {"value": pairs_s / (ms_s * 1e-3), "frame_time_ms": ms_s, "build_R_kernel_ms": float(kr_s[0]),
                                      "contributions_per_s_kernel": float(kpairs_s[0]) / (float(kr_s[0]) * 1e-3),
                                      "render_kernel_ms": st_s.msTransportKernelRender}
        if world == 1 and not a.no_parity:
            try:
                line["parity"] = parity_sample(pkg, g, scene, vrls, params, a.parity_seconds)
            except Exception as e:                                  # the checker must never take the bench line down
                line["parity"] = {"error": repr(e)}
        if not a.no_cpu_baseline and world == 1:
            line["cpu_baseline"], _ = cpu_baseline(pkg, scene, vrls, params, hg, a.cpu_seconds, desc["workload"])
        sys.stdout.flush()
        if saved_stdout is not None:
            os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        if saved_stdout is not None:
            os.dup2(2, 1)
    group.close()
    if multi:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
