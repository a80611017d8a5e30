#!/usr/bin/env python3
"""bench.py — the hot path's headline metric on B200.

Metric (BASELINE.json): Mpix·depth-candidates/s = computeCost evaluations per second.
Workload `bf128_l0` (default): 16-camera FTHETA ring rig, 2048x2048, 128 candidates — the level-0
brute-force sphere sweep the north-star quotes its roofline target on.  One step = one frame:
for each of the 16 destination cameras, reprojectColors/precomputeProjections (K2-K4) followed by
the fused sweep + cost + WTA kernel (K6).  At N GPUs every rank processes its own frame per step
(frames shard with no data-path collective: weak scaling).

`value`  : inputs already resident in HBM when the timed region starts.
`e2e`    : the same step through the C-ABI with HOST buffers: pinned-host -> device copy of the 16
           colour images and device -> host read-back of the 16 disparity maps inside the timed region.
`roofline`: the dominant kernel (sweepKernel) timed live with CUDA events on its launching stream;
           algorithmic bytes B_stream = 20 B x (pixel,candidate,source) triples + 30 B x pixels
           (SURVEY.md §8(d)), both counted exactly by the kernel's own work counters.
`cpu_baseline`: the oracle (a port of the reference CPU path, the reference itself cannot be built
           here) on this box's host cores, on a bounded sample: 1 destination camera x 8 evenly
           spaced candidates x the full 2048x2048 frame.

--impl reference times that same CPU port as the reference arm (rank 0 only).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))


def _baseline_metric():
    """The metric name exactly as BASELINE.json spells it (the driver matches the bench line against that file)."""
    try:
        return json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]
    except Exception:
        return "Mpix·depth-candidates/s at 16-cam 2K×2K, 1/2/4/8 GPU vs ISPC CPU"


METRIC = _baseline_metric()

sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (num_cams, width, height, num_depths, kind)
    "bf128_l0": (16, 2048, 2048, 128, "FTHETA"),
    "bf32_cfg1": (4, 512, 512, 32, "RECTILINEAR"),  # BASELINE.json configs[0] (parity case; small)
    # BASELINE.json configs[1] as the reference runs it: 5-level coarse-to-fine frame (see CoarseToFine)
    "c2f5": (16, 2048, 2048, 128, "FTHETA"),
    # BASELINE.json configs[3]: ONE 24-camera 4096^2 frame, 256 candidates + bilateral, destination cameras dealt to the GPUs
    # (strong scaling); two levels so that the mismatch stage and its all-gather of disparity planes are on the path
    "cfg4": (24, 4096, 4096, 256, "FTHETA"),
    # BASELINE.json configs[4]: 16 cameras x 30 frames, 5-level coarse-to-fine per frame with the temporal filter after every
    # level (scripts/render/pipeline.py:364-408), contiguous frame blocks per GPU, halo frames exchanged over NVLink
    "cfg5": (16, 2048, 2048, 128, "FTHETA"),
}
CFG5_FRAMES = 30
C2F_LEVELS = 5
C2F_CPU_LEVEL = 3  # the level the CPU arm times (256 x 256 at 2048 full size): seconds, not minutes, of reference code
MIN_DEPTH, MAX_DEPTH = 0.5, 1e4


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_inputs(workload, device, rank=0):
    from facebook360_dep_b200 import synth
    S, W, H, D, kind = WORKLOADS[workload]
    rig = synth.ring_rig(S, W, H, kind=kind, hfov_deg=120.0 if kind == "RECTILINEAR" else None)
    t0 = time.time()
    scene = synth.Scene(seed=42 + rank)  # one frame per rank: same rig, different scene seed
    colors, _ = synth.render_rig(rig, W, H, scene=scene, device=device)
    log("[bench] rendered %d x %dx%d synthetic frames in %.1fs on %s" % (S, W, H, time.time() - t0, device))
    return rig, colors


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile(prefix="clocks_", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=f, stderr=subprocess.DEVNULL)
        except Exception as e:  # nvidia-smi missing
            log("[bench] clock sampling unavailable:", e)
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()  # the exact PID we started
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no_samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_capture(workload):
    """Per-launch counters of sweepKernel from the committed ncu capture (profiles/sweep_traffic.json), if one exists
    for this workload: dram bytes and executed warp instructions."""
    p = os.path.join(ROOT, "profiles", "sweep_traffic.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            if j.get("workload") == workload:
                return j
        except Exception:
            pass
    return {}


def ncu_traffic_per_launch(workload):
    return ncu_capture(workload).get("dram_bytes_per_launch")


def issue_roof(workload, ms_per_launch, sm_mhz, sms=148):
    """The roof that actually binds the sweep: warp-instruction issue.  Instructions per launch come from the
    committed ncu capture of the same workload (a static property of kernel + inputs), time from this run's CUDA
    events, peak = SMs x 4 schedulers x 1 instruction/clk at the SM clock sampled during this run."""
    n = ncu_capture(workload).get("warp_instructions_per_launch")
    if not n or not ms_per_launch or not sm_mhz:
        return None
    achieved = n / (ms_per_launch / 1e3) / 1e9
    peak = sms * 4 * sm_mhz * 1e6 / 1e9
    return {"warp_instructions_per_launch": n, "achieved": achieved, "peak": peak, "unit": "G warp-instr/s",
            "frac": achieved / peak}


def host_threads():
    """Threads this process may really use: the scheduler affinity mask, capped by the cgroup CPU quota (os.cpu_count()
    ignores both), and the number of distinct physical cores among them."""
    try:
        cpus = sorted(os.sched_getaffinity(0))
    except AttributeError:
        cpus = list(range(os.cpu_count() or 1))
    n = len(cpus)
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(float(txt[0]) / float(txt[1]))))
            else:
                q = int(txt[0])
                per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                if q > 0:
                    n = min(n, max(1, q // per))
            break
        except Exception:
            continue
    phys = set()
    try:
        cur = {}
        for line in open("/proc/cpuinfo"):
            if ":" in line:
                k, v = [x.strip() for x in line.split(":", 1)]
                cur[k] = v
            elif not line.strip():
                if "processor" in cur and int(cur["processor"]) in cpus:
                    phys.add((cur.get("physical id", "0"), cur.get("core id", cur["processor"])))
                cur = {}
    except Exception:
        pass
    return n, (len(phys) or None)


def candidate_table(D):
    """probeDisparity (ImageUtil.cpp:100-107) as Derp.cpp:279-285 calls it: fp32 ends, fp64 mix, fp32 result."""
    dmin = np.float32(1.0) / np.float32(MAX_DEPTH)
    dmax = np.float32(1.0) / np.float32(MIN_DEPTH)
    f = np.arange(D, dtype=np.float64) / float(D - 1)
    return (f * float(dmin) + (1.0 - f) * float(dmax)).astype(np.float32)


class CpuArm:
    """The reference's CPU implementation of the hot path on this box's host cores, on bounded samples.

    kind "reference": oracle/_ref/libderp_ref.so = the reference's own Derp.cpp / DerpUtil.cpp / Camera.cpp / CvUtil.h ...
    compiled against stand-in headers (oracle/ref_bridge.cpp); one sample = `threads` candidate slices of destination 0
    over a band of rows, ONE thread per candidate slice — the reference's own task structure (Derp.cpp:288-304 spawns
    one ThreadPool task per candidate and joins a batch of `threads` tasks at a time).
    kind "port": the same through the oracle restatement when oracle/_ref is not there (row-parallel, brute force)."""

    def __init__(self, workload, rig, colors):
        from facebook360_dep_b200 import capi
        from tests import oracle_libs  # the checker: cpu_baseline / --impl reference legs only
        self.S, self.W, self.H, self.D, self.kind_cam = WORKLOADS[workload]
        self.threads, self.physical = host_threads()
        self.lib = oracle_libs.load_ref()
        self.kind = "reference" if self.lib is not None else "port"
        if self.lib is None:
            self.lib = oracle_libs.load_oracle()
        self.lib.set_threads(self.threads)
        self.table = candidate_table(self.D)
        t0 = time.perf_counter()
        self.ctx = capi.Context(self.lib, capi.rig_descs(rig), dst_to_src=[0])
        self.ctx.level_begin(self.W, self.H)
        self.ctx.set_colors(colors)
        self.ctx.reproject(0)
        self.fov = self.ctx.get_fov_mask(0)
        log("[bench] cpu arm (%s): %d threads (%s physical cores), tables of destination 0 built in %.1fs" % (
            self.kind, self.threads, self.physical, time.perf_counter() - t0))
        n = min(self.threads, self.D)
        self.slices = np.unique(np.round(np.linspace(0, self.D - 1, n)).astype(int))
        self.rows = None
        self.vbar = None

    def _band(self, rows):
        y0 = max(1, (self.H - rows) // 2)
        return y0, min(self.H - 1, y0 + rows)

    def _run(self, y0, y1, want_costs=False):
        import ctypes as C
        n = len(self.slices)
        disp = np.ascontiguousarray(self.table[self.slices])
        active = int(self.fov[y0:y1, 1:self.W - 1].astype(bool).sum())
        if self.kind == "reference":
            f = self.lib.lib.derp_ref_cost_slices
            f.restype = C.c_int
            f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
            costs = np.empty((n, y1 - y0, self.W), np.float32) if want_costs else None
            t0 = time.perf_counter()
            self.lib.check(f(self.ctx.h, 0, disp.ctypes.data, n, y0, y1, None if costs is None else costs.ctypes.data, None))
            dt = time.perf_counter() - t0
            return active * n, dt, costs
        # port: one eval_cost per slice on a full constant-disparity map (row-parallel inside the oracle)
        costs = np.empty((n, y1 - y0, self.W), np.float32) if want_costs else None
        t0 = time.perf_counter()
        for k, d in enumerate(disp):
            c, _ = self.ctx.eval_cost(0, np.full((self.H, self.W), d, np.float32))
            if costs is not None:
                costs[k] = np.where(self.fov[y0:y1].astype(bool), c[y0:y1], np.nan)
        dt = time.perf_counter() - t0
        full_active = int(self.fov[1:self.H - 1, 1:self.W - 1].astype(bool).sum())
        return full_active * n, dt, costs

    def calibrate(self, target_s=2.5):
        """Choose the band height so that one step is about target_s of CPU work (outside every timed region)."""
        rows = max(8, self.H // 64)
        y0, y1 = self._band(rows)
        evals, dt, _ = self._run(y0, y1)
        per_row = dt / max(1, y1 - y0)
        self.rows = int(min(self.H - 2, max(8, target_s / max(per_row, 1e-9))))
        if self.kind == "port":
            self.rows = self.H - 2
        log("[bench] cpu arm: calibration %d rows in %.2fs -> %d rows per step" % (y1 - y0, dt, self.rows))

    def sample(self, steps, warmup, want_costs=False):
        if self.rows is None:
            self.calibrate()
        y0, y1 = self._band(self.rows)
        rates, times, costs = [], [], None
        for i in range(warmup + steps):
            last = i == warmup + steps - 1
            evals, dt, c = self._run(y0, y1, want_costs and last)
            if c is not None:
                costs = c
            if i >= warmup:
                rates.append(evals / dt)
                times.append(dt)
            log("[bench] cpu arm step %d%s: %.2fs, %.3f Mpix·cand/s" % (i, " (warm-up)" if i < warmup else "", dt, evals / dt / 1e6))
        desc = ("destination 0 of the %d-camera rig, %d candidate slices (one host thread each, the reference's task "
                "structure) x rows %d..%d of the %dx%d frame" % (self.S, len(self.slices), y0, y1 - 1, self.W, self.H))
        return {"rates": rates, "times": times, "sample": desc, "band": (y0, y1), "costs": costs}

    def baseline_object(self, res):
        return {"value": statistics.mean(res["rates"]) / 1e6, "unit": "Mpix·cand/s", "cores": self.threads,
                "physical_cores": self.physical, "kind": self.kind, "sample": res["sample"],
                "steps": len(res["rates"]), "seconds_per_step": statistics.mean(res["times"])}

    def close(self):
        self.ctx.close()


def parity_vs_cpu(arm, res, gpu_ctx):
    """The CPU sample's cost maps against the CUDA library's computeCost of the SAME candidates on the SAME frame
    (derp_eval_cost on constant-disparity maps, destination 0): mismatching cost fraction and winner-index flips."""
    y0, y1 = res["band"]
    cpu = res["costs"]
    if cpu is None:
        return None
    gpu_ctx.reproject(0)
    n = len(arm.slices)
    gpu = np.empty_like(cpu)
    for k in range(n):
        c, _ = gpu_ctx.eval_cost(0, np.full((arm.H, arm.W), arm.table[arm.slices[k]], np.float32))
        gpu[k] = c[y0:y1]
    valid = ~np.isnan(cpu)
    same = (cpu.view(np.uint32) == gpu.view(np.uint32)) & valid
    pixels = int(valid.sum())
    # winner over the sampled slices, first strict minimum in index order (Derp.cpp:323-333); NaN never wins
    def wta(v):
        w = np.where(np.isnan(v), np.float32(np.inf), v)
        return np.argmin(w, axis=0)
    col = valid.any(axis=0)
    flips = int((wta(cpu) != wta(np.where(valid, gpu, np.nan)))[col].sum())
    return {"against": arm.kind, "slices": n, "rows": [int(y0), int(y1)], "pixel_candidates": pixels,
            "cost_bit_mismatches": int(pixels - same.sum()), "cost_mismatch_frac": float(1.0 - same.sum() / max(1, pixels)),
            "index_mismatches": flips, "pixels": int(col.sum())}


class CoarseToFine:
    """BASELINE.json configs[1] as the reference runs it (DerpCLI defaults): brute force with D candidates at the coarsest
    of 5 levels, then random proposals (2) + ping-pong (1) + joint bilateral + median on every finer level.

    The pyramid is what scripts/render/resize.py:79 feeds DerpCLI: every level cv::resize(INTER_AREA) of the FULL-SIZE
    image — built on the device by derp_downscale_area.  Levels are handed over in HBM (derp_level_keep /
    derp_upsample_from_kept) instead of the reference's PFM round trip (DerpCLI.cpp:287-288).
      frame(resident=True)  : level images already in HBM; times level_begin .. process_level of the 5 levels
      frame(resident=False) : end to end — pinned host full-size images -> device, pyramid build, the 5 levels, the 16
                              level-0 disparity planes back to pinned host memory"""

    def __init__(self, cuda, ctx, pin_colors, S, W, H, D, stream, device, levels=5):
        import torch
        self.torch, self.cuda, self.ctx, self.pin = torch, cuda, ctx, pin_colors
        self.S, self.W, self.H, self.D, self.stream, self.dev, self.levels = S, W, H, D, stream, device, levels
        self.full = [torch.empty((H, W, 3), dtype=torch.uint16, device=device) for _ in range(S)]
        self.lvl = [[torch.empty((H >> k, W >> k, 3), dtype=torch.uint16, device=device) for _ in range(S)]
                    for k in range(1, levels)]
        self.out = [torch.empty((H, W), dtype=torch.float32).pin_memory() for _ in range(S)]
        self.evals = self.hits = 0
        self.level_out = {}
        self.level_evals = {}

    def upload(self):
        for s in range(self.S):
            self.full[s].copy_(self.pin[s], non_blocking=True)

    def build_pyramid(self):
        for k in range(1, self.levels):
            w, h = self.W >> k, self.H >> k
            for s in range(self.S):
                self.cuda.check(self.cuda.lib.derp_downscale_area(self.dev.index or 0, self.full[s].data_ptr(), self.W, self.H,
                                                                 self.lvl[k - 1][s].data_ptr(), w, h))

    def levels_pass(self, keep_level=None):
        ctx = self.ctx
        self.evals = self.hits = 0
        for level in range(self.levels - 1, -1, -1):
            w, h = self.W >> level, self.H >> level
            ctx.level_begin(w, h, level=level, num_levels=self.levels, full_width=self.W, full_height=self.H)
            imgs = self.full if level == 0 else self.lvl[level - 1]
            ctx.set_colors_ptr([t.data_ptr() for t in imgs])
            if level < self.levels - 1:
                for d in range(self.S):
                    ctx.upsample_from_kept(d)
            ctx.process_level(num_depths=self.D, min_depth_m=MIN_DEPTH, max_depth_m=MAX_DEPTH, partial_coverage=True)
            e, h_ = ctx.get_counters()
            self.evals += e
            self.hits += h_
            self.level_evals[level] = e
            if level > 0:
                ctx.level_keep()
            if keep_level is not None and level in keep_level:
                self.level_out[level] = [ctx.get_disparity(d, want_cost=False) for d in range(self.S)]

    def download(self):
        for d in range(self.S):
            self.cuda.check(self.cuda.lib.derp_get_disparity(self.ctx.h, d, self.out[d].numpy().ctypes.data, None, None))

    def frame(self, resident, keep_level=None):
        if not resident:
            self.upload()
            self.build_pyramid()
        self.levels_pass(keep_level)
        if not resident:
            self.download()


def static_config(workload):
    """The workload description both arms print verbatim (the driver compares the two `config` objects)."""
    S, W, H, D, kind = WORKLOADS[workload]
    return {"workload": workload, "cameras": S, "width": W, "height": H, "candidates": D, "camera_model": kind,
            "min_depth_m": MIN_DEPTH, "max_depth_m": MAX_DEPTH, "frames_per_step_per_gpu": 1,
            "l2": "inputs larger than L2 (per destination %.0f MB of pair tables vs 126 MB L2)" % (
                (S - 1) * W * H * 24 / 1e6)}


def run_reference(args):
    """Reference arm: the reference's own CPU code of the path (oracle/_ref, see CpuArm) on this box's host cores.
    Each step is a bounded sample of the workload (a band of rows x `threads` candidate slices of destination 0);
    --warmup / --steps are honoured like in the GPU arm (at least one warm-up step)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # input generation is not part of the measurement: render on the GPU when there is one (torchrun pins
    # OMP_NUM_THREADS=1, which makes the CPU renderer take minutes), else on all host threads
    import torch
    if torch.cuda.is_available():
        gen_dev = "cuda:%d" % int(os.environ.get("LOCAL_RANK", "0"))
    else:
        torch.set_num_threads(os.cpu_count() or 1)
        gen_dev = "cpu"
    rig, colors = make_inputs(args.workload, gen_dev)
    if args.workload == "c2f5":
        return run_reference_c2f(args, rig, [np.ascontiguousarray(c) for c in colors])
    arm = CpuArm(args.workload, rig, [np.ascontiguousarray(c) for c in colors])
    warm = max(1, args.warmup)
    arm.calibrate(target_s=max(0.5, min(2.5, 150.0 / (args.steps + warm))))  # the whole run stays within a few minutes
    res = arm.sample(args.steps, warm)
    base = arm.baseline_object(res)
    value = base["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpix·cand/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": warm,
        "ms_per_step": 1e3 * statistics.mean(res["times"]), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 cost, f64 projection, u16 texels", "data": "synthetic",
        "config": static_config(args.workload),
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": "Mpix·cand/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    arm.close()
    print(json.dumps(line), flush=True)
    return 0


def c2f_cpu_level(lib, rig, level_colors, coarse_disps, threads, steps, warmup):
    """The reference's processLevel (Derp.cpp:1005-1034) of ONE fine level of the coarse-to-fine frame on the host cores:
    level C2F_CPU_LEVEL of C2F_LEVELS, all cameras, started from the given coarser-level disparities.  Returns
    (seconds per step list, disparities of the level)."""
    from facebook360_dep_b200 import capi
    S, W, H, D, kind = WORKLOADS["c2f5"]
    w, h = W >> C2F_CPU_LEVEL, H >> C2F_CPU_LEVEL
    lib.set_threads(threads)
    ctx = capi.Context(lib, capi.rig_descs(rig))
    times, out = [], None
    for i in range(warmup + steps):
        ctx.level_begin(w, h, level=C2F_CPU_LEVEL, num_levels=C2F_LEVELS, full_width=W, full_height=H)
        ctx.set_colors(level_colors)
        for d in range(S):
            ctx.upsample_from(d, coarse_disps[d])
        t0 = time.perf_counter()
        ctx.process_level(num_depths=150, min_depth_m=MIN_DEPTH, max_depth_m=MAX_DEPTH, partial_coverage=True)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        log("[bench] cpu arm (c2f level %d) step %d%s: %.2fs" % (C2F_CPU_LEVEL, i, " (warm-up)" if i < warmup else "", dt))
        out = [ctx.get_disparity(d, want_cost=False) for d in range(S)]
    evals = ctx.get_counters()[0]
    ctx.close()
    return times, out, evals


def run_c2f(args):
    """Workload c2f5: one 5-level coarse-to-fine frame of the 16-camera 2048^2 rig per step and per GPU."""
    import torch
    import torch.distributed as dist
    from facebook360_dep_b200 import capi, shard
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    S, W, H, D, kind = WORKLOADS["c2f5"]
    rig, colors = make_inputs("c2f5", dev, rank)
    pin = [torch.from_numpy(np.ascontiguousarray(c)).pin_memory() for c in colors]
    cuda = capi.load_cuda()
    ctx = capi.Context(cuda, capi.rig_descs(rig), device=local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    c2f = CoarseToFine(cuda, ctx, pin, S, W, H, D, stream, dev, C2F_LEVELS)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    c2f.upload()
    c2f.build_pyramid()
    for _ in range(max(1, args.warmup)):
        c2f.frame(True)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ctx.profile(True)
    l0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        c2f.frame(True)
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - l0
    pp_ms, pp_n, pp_evals, pp_hits = ctx.get_profile_ping_pong()
    ctx.profile(False)
    clocks = sampler.stop() if rank == 0 else None
    evals_step, hits_step = c2f.evals, c2f.hits
    # end to end: host images in, host disparities out
    c2f.frame(False)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        c2f.frame(False)
    e1.record(stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    ms, evals_all = shard.reduce_step(ms, evals_step, dev)
    e2e_max, _ = shard.reduce_step(e2e_ms, 0.0, dev)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    peak, peak_src = measured_peak_gbs()
    value = evals_all * args.steps / (ms / 1e3) / 1e6
    pp_launch_ms = pp_ms / max(1, pp_n)
    pp_bytes = (20.0 * pp_hits + 30.0 * (pp_evals / 9.0)) / max(1, pp_n)
    achieved = pp_bytes / (pp_launch_ms / 1e3) / 1e9 if pp_n else None
    line = {
        "metric": METRIC, "value": value, "unit": "Mpix·cand/s", "n_gpus": world, "steps": args.steps, "warmup": max(1, args.warmup),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 cost, f64 projection, u16 texels", "data": "synthetic", "config": static_config("c2f5"),
        "work": {"cost_evaluations_per_frame": evals_step, "triples_per_frame": hits_step,
                 "per_level": {str(k): int(v) for k, v in sorted(c2f.level_evals.items())}},
        "clocks": clocks,
        "e2e": {"value": evals_all * args.steps / (e2e_max / 1e3) / 1e6, "unit": "Mpix·cand/s",
                "h2d_bytes_per_step": S * W * H * 6, "d2h_bytes_per_step": S * W * H * 4, "ms_per_step": e2e_max / args.steps,
                "includes": "pinned host -> device of the 16 full-size images, INTER_AREA pyramid on the device, 5 levels "
                            "with in-HBM hand-off, 16 level-0 disparity planes back to pinned host memory"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "pingPongKernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": pp_bytes, "ms_per_launch": pp_launch_ms, "launches_timed": int(pp_n),
                     "kernel_share_of_step": (pp_ms / ms) if ms else None,
                     "triples_per_s": pp_hits / (pp_ms / 1e3) if pp_ms else None,
                     "note": "B_stream = 20 B x (pixel,cand,source) triples + 30 B x pixels, pixels taken as evaluations / 9 "
                             "(<= 9 neighbour candidates per active pixel); the kernel gathers scattered 4x4 texel blocks "
                             "and is L1 / latency bound, not HBM bound (profiles/README.md)"},
    }
    if world == 1 and not args.no_cpu_baseline:
        from tests import oracle_libs
        lib = oracle_libs.load_ref()
        kind_cpu = "reference" if lib is not None else "port"
        if lib is None:
            lib = oracle_libs.load_oracle()
        threads, physical = host_threads()
        c2f.frame(True, keep_level={C2F_CPU_LEVEL + 1, C2F_CPU_LEVEL})
        lvl_colors = [t.cpu().numpy() for t in c2f.lvl[C2F_CPU_LEVEL - 1]]
        times, cpu_disp, _ = c2f_cpu_level(lib, rig, lvl_colors, c2f.level_out[C2F_CPU_LEVEL + 1], threads, steps=3, warmup=1)
        evals_lvl = c2f.level_evals[C2F_CPU_LEVEL]
        line["cpu_baseline"] = {"value": evals_lvl / statistics.mean(times) / 1e6, "unit": "Mpix·cand/s", "cores": threads,
                                "physical_cores": physical, "kind": kind_cpu, "steps": len(times),
                                "seconds_per_step": statistics.mean(times),
                                "sample": "processLevel of level %d of %d (%dx%d, all %d cameras: random proposals, ping-pong, "
                                          "joint bilateral, median) started from the CUDA library's level-%d disparities; work = "
                                          "the %d cost evaluations the CUDA library counts for that level" % (
                                              C2F_CPU_LEVEL, C2F_LEVELS, W >> C2F_CPU_LEVEL, H >> C2F_CPU_LEVEL, S,
                                              C2F_CPU_LEVEL + 1, evals_lvl)}
        good = tot = nanbad = 0
        for g, o in zip(c2f.level_out[C2F_CPU_LEVEL], cpu_disp):
            fin = ~np.isnan(o)
            nanbad += int((np.isnan(g) != np.isnan(o)).sum())
            good += int((np.abs(g - o)[fin] <= 1e-3 * np.abs(o)[fin]).sum())
            tot += int(fin.sum())
        line["parity"] = {"against": kind_cpu, "level": C2F_CPU_LEVEL, "pixels": tot, "within_1e-3_rel": good,
                          "fraction": good / max(1, tot), "nan_pattern_mismatches": nanbad}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_reference_c2f(args, rig, colors):
    """Reference arm of workload c2f5, reference code only: the reference's own processLevel of the coarsest level (its
    compile-time 150 candidates) feeds its own processLevel of level C2F_CPU_LEVEL, which is the timed step.  The work of
    that step is counted by the oracle restatement (bit-identical to the reference, tests/test_reference_pin.py) run
    once, untimed, on the same input — the reference keeps no counters."""
    from facebook360_dep_b200 import capi
    from tests import oracle_libs
    S, W, H, D, kind = WORKLOADS["c2f5"]
    lib = oracle_libs.load_ref()
    kind_cpu = "reference" if lib is not None else "port"
    oracle = oracle_libs.load_oracle()
    if lib is None:
        lib = oracle
    threads, physical = host_threads()
    lib.set_threads(threads)
    oracle.set_threads(threads)
    lv = {k: [lib.downscale_area(c, W >> k, H >> k) for c in colors] for k in (C2F_CPU_LEVEL + 1, C2F_CPU_LEVEL)}
    ctx = capi.Context(lib, capi.rig_descs(rig))
    k = C2F_CPU_LEVEL + 1
    ctx.level_begin(W >> k, H >> k, level=k, num_levels=C2F_LEVELS, full_width=W, full_height=H)
    ctx.set_colors(lv[k])
    if k < C2F_LEVELS - 1:
        raise SystemExit("C2F_CPU_LEVEL must be the second coarsest level")
    ctx.process_level(num_depths=150, min_depth_m=MIN_DEPTH, max_depth_m=MAX_DEPTH, partial_coverage=True)
    coarse = [ctx.get_disparity(d, want_cost=False) for d in range(S)]
    ctx.close()
    _, _, evals = c2f_cpu_level(oracle, rig, lv[C2F_CPU_LEVEL], coarse, threads, steps=1, warmup=0)
    warm = max(1, args.warmup)
    times, _, _ = c2f_cpu_level(lib, rig, lv[C2F_CPU_LEVEL], coarse, threads, steps=args.steps, warmup=warm)
    value = evals / statistics.mean(times) / 1e6
    base = {"value": value, "unit": "Mpix·cand/s", "cores": threads, "physical_cores": physical, "kind": kind_cpu,
            "steps": len(times), "seconds_per_step": statistics.mean(times),
            "sample": "processLevel of level %d of %d (%dx%d, all %d cameras) from the reference's own level-%d result; %d cost "
                      "evaluations per step (counted by the oracle restatement)" % (
                          C2F_CPU_LEVEL, C2F_LEVELS, W >> C2F_CPU_LEVEL, H >> C2F_CPU_LEVEL, S, C2F_CPU_LEVEL + 1, evals)}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Mpix·cand/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": warm, "ms_per_step": 1e3 * statistics.mean(times), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 cost, f64 projection, u16 texels", "data": "synthetic",
            "config": static_config("c2f5"), "cpu_baseline": base,
            "e2e": {"value": value, "unit": "Mpix·cand/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def _dist_setup():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    return torch, dist, world, rank, local_rank, dev


def run_cfg4(args):
    """Workload cfg4 (BASELINE.json configs[3]): one 24-camera 4096^2 frame per step, STRONG scaling — the destination
    cameras are dealt round-robin to the GPUs (shard.camera_shard).  Level 1 (2048^2): brute force with 256 candidates;
    level 0 (4096^2): proposals, ping-pong, mismatch handling, joint bilateral (radius 5), median.  Every stage is
    independent per destination except mismatch handling, which reads every camera's pre-update disparity: ONE NCCL
    all-gather of the per-camera planes per mismatch level, inside the timed region."""
    torch, dist, world, rank, local_rank, dev = _dist_setup()
    from facebook360_dep_b200 import capi, pipeline, shard, synth
    S, W, H, D, kind = WORKLOADS["cfg4"]
    rig = synth.ring_rig(S, W, H, kind=kind)
    t0 = time.time()
    colors, _ = synth.render_rig(rig, W, H, scene=synth.Scene(seed=42), device=dev)
    log("[bench] rank %d rendered %d x %dx%d frames in %.1fs" % (rank, S, W, H, time.time() - t0))
    cuda = capi.load_cuda()
    own = shard.camera_shard(S, world, rank)
    ctx = capi.Context(cuda, capi.rig_descs(rig), own, device=local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    full = [torch.from_numpy(np.ascontiguousarray(c)).to(dev) for c in colors]
    half = [torch.empty((H // 2, W // 2, 3), dtype=torch.uint16, device=dev) for _ in range(S)]
    for s in range(S):
        cuda.check(cuda.lib.derp_downscale_area(local_rank, full[s].data_ptr(), W, H, half[s].data_ptr(), W // 2, H // 2))
    del colors
    kw = dict(num_depths=D, min_depth_m=MIN_DEPTH, max_depth_m=MAX_DEPTH, partial_coverage=True, mismatches_start_level=0)
    stats = {"evals": 0, "gather_bytes": 0, "gather_ms": 0.0}

    def frame():
        stats["evals"] = 0
        stats["gather_bytes"] = 0
        stats["gather_ms"] = 0.0
        for level, imgs, (w, h) in ((1, half, (W // 2, H // 2)), (0, full, (W, H))):
            ctx.level_begin(w, h, level=level, num_levels=2, full_width=W, full_height=H)
            ctx.set_colors_ptr([t.data_ptr() for t in imgs])
            if level == 0:
                for d in range(len(own)):
                    ctx.upsample_from_kept(d)
            ctx.level_estimate(**kw)
            stats["evals"] += ctx.get_counters()[0]
            if level == 0:  # Derp.cpp:726-728: mismatch handling on levels <= mismatches_start_level, not the coarsest
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record(stream)
                stats["gather_bytes"] += pipeline.all_gather_disparities_device(ctx, S, dev)
                g1.record(stream)
                torch.cuda.synchronize()
                stats["gather_ms"] += g0.elapsed_time(g1)
                if world > 1:
                    dist.barrier()  # every rank holds its copy before any rank updates its planes
                ctx.mismatches_gathered()
            ctx.level_filter(**kw)
            if level == 1:
                ctx.level_keep()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(1, args.warmup)):
        frame()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ctx.profile(True)
    l0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        frame()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - l0
    sweep_ms, sweep_n = ctx.get_profile()
    ctx.profile(False)
    clocks = sampler.stop() if rank == 0 else None
    ms, evals_all = shard.reduce_step(ms, stats["evals"], dev)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    peak, peak_src = measured_peak_gbs()
    line = {
        "metric": METRIC, "value": evals_all * args.steps / (ms / 1e3) / 1e6, "unit": "Mpix·cand/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(1, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32 cost, f64 projection, u16 texels", "data": "synthetic",
        "config": dict(static_config("cfg4"), parallelism="destination cameras round-robin over %d GPU(s)" % world, levels=2,
                       frames_per_step=1),
        "work": {"cost_evaluations_per_frame": evals_all},
        "clocks": clocks, "gpu_launches": int(launches),
        "exchange": {"collective": "NCCL all-gather of the per-camera disparity planes before mismatch handling (level 0)",
                     "bytes_received_per_rank_per_frame": stats["gather_bytes"], "ms_per_frame_rank0": stats["gather_ms"],
                     "GBps_rank0": (stats["gather_bytes"] / 1e9) / (stats["gather_ms"] / 1e3) if stats["gather_ms"] else None},
        "roofline": {"bound": "hbm", "kernel": "sweep (rank 0's destinations, level 1)", "peak": peak, "unit": "GB/s",
                     "peak_source": peak_src, "ms_per_launch": sweep_ms / max(1, sweep_n), "launches_timed": int(sweep_n),
                     "kernel_share_of_step": sweep_ms / ms if ms else None, "achieved": None, "frac": None, "traffic": None,
                     "note": "see workload bf128_l0 for the sweep's roofline; here the share of the step and the exchange are "
                             "what is measured"},
        "e2e": None,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_cfg5(args):
    """Workload cfg5 (BASELINE.json configs[4]): a 30-frame sequence of the 16-camera rig, contiguous frame blocks per
    GPU.  Like scripts/render/pipeline.py:364-408, every pyramid level is estimated for all frames (processLevel), then
    temporally filtered (TemporalBilateralFilter, time_radius 2), and the FILTERED level feeds the next finer one.  The
    filter needs the colour and disparity of the +-2 neighbouring frames: the boundary frames travel rank to rank over
    NVLink (NCCL send/recv of device tensors) inside the timed region."""
    torch, dist, world, rank, local_rank, dev = _dist_setup()
    from facebook360_dep_b200 import capi, pipeline, shard, synth
    S, W, H, D, kind = WORKLOADS["cfg5"]
    F = CFG5_FRAMES if not args.frames else args.frames
    rig = synth.ring_rig(S, W, H, kind=kind)
    cuda = capi.load_cuda()
    ctx = capi.Context(cuda, capi.rig_descs(rig), device=local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    first, last = shard.frame_block(F, world, rank)
    levels = C2F_LEVELS
    pyr = {}  # frame -> level -> uint16 [S,h,w,3] on the device
    t0 = time.time()
    for f in range(first, last):
        colors, _ = synth.render_rig(rig, W, H, scene=synth.Scene(seed=42, shift=(0.01 * f, 0, 0)), device=dev)
        lv = {0: torch.stack([torch.from_numpy(np.ascontiguousarray(c)).to(dev) for c in colors])}
        for k in range(1, levels):
            lv[k] = torch.empty((S, H >> k, W >> k, 3), dtype=torch.uint16, device=dev)
            for s in range(S):
                cuda.check(cuda.lib.derp_downscale_area(local_rank, lv[0][s].data_ptr(), W, H, lv[k][s].data_ptr(), W >> k, H >> k))
        pyr[f] = lv
    log("[bench] rank %d: frames %d..%d rendered and resized in %.1fs" % (rank, first, last - 1, time.time() - t0))
    fov = {}
    for k in range(levels):
        ctx.level_begin(W >> k, H >> k, level=k, num_levels=levels, full_width=W, full_height=H)
        fov[k] = torch.stack([torch.from_numpy(ctx.get_fov_mask(d)).to(dev) for d in range(S)])
    kw = dict(num_depths=D, min_depth_m=MIN_DEPTH, max_depth_m=MAX_DEPTH, partial_coverage=True)
    stats = {"evals": 0, "halo_bytes": 0, "halo_ms": 0.0}

    def sequence():
        stats["evals"] = 0
        stats["halo_bytes"] = 0
        filtered = {}  # frame -> f32 [S,h,w] of the level just finished
        for k in range(levels - 1, -1, -1):
            w, h = W >> k, H >> k
            est = {}
            for f in range(first, last):
                ctx.level_begin(w, h, level=k, num_levels=levels, full_width=W, full_height=H)
                ctx.set_colors_ptr([pyr[f][k][s].data_ptr() for s in range(S)])
                if k < levels - 1:
                    for d in range(S):
                        cuda.check(cuda.lib.derp_upsample_from(ctx.h, d, filtered[f][d].data_ptr(), W >> (k + 1), H >> (k + 1), None, None))
                ctx.process_level(**kw)
                stats["evals"] += ctx.get_counters()[0]
                out = torch.empty((S, h, w), dtype=torch.float32, device=dev)
                for d in range(S):
                    cuda.check(cuda.lib.derp_get_disparity(ctx.h, d, out[d].data_ptr(), None, None))
                est[f] = out
            ctx.sync()
            local = {f: (pyr[f][k], est[f]) for f in range(first, last)}
            filtered, nbytes = pipeline.temporal_filter_block_device(cuda, local, F, fov[k], time_radius=2, gpu=local_rank)
            stats["halo_bytes"] += nbytes
        return filtered

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(1, args.warmup)):
        sequence()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        sequence()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    # the exchange alone, for its bandwidth: level-0 boundary frames
    local0 = {f: (pyr[f][0], torch.zeros((S, H, W), dtype=torch.float32, device=dev)) for f in range(first, last)}
    barrier()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record(stream)
    _, halo0 = pipeline.exchange_halos_device(local0, F, 2)
    h1.record(stream)
    barrier()
    halo_ms = h0.elapsed_time(h1)
    ms, evals_all = shard.reduce_step(ms, stats["evals"], dev)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    line = {
        "metric": METRIC, "value": evals_all * args.steps / (ms / 1e3) / 1e6, "unit": "Mpix·cand/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(1, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32 cost, f64 projection, u16 texels", "data": "synthetic",
        "config": dict(static_config("cfg5"), frames=F, levels=levels, time_radius=2,
                       parallelism="contiguous frame blocks over %d GPU(s), +-2-frame halo over NCCL" % world),
        "work": {"cost_evaluations_per_sequence": evals_all, "frames_per_s": F * args.steps / (ms / 1e3)},
        "clocks": clocks, "gpu_launches": int(launches),
        "exchange": {"collective": "NCCL send/recv of the +-2 boundary frames (colour + disparity) after every level",
                     "bytes_received_rank0_per_sequence": stats["halo_bytes"],
                     "level0_bytes_received_rank0": halo0, "level0_ms": halo_ms,
                     "level0_GBps_rank0": (halo0 / 1e9) / (halo_ms / 1e3) if halo_ms and halo0 else None},
        "roofline": None, "e2e": None,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def cfg1_full(cuda, device):
    """BASELINE.json configs[0] in full on both sides: 4-camera rectilinear rig, 512x512, 32 candidates, one level.
    CPU: every candidate slice of every destination through the reference's own code (winner = first strict minimum
    over the slices, Derp.cpp:323-333); GPU: derp_brute_force.  Reports both rates and the winner-index comparison."""
    import ctypes as C
    import torch
    from facebook360_dep_b200 import capi
    from tests import oracle_libs
    S, W, H, D, kind = WORKLOADS["bf32_cfg1"]
    rig, colors = make_inputs("bf32_cfg1", device)
    colors = [np.ascontiguousarray(c) for c in colors]
    table = candidate_table(D)
    lib = oracle_libs.load_ref()
    kind_cpu = "reference" if lib is not None else "port"
    if lib is None:
        lib = oracle_libs.load_oracle()
    threads, physical = host_threads()
    lib.set_threads(threads)
    cctx = capi.Context(lib, capi.rig_descs(rig))
    cctx.level_begin(W, H)
    cctx.set_colors(colors)
    gctx = capi.Context(cuda, capi.rig_descs(rig), device=device.index or 0)
    gctx.level_begin(W, H)
    gctx.set_colors(colors)
    cpu_s = gpu_s = 0.0
    evals = flips = pixels = 0
    for d in range(S):
        cctx.reproject(d)
        if kind_cpu == "reference":
            f = lib.lib.derp_ref_cost_slices
            f.restype = C.c_int
            f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
            costs = np.empty((D, H - 2, W), np.float32)
            t0 = time.perf_counter()
            lib.check(f(cctx.h, d, table.ctypes.data, D, 1, H - 1, costs.ctypes.data, None))
            cpu_s += time.perf_counter() - t0
            w = np.where(np.isnan(costs), np.float32(np.inf), costs)
            cpu_idx = np.where(np.isinf(w).all(axis=0) | (w.min(axis=0) >= np.float32(3.4028235e38)), -1, np.argmin(w, axis=0))[:, 1:W - 1]
        else:
            t0 = time.perf_counter()
            cpu_idx = cctx.brute_force(d, num_depths=D, min_depth_m=MIN_DEPTH, max_depth_m=MAX_DEPTH)[1:H - 1, 1:W - 1]
            cpu_s += time.perf_counter() - t0
        gctx.reproject(d)
        gctx.brute_force(d, num_depths=D, min_depth_m=MIN_DEPTH, max_depth_m=MAX_DEPTH)  # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        gi = gctx.brute_force(d, num_depths=D, min_depth_m=MIN_DEPTH, max_depth_m=MAX_DEPTH, want_index=False)
        gctx.sync()
        gpu_s += time.perf_counter() - t0
        gi = gctx.brute_force(d, num_depths=D, min_depth_m=MIN_DEPTH, max_depth_m=MAX_DEPTH)[1:H - 1, 1:W - 1]
        evals += gctx.get_counters()[0]
        fov = cctx.get_fov_mask(d)[1:H - 1, 1:W - 1].astype(bool)
        flips += int((gi != cpu_idx)[fov].sum())
        pixels += int(fov.sum())
    cctx.close()
    gctx.close()
    return {"workload": "bf32_cfg1", "pixel_candidates": int(evals), "cpu": {"value": evals / cpu_s / 1e6, "unit": "Mpix·cand/s",
            "kind": kind_cpu, "cores": threads, "seconds": cpu_s}, "gpu": {"value": evals / gpu_s / 1e6, "unit": "Mpix·cand/s",
            "seconds": gpu_s, "note": "reprojection excluded, wall clock around derp_brute_force"},
            "parity": {"pixels": pixels, "index_mismatches": flips}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="bf128_l0", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-c2f", action="store_true")
    ap.add_argument("--no-cfg1", action="store_true")
    ap.add_argument("--frames", type=int, default=0, help="cfg5: sequence length (default 30)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "c2f5":
        return run_c2f(args)
    if args.workload == "cfg4":
        return run_cfg4(args)
    if args.workload == "cfg5":
        return run_cfg5(args)

    import torch
    import torch.distributed as dist
    from facebook360_dep_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    S, W, H, D, kind = WORKLOADS[args.workload]
    rig, colors = make_inputs(args.workload, dev, rank)
    # pinned host staging of the inputs and outputs (e2e path)
    pin_colors = [torch.from_numpy(c).pin_memory() for c in colors]
    pin_np = [t.numpy() for t in pin_colors]
    pin_out = [torch.empty((H, W), dtype=torch.float32).pin_memory() for _ in range(S)]
    pin_out_np = [t.numpy() for t in pin_out]

    cuda = capi.load_cuda()  # no fallback: raises if the extension is missing
    ctx = capi.Context(cuda, capi.rig_descs(rig), device=local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    ctx.level_begin(W, H)
    ctx.set_colors(pin_np)
    ctx.sync()

    def sweep_all():
        for d in range(S):
            ctx.reproject(d)
            ctx.brute_force(d, num_depths=D, min_depth_m=MIN_DEPTH, max_depth_m=MAX_DEPTH, partial_coverage=True,
                            want_index=False)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- work count of one step (exact, from the kernel's counters) ----
    evals_step = hits_step = 0
    for d in range(S):
        ctx.reproject(d)
        ctx.brute_force(d, num_depths=D, min_depth_m=MIN_DEPTH, max_depth_m=MAX_DEPTH, partial_coverage=True,
                        want_index=False)
        e, h = ctx.get_counters()
        evals_step += e
        hits_step += h
    vbar = hits_step / max(1, evals_step)
    log("[bench] rank %d: %.3f G pixel·cand per step, vbar %.2f" % (rank, evals_step / 1e9, vbar))

    # ---- value: inputs resident in HBM ----
    for _ in range(max(0, args.warmup - 1)):  # the counting pass above was one more warm-up step
        sweep_all()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ctx.profile(True)
    l0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        sweep_all()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - l0
    sweep_ms, sweep_n = ctx.get_profile()
    ctx.profile(False)
    clocks = sampler.stop() if rank == 0 else None
    refined, seeds = ctx.sweep_stats()  # of the last destination: exact evaluations the filtered sweep performed
    refine_frac = (refined + seeds) / max(1, evals_step / S) if seeds else None

    # ---- e2e: host buffers in, host buffers out, through the C ABI ----
    e2e_ms = None
    h2d = S * W * H * 6
    d2h = S * W * H * 4
    if not args.no_e2e:
        def e2e_step():
            ctx.set_colors(pin_np)  # pinned host -> device + variance
            for d in range(S):
                ctx.reproject(d)
                ctx.brute_force(d, num_depths=D, min_depth_m=MIN_DEPTH, max_depth_m=MAX_DEPTH, partial_coverage=True,
                                want_index=False)
            for d in range(S):
                cuda.check(cuda.lib.derp_get_disparity(ctx.h, d, pin_out_np[d].ctypes.data, None, None))
        e2e_step()  # warm-up
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            e2e_step()
        e1.record(stream)
        barrier()
        e2e_ms = e0.elapsed_time(e1)

    # ---- max over ranks (time) / sum over ranks (work): facebook360_dep_b200/shard.py ----
    from facebook360_dep_b200 import shard
    ms, evals_all = shard.reduce_step(ms, evals_step, dev)
    e2e_max, _ = shard.reduce_step(e2e_ms if e2e_ms is not None else 0.0, 0.0, dev)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    value = evals_all * args.steps / (ms / 1e3) / 1e6
    peak, peak_src = measured_peak_gbs()
    # roofline of the dominant kernel (rank 0's launches): B_stream bytes per launch / mean launch time
    alg_bytes_step = 20.0 * hits_step + 30.0 * (evals_step / D)
    alg_bytes_launch = alg_bytes_step / S
    sweep_ms_launch = sweep_ms / max(1, sweep_n)
    achieved = alg_bytes_launch / (sweep_ms_launch / 1e3) / 1e9 if sweep_n else None
    line = {
        "metric": METRIC, "value": value, "unit": "Mpix·cand/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 cost, f64 projection, u16 texels", "data": "synthetic",
        "config": static_config(args.workload),
        "work": {"vbar": round(vbar, 3), "pixel_cand_per_step_per_gpu": evals_step, "triples_per_step_per_gpu": hits_step},
        "clocks": clocks,
        "e2e": None if e2e_ms is None else {
            "value": evals_all * args.steps / (e2e_max / 1e3) / 1e6, "unit": "Mpix·cand/s",
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_max / args.steps},
        "gpu_launches": int(launches),
        "roofline": {
            "bound": "hbm", "kernel": "filtered sweep: sweepLowerKernel (97 %) + sweepSeedKernel + refineListKernel + refineKernel",
            "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": (achieved / peak) if achieved else None, "traffic": ncu_traffic_per_launch(args.workload),
            "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes_launch,
            "ms_per_launch": sweep_ms_launch, "launches_timed": int(sweep_n),
            "kernel_share_of_step": (sweep_ms / ms) if ms else None,
            "note": "B_stream = 20 B x (pixel,cand,source) triples + 30 B x pixels (SURVEY.md 8(d)); ms_per_launch = the whole "
                    "sweep of one destination (all four kernels). The sweep is co-limited by instruction issue (64 %) and the "
                    "L1/shared-memory pipe (70 %), not by HBM - see DESIGN.md",
            "exact_evaluations_fraction": refine_frac,
            "triples_per_s": hits_step / S / (sweep_ms_launch / 1e3) if sweep_n else None,
            "issue": issue_roof(args.workload, sweep_ms_launch if sweep_n else None, (clocks or {}).get("sm_mhz"))},
    }
    if world == 1 and args.workload == "bf128_l0" and not args.no_c2f:
        c2f = CoarseToFine(cuda, ctx, pin_colors, S, W, H, D, stream, dev, C2F_LEVELS)
        c2f.frame(False)  # uploads, pyramid, warm-up
        c2f.frame(True)
        t = []
        for resident in (True, False):
            torch.cuda.synchronize()
            q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            q0.record(stream)
            c2f.frame(resident)
            q1.record(stream)
            torch.cuda.synchronize()
            t.append(q0.elapsed_time(q1))
        line["coarse_to_fine_5level"] = {
            "ms_per_frame": t[0], "e2e_ms_per_frame": t[1], "cost_evaluations": c2f.evals, "value": c2f.evals / t[0] / 1e3,
            "unit": "Mpix·cand/s", "levels": C2F_LEVELS,
            "note": "configs[1] as the reference runs it; full line: python bench.py --workload c2f5"}
    if world == 1 and not args.no_cpu_baseline:
        arm = CpuArm(args.workload, rig, [np.ascontiguousarray(c) for c in colors])
        arm.calibrate(target_s=2.5)
        res = arm.sample(steps=3, warmup=1, want_costs=True)
        line["cpu_baseline"] = arm.baseline_object(res)
        # parity at the benched size, inside the run that benches it: same frame, same candidates, CPU vs CUDA
        ctx.level_begin(W, H)
        ctx.set_colors(pin_np)
        line["parity"] = parity_vs_cpu(arm, res, ctx)
        arm.close()
        if args.workload == "bf128_l0" and not args.no_cfg1:
            line["cfg1_full"] = cfg1_full(cuda, dev)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
