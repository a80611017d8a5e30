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
}
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


def cpu_sample(workload, rig, colors, steps=1, warmup=0):
    """Oracle (port of the reference CPU path) on a bounded sample: 1 dst camera x 8 evenly spaced candidates
    x full frame, all host threads.  Returns (evals_per_s list, cores, sample description, vbar)."""
    from facebook360_dep_b200 import capi
    S, W, H, D, kind = WORKLOADS[workload]
    oracle = capi.load_oracle()  # bench.py's cpu_baseline / --impl reference legs only
    ncpu = os.cpu_count() or 1
    oracle.set_threads(ncpu)
    ctx = capi.Context(oracle, capi.rig_descs(rig))
    ctx.level_begin(W, H)
    ctx.set_colors(colors)
    ctx.reproject(0)
    rates = []
    vbar = None
    ncand = 8 if W >= 1024 else D
    # all the host threads it can use: SMT siblings can hurt this fp-heavy loop, so calibrate on 2 candidates
    cores, best = ncpu, 0.0
    for t in sorted({ncpu, max(1, ncpu // 2)}, reverse=True):
        oracle.set_threads(t)
        t0 = time.perf_counter()
        ctx.brute_force(0, num_depths=2, min_depth_m=MIN_DEPTH, max_depth_m=MAX_DEPTH, partial_coverage=True,
                        want_index=False)
        r = ctx.get_counters()[0] / (time.perf_counter() - t0)
        if r > best:
            best, cores = r, t
    oracle.set_threads(cores)
    log("[bench] cpu baseline uses %d of %d host threads" % (cores, ncpu))
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        ctx.brute_force(0, num_depths=ncand, min_depth_m=MIN_DEPTH, max_depth_m=MAX_DEPTH, partial_coverage=True,
                        want_index=False)
        dt = time.perf_counter() - t0
        evals, hits = ctx.get_counters()
        vbar = hits / max(evals, 1)
        if i >= warmup:
            rates.append(evals / dt)
        log("[bench] cpu sample step %d: %.2fs, %.3f Mpix·cand/s, vbar %.2f" % (i, dt, evals / dt / 1e6, vbar))
    ctx.close()
    sample = "1 dst camera x %d evenly spaced candidates x full %dx%d frame (%d-cam rig), brute force only" % (
        ncand, W, H, S)
    return rates, cores, sample, vbar


def coarse_to_fine(ctx, colors, S, W, H, D, stream, levels=5):
    """BASELINE.json configs[1] as the reference runs it (DerpCLI defaults): brute force with D candidates at the
    coarsest of 5 levels, then random proposals (2) + ping-pong (1) + joint bilateral + median on every finer
    level, levels handed over in HBM->host->HBM like the PFM round trip.  Reported beside the headline; the
    second of two passes is timed (geometry caches warm, as from the second frame of a sequence on)."""
    import torch
    from facebook360_dep_b200 import synth
    pyr = [colors]
    for _ in range(1, levels):
        pyr.append([synth.downscale_area(c, 2) for c in pyr[-1]])
    out = None
    for rep in range(2):
        prev, tot_ms, tot_e = None, 0.0, 0
        for level in range(levels - 1, -1, -1):
            w, h = W >> level, H >> level
            ctx.level_begin(w, h, level=level, num_levels=levels, full_width=W, full_height=H)
            ctx.set_colors(pyr[level])
            if prev is not None:
                for d in range(S):
                    ctx.upsample_from(d, prev[d])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.process_level(num_depths=D, min_depth_m=MIN_DEPTH, max_depth_m=MAX_DEPTH, partial_coverage=True)
            e1.record(stream)
            torch.cuda.synchronize()
            tot_ms += e0.elapsed_time(e1)
            tot_e += ctx.get_counters()[0]
            prev = [ctx.get_disparity(d, want_cost=False) for d in range(S)]
        out = {"ms_per_frame": tot_ms, "cost_evaluations": tot_e, "value": tot_e / tot_ms / 1e3, "unit": "Mpix·cand/s",
               "levels": levels, "note": "process_level time only (uploads / level hand-over excluded)"}
    return out


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path.  The reference cannot be compiled in this
    image (OpenCV C++/Eigen/Boost/gflags/glog/folly absent), so this is the oracle port, all host threads."""
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
    t0 = time.perf_counter()
    rates, cores, sample, vbar = cpu_sample(args.workload, rig, colors, steps=args.steps, warmup=min(args.warmup, 1))
    total = time.perf_counter() - t0
    value = statistics.mean(rates) / 1e6
    S, W, H, D, kind = WORKLOADS[args.workload]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpix·cand/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1),
        "ms_per_step": 1e3 * total / max(1, args.steps + min(args.warmup, 1)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 cost, f64 projection, u16 texels", "data": "synthetic",
        "config": {"workload": args.workload, "cameras": S, "width": W, "height": H, "candidates": D,
                   "camera_model": kind, "note": "each step = bounded sample of the workload"},
        "cpu_baseline": {"value": value, "unit": "Mpix·cand/s", "cores": cores, "kind": "port", "sample": sample,
                         "vbar": vbar},
        "e2e": {"value": value, "unit": "Mpix·cand/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


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
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

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
        "config": {"workload": args.workload, "cameras": S, "width": W, "height": H, "candidates": D,
                   "camera_model": kind, "frames_per_step_per_gpu": 1, "vbar": round(vbar, 3),
                   "pixel_cand_per_step_per_gpu": evals_step,
                   "l2": "inputs larger than L2 (per destination %.0f MB of pair tables vs 126 MB L2)" % (
                       (S - 1) * W * H * 24 / 1e6)},
        "clocks": clocks,
        "e2e": None if e2e_ms is None else {
            "value": evals_all * args.steps / (e2e_max / 1e3) / 1e6, "unit": "Mpix·cand/s",
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_max / args.steps},
        "gpu_launches": int(launches),
        "roofline": {
            "bound": "hbm", "kernel": "sweepKernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": (achieved / peak) if achieved else None, "traffic": ncu_traffic_per_launch(args.workload),
            "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes_launch,
            "ms_per_launch": sweep_ms_launch, "launches_timed": int(sweep_n),
            "kernel_share_of_step": (sweep_ms / ms) if ms else None,
            "note": "B_stream = 20 B x (pixel,cand,source) triples + 30 B x pixels (SURVEY.md 8(d)); the kernel is "
                    "FP32/FP64-issue bound, not HBM bound - see DESIGN.md",
            "triples_per_s": hits_step / S / (sweep_ms_launch / 1e3) if sweep_n else None,
            "issue": issue_roof(args.workload, sweep_ms_launch if sweep_n else None, (clocks or {}).get("sm_mhz"))},
    }
    if world == 1 and args.workload == "bf128_l0" and not args.no_c2f:
        line["coarse_to_fine_5level"] = coarse_to_fine(ctx, colors, S, W, H, D, stream)
    if world == 1 and not args.no_cpu_baseline:
        cpu_colors = [np.ascontiguousarray(c) for c in colors]
        rates, cores, sample, cvbar = cpu_sample(args.workload, rig, cpu_colors, steps=1, warmup=0)
        line["cpu_baseline"] = {"value": statistics.mean(rates) / 1e6, "unit": "Mpix·cand/s", "cores": cores,
                                "kind": "port", "sample": sample, "vbar": cvbar}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
