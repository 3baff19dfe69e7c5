#!/usr/bin/env python
"""Benchmark of the per-frame geometry pipeline (BASELINE.json metric: 4K
visibility-buffer frames/s, culled Mtri/s, % HBM roofline).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2]

A step is one frame: reset -> instance cull + scan -> meshlet cull -> vis-buffer
raster -> (composite) -> depth min/max -> cascade matrices -> shadow cull/raster x4
-> lighting resolve (src/rendering.cpp steps 1-11) of BASELINE config 2
(10 000 instances of a 100 000-triangle torus at 3840x2160).  `value` times the
frame with every input resident in HBM (CUDA events on the context stream, L2
flushed between frames); `e2e` times lh_render_frame_host: host Uniforms +
host Instance array in, frame statistics out, copies inside the timed region.
--impl reference times the CPU restatement of the reference (oracle/) on the
host cores; it is a reported baseline, not the target.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "visbuffer_frames_per_s"
UNIT = "frames/s"
WORKLOADS = {
    "c1": "BASELINE config 1: 1000 x torus 20k tri, 1920x1080",
    "c2": "BASELINE config 2: 10000 x torus 100k tri (1.0 G tri), 3840x2160, frustum+cone cull, vis-buffer, 4 cascades, lighting resolve",
    "c4": "BASELINE config 4: ~100 M tri mixed meshes, 5% alpha-clip, BC7 textures, 3840x2160",
    "c5": "BASELINE config 5: 10000 x torus 100k tri, 7680x4320",
    "tiny": "unit-test scene",
}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class NvmlSampler:
    """SM clock and throttle reasons sampled in-process through NVML every 10 ms DURING the timed region (the timed
    region of this bench is tens of milliseconds: nvidia-smi's 100 ms loop would see next to nothing of it)."""

    def __init__(self, device):
        self.device = device
        self.ok = False
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for n, b in names.items():
                    if bits & b:
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.010)

    def start(self):
        if self.ok:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        if not self.ok:
            return None
        self._stop.set()
        self._thread.join(timeout=2)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz, "samples": len(self.sm),
                "reasons": sorted(self.reasons), "how": "NVML, 10 ms period, in-process"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for k, n in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.path)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def build_scene(workload):
    from lighthugger_b200.scene import Scene
    return Scene.preset(workload)


def scene_textures(scene, size):
    from helpers import Textures
    return Textures(scene, size) if scene.num_texture_sets else None


# ---------------------------------------------------------------- reference arm / cpu baseline
def cpu_frame_times(scene, tex, sky, sample_instances, threads=0):
    """Times the oracle (CPU restatement of the reference) pass by pass on the first `sample_instances` instances at
    the workload's full resolution.  Returns (seconds of instance-proportional passes, seconds of per-pixel passes)."""
    from helpers import make_oracle
    o = make_oracle(scene, textures=tex, skybox=sky, threads=threads, instance_range=(0, sample_instances))
    t = {}

    def run(name, fn):
        t0 = time.perf_counter()
        fn()
        t[name] = time.perf_counter() - t0

    run("reset", o.reset)
    run("cull_instances", o.cull_instances)
    run("cull_meshlets", o.cull_meshlets)
    run("visbuffer", o.raster_visbuffer)
    run("depth_reduction", o.reduce_depth)
    run("shadow_matrices", o.generate_shadow_matrices)
    run("cull_instances_shadows", o.cull_instances_shadows)
    run("shadowmaps", lambda: [o.shadow_cascade(c) for c in range(4)])
    run("render_geometry", o.resolve)
    proportional = t["cull_instances"] + t["cull_meshlets"] + t["visbuffer"] + t["cull_instances_shadows"] + t["shadowmaps"]
    fixed = t["reset"] + t["depth_reduction"] + t["shadow_matrices"] + t["render_geometry"]
    return proportional, fixed, t


def cpu_baseline(scene, tex, sky, budget_s):
    """One bounded sample: instance subset sized from a calibration run so that the sample takes about `budget_s`."""
    n = scene.num_instances
    cores = os.cpu_count() or 1
    probe = max(1, min(n, n // 100))
    p, f, _ = cpu_frame_times(scene, tex, sky, probe)
    per_instance = max(p / probe, 1e-9)
    sample = int(max(probe, min(n, (budget_s - f) / per_instance)))
    p, f, passes = cpu_frame_times(scene, tex, sky, sample)
    full_frame_s = f + p * (n / sample)
    return {
        "value": 1.0 / full_frame_s, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": (f"oracle (CPU restatement of the reference; lavapipe not installed) on instances [0,{sample}) of {n} at full "
                   f"resolution, all passes: {p + f:.2f} s measured; instance-proportional passes scaled x{n / sample:.2f}, "
                   f"per-pixel passes as measured -> {full_frame_s:.2f} s/frame"),
        "culled_mtri_per_s": scene.total_triangles / full_frame_s / 1e6,
        "passes_s": {k: round(v, 4) for k, v in passes.items()},
    }


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    scene = build_scene(args.workload)
    tex = scene_textures(scene, args.tex_size)
    sky = scene.make_skybox(32)
    total = max(1, args.steps + args.warmup)
    budget = max(2.0, min(30.0, 150.0 / total))
    times = []
    base = None
    for i in range(total):
        base = cpu_baseline(scene, tex, sky, budget)
        if i >= args.warmup:
            times.append(1.0 / base["value"])
    ms = 1e3 * float(np.mean(times))
    value = 1e3 / ms
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "impl": "reference",
        "config": {"workload": WORKLOADS[args.workload], "resolution": [scene.width, scene.height], "instances": scene.num_instances,
                   "input_triangles": scene.total_triangles},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": base["cores"], "kind": base["kind"], "sample": base["sample"]},
        "culled_mtri_per_s": scene.total_triangles / (ms / 1e3) / 1e6,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def meshlet_builder():
    """Which meshlet builder the host library was built with (the reference's vendored meshoptimizer, or the greedy fallback)."""
    from lighthugger_b200.scene import host_lib
    return "meshoptimizer (reference's vendored copy)" if host_lib().lh_host_has_meshoptimizer() else "greedy fallback"


# ---------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    from lighthugger_b200 import _structs as S
    from helpers import make_context

    rank, local_rank, world = dist_env()
    n_gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    scene = build_scene(args.workload)
    tex = scene_textures(scene, args.tex_size)
    sky = scene.make_skybox(32)
    from lighthugger_b200 import sortlast
    if args.stripes_per_rank > 0:
        sortlast.STRIPES_PER_RANK = args.stripes_per_rank
    n = scene.num_instances
    (first, count), _ = sortlast.main_plan(n, rank, world, args.partition)
    end = first + count
    caps = sortlast.capacities(scene, world, args.shadow_split)
    emulate = None
    if args.emulate_rank and world == 1:
        # profiling aid: one GPU does the local work of rank R of N (stripe partition, no exchange), so that a launch list
        # of a single rank of a large world can be captured on one device; not a bench value
        emulate = tuple(int(x) for x in args.emulate_rank.split("/"))
        caps = sortlast.capacities(scene, emulate[1], args.shadow_split)
    ctx = make_context(scene, textures=tex, skybox=sky, caps=caps, flags=S.LH_CONFIG_WRITE_DRAW_COMMANDS | S.LH_CONFIG_PASS_TIMINGS,
                       device=local_rank)
    if world > 1:
        ids = [ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(rank, world, ids[0])
        ctx.comm_set_exchange(S.LH_EXCHANGE_NCCL if args.exchange == "nccl" else S.LH_EXCHANGE_PEER)
        sortlast.apply(ctx, n, rank, world, args.shadow_split, args.partition)
    if emulate:
        sortlast.apply(ctx, n, emulate[0], emulate[1], args.shadow_split, args.partition)
    u = ctx.uniforms()
    visible_fraction = None
    if args.visible > 0:
        # culling-ratio sweep (BASELINE config 5): the camera moves along its viewing axis until the instance
        # cull keeps the requested share of the instances (bisection on the GPU's own count, a few untimed frames)
        cam = S.LhCamera.from_buffer_copy(bytes(scene.camera))
        z0 = float(cam.position[2])
        lo_z, hi_z = z0 * 6.0, -z0 * 0.9          # far behind the scene ... well inside it (z0 < 0: the camera looks at +z)
        for _ in range(14):
            mid = 0.5 * (lo_z + hi_z)
            cam.position[2] = mid
            u = ctx.uniforms(camera=cam)
            ctx.frame_begin(u); ctx.reset(); ctx.cull_instances(); ctx.sync()
            frac = ctx.sector(0)[0] / max(1, n)
            if frac > args.visible:
                lo_z = mid      # too many visible: move in
            else:
                hi_z = mid
        if dist:                # every rank must use the same camera
            z = torch.tensor([float(cam.position[2])], device="cuda", dtype=torch.float64)
            dist.broadcast(z, 0)
            cam.position[2] = float(z.item())
            u = ctx.uniforms(camera=cam)
        ctx.frame_begin(u); ctx.reset(); ctx.cull_instances(); ctx.sync()
        visible_fraction = ctx.sector(0)[0] / max(1, n) if world == 1 else args.visible

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()

    def frame_resident():
        ctx.frame_begin(u)
        ctx.render_frame()

    # ---- sort-last load balancing (untimed): the stripe bounds follow the ranks' measured pass times ----
    balance_log = None
    if world > 1 and args.partition == "stripes" and args.balance > 0:
        bal = sortlast.StripeBalancer(world)
        # ... and the row slices the ranks composite and resolve follow their measured resolve times
        rows = sortlast.RowBalancer(scene.height, world) if args.exchange == "peer" and not args.equal_rows else None
        shadow_striped = args.shadow_split == "instances"
        bal.apply(ctx, rank, views=(0, 1) if shadow_striped else (0,))
        balance_log = []
        for _ in range(args.balance):
            ctx.flush_l2()
            frame_resident()
            ctx.sync()
            t = ctx.timings()
            # (the composite's time includes its barrier, i.e. the other ranks' raster imbalance: the rows follow the resolve alone)
            mine = torch.tensor([t["cull_meshlets"] + t["visbuffer"], t["shadowmaps"], t["render_geometry"]], device="cuda", dtype=torch.float64)
            allc = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allc, mine)
            costs = [[float(c[v]) for c in allc] for v in range(3)]
            balance_log.append([round(bal.imbalance(costs[0]), 3), round(bal.imbalance(costs[1]), 3), round(bal.imbalance(costs[2]), 3)])
            bal.update(0, costs[0])
            if shadow_striped:
                bal.update(1, costs[1])
            bal.apply(ctx, rank, views=(0, 1) if shadow_striped else (0,))
            if rows:
                rows.update(costs[2])
                rows.apply(ctx, rank)

    # ---- warm-up (>= 3) ----
    warm = max(3, args.warmup)
    for _ in range(warm):
        ctx.flush_l2()
        frame_resident()
    ctx.sync()
    stats = ctx.stats()
    launches_before = ctx.launch_count()

    # ---- timed: device-resident frames, CUDA events on the context stream, L2 flushed between frames ----
    sampler = NvmlSampler(local_rank)
    if not sampler.ok:
        sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    frame_ms, pass_ms = [], []
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.flush_l2()
        ctx.timer_start()
        frame_resident()
        frame_ms.append(ctx.timer_stop())
        pass_ms.append(ctx.timings())
    barrier()
    wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    launches = (ctx.launch_count() - launches_before - args.steps) // max(1, args.steps)  # minus the L2 flush
    total_ms = float(np.sum(frame_ms))
    if dist:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = 1e3 / ms_per_step
    # per-rank medians of every pass (min / max over ranks name the limiter of the scaling curve)
    my_med = {k: float(np.median([p[k] for p in pass_ms])) for k in pass_ms[0]}
    rank_spread = None
    if dist:
        keys = sorted(my_med)
        mine = torch.tensor([my_med[k] for k in keys], device="cuda", dtype=torch.float64)
        allm = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allm, mine)
        m = torch.stack(allm).cpu().numpy()
        rank_spread = {k: {"min": round(float(m[:, i].min()), 4), "max": round(float(m[:, i].max()), 4)} for i, k in enumerate(keys) if m[:, i].max() > 0.002}

    # ---- e2e: the public host call, host buffers in, stats out, copies inside the timed region ----
    for _ in range(2):
        ctx.render_frame_host(u)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.render_frame_host(u)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if dist:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = args.steps / e2e_s
    h2d = C.sizeof(S.LhUniforms) + n * C.sizeof(S.LhInstance)
    d2h = C.sizeof(S.LhFrameStats)

    # ---- one untimed frame with the rasteriser's sample counter on (one 8-byte reduction per covered sample) ----
    flags = ctx.cfg.flags
    ctx.set_config_flags(flags | S.LH_CONFIG_COUNT_SAMPLES)
    frame_resident()
    ctx.sync()
    samples = [int(x) for x in ctx.stats().emitted_samples]
    ctx.set_config_flags(flags)

    if rank != 0:
        ctx.close()
        if dist:
            dist.destroy_process_group()
        return

    # ---- roofline: per-pass algorithmic bytes (SURVEY.md 8d / DESIGN.md) over the event-timed pass durations ----
    peak, peak_src = measured_peak()
    W, H = scene.width, scene.height
    px = W * H
    med = {k: float(np.median([p[k] for p in pass_ms])) for k in pass_ms[0]}
    inst_here = end - first
    avg_tris = scene.total_triangles / max(1, scene.total_meshlets)
    survivors = stats.num_opaque + stats.num_alpha_clip
    # meshlets that reach the vertex phase of the main view's raster kernel: survivors of the cull minus the early-outs
    rasterised = max(0, survivors - int(stats.meshlets_skipped_no_sample[0]) - int(stats.meshlets_skipped_occluded[0]))
    y0, y1 = ctx.comm_rows()           # rows rank 0 composites and resolves (the whole frame on one GPU)
    px_here = (y1 - y0) * W
    peer = n_gpus > 1 and args.exchange == "peer"
    alg = {
        "cull_instances": 112 * inst_here + 8 * stats.num_visible_instances,
        # 56 B per candidate; per survivor the 8 B reference, the 16 B draw command, and the rasteriser's 16 B bound + 4 B key
        "cull_meshlets": 56 * stats.num_candidate_meshlets + (8 + 16 + 16 + 4) * survivors + 8 * stats.num_visible_instances,
        # peer exchange: the composite kernel reads this rank's rows of every rank's keys and writes them once, with
        # the depth bounds folded in (the depth pass is then an 8-byte all-reduce)
        "composite": 8 * px_here * (n_gpus + 1) if peer else (2 * 8 * px * (n_gpus - 1) // n_gpus if n_gpus > 1 else 0),
        "depth_reduction": 0 if peer else 8 * px,
        # 8 B visibility word in + 8 B RGBA16F out per pixel (over the part of the pass on the critical path)
        "render_geometry": 16 * px_here,
        # raster geometry fetch per rasterised meshlet: 56 + 3 tri + 2 idx + 6 pos per vertex + 64 B of instance transform
        "raster_meshlets_kernel": rasterised * (56 + 3 * avg_tris + (2 + 6) * 64 + 64),
    }
    passes = {}
    for k, b in alg.items():
        ms = med.get(k, 0.0)
        if ms > 0 and b > 0:
            passes[k] = {"ms": round(ms, 4), "alg_bytes": int(b), "gbs": round(b / ms / 1e6, 1), "frac": round(b / ms / 1e6 / peak, 4)}
    kernel_share = {k: round(v / med["frame"], 4) for k, v in med.items() if k not in ("frame", "raster_meshlets_kernel_launches") and med["frame"] > 0}
    # ---- roofline of the dominant kernel (the rasteriser is bound by L2 atomics / instruction issue, not HBM) ----
    # raster_meshlets_kernel<cascades> is >= half of the GPU's work (the four cascades run on four streams; its time base
    # is the shadowmaps span, which also holds the cascades' cull / selection launches, so the fraction is a lower bound).
    # Issue roofline: warp instructions of the kernel per frame (ncu smsp__inst_executed.sum of the same command,
    # profiles/raster_ceilings.json) over that span, against 148 SMs x 4 schedulers x SM clock.  Atomic roofline: the
    # reductions counted live (one per depth-tested sample) over the live kernel time, against the measured ceiling of
    # tools/atomic_probe.cu on the same GPU type (profiles/r2a_atomic_probe.txt).
    ceil = {}
    try:
        with open(os.path.join(ROOT, "profiles", "raster_ceilings.json")) as f:
            ceil = json.load(f)
    except Exception:
        pass
    n_launch = max(1, int(round(med.get("raster_meshlets_kernel_launches", 1))))
    main_ms = med.get("raster_meshlets_kernel", 0.0)
    shadow_ms = med.get("shadowmaps", 0.0)
    sm_mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0
    issue_peak = 148 * 4 * sm_mhz / 1e3   # G warp instructions / s
    inst = ceil.get("raster_meshlets_kernel_warp_instructions_per_frame", {}) if (args.workload == "c2" and n_gpus == 1) else {}
    atomic_peak = ceil.get("l2_atomic_greductions_per_s", {}).get("8k_265MB_8x4_block_local" if px * 8 > 100e6 else "4k_66MB_8x4_block_local")
    traffic = None
    try:   # per-launch DRAM bytes of the same kernel from the committed ncu capture (profiles/)
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic = json.load(f).get("raster_meshlets_kernel<cascades>", {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    cascade_inst = inst.get("cascades")
    achieved_issue = cascade_inst / shadow_ms / 1e6 if (cascade_inst and shadow_ms > 0) else None
    atom_main = samples[0] / main_ms / 1e6 if main_ms > 0 else None
    atom_casc = sum(samples[1:]) / shadow_ms / 1e6 if shadow_ms > 0 else None
    roofline = {
        "bound": "issue", "kernel": "raster_meshlets_kernel<cascades> (one warp per meshlet; 24 launches on 4 streams)",
        "launches_per_step": 24, "avg_launch_ms": round(shadow_ms / 24, 4),
        "achieved": round(achieved_issue, 1) if achieved_issue else None, "peak": round(issue_peak, 1), "unit": "G warp-instructions/s",
        "frac": round(achieved_issue / issue_peak, 4) if achieved_issue else None,
        "traffic": traffic, "peak_source": "148 SMs x 4 warp schedulers x measured SM clock; instruction count from profiles/raster_ceilings.json (ncu, same command)",
        "l2_atomic": {
            "kernel": "raster_meshlets_kernel<main view>", "launches_per_step": n_launch, "kernel_ms": round(main_ms, 4),
            "reductions_per_step": samples[0], "achieved": round(atom_main, 2) if atom_main else None, "peak": atomic_peak,
            "unit": "G reductions/s", "frac": round(atom_main / atomic_peak, 4) if (atom_main and atomic_peak) else None,
            "cascades": {"reductions_per_step": sum(samples[1:]), "achieved": round(atom_casc, 2) if atom_casc else None,
                         "frac": round(atom_casc / atomic_peak, 4) if (atom_casc and atomic_peak) else None},
            "peak_source": "tools/atomic_probe.cu, 64-bit min reductions into a key buffer of this footprint, 8x4-block-local (profiles/r2a_atomic_probe.txt)"},
        "note": ("the rasteriser's working set lives in L2 (DRAM traffic per launch is a few MB): its ceilings are instruction issue and "
                 "L2 reductions, not HBM; the HBM-bound passes (culls, depth reduction, resolve) are under 'passes' against the measured copy peak"),
    }
    cpu = None
    if n_gpus == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(scene, tex, sky, budget_s=args.cpu_budget)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload], "resolution": [W, H], "instances": n, "input_triangles": scene.total_triangles,
                   "candidate_meshlets": scene.total_meshlets, "l2": "flushed between timed frames (256 MB write)",
                   "parallelism": (f"sort-last x{n_gpus}: instances split by {args.partition}, {args.exchange} exchange, shadows split by {args.shadow_split}"
                                   if n_gpus > 1 else "single GPU"), "texture_format": "BC7_SRGB",
                   "texture_size": args.tex_size, "texture_sets": scene.num_texture_sets, "meshlet_builder": meshlet_builder(),
                   "visible_instance_fraction": (round(visible_fraction, 4) if visible_fraction is not None else None),
                   **({"emulated_rank": args.emulate_rank, "note": "profiling aid: one rank's local work, NOT a bench value"} if emulate else {})},
        "culled_mtri_per_s": scene.total_triangles / (ms_per_step / 1e3) / 1e6,
        "frame": {"visible_instances": stats.num_visible_instances, "candidate_meshlets": stats.num_candidate_meshlets,
                  "rasterised_meshlets": survivors, "deferred_triangle_tiles": stats.num_large_triangles,
                  "rasterised_triangles_nonempty": int(stats.rasterized_triangles),
                  "meshlets_skipped_no_sample": list(stats.meshlets_skipped_no_sample),
                  "meshlets_skipped_occluded": list(stats.meshlets_skipped_occluded),
                  "meshlets_queued": list(stats.meshlets_queued), "emitted_samples": samples,
                  "shadow_rasterised_meshlets": [int(a) + int(b) for a, b in zip(stats.shadow_opaque, stats.shadow_alpha_clip)]},
        "pass_ms": {k: round(v, 4) for k, v in med.items()},
        "pass_notes": ("render_geometry is the resolve's shadow phase (critical path); its material phase (resolve_kernel<1>) runs on a "
                       "low-priority stream beside the depth and shadow passes; shadowmaps is the span of the four cascades on four "
                       "concurrent streams; raster_meshlets_kernel sums the main view's launches of that kernel"),
        "pass_share": kernel_share,
        "pass_ms_over_ranks": rank_spread,
        "stripe_balance": ({"iterations": args.balance, "imbalance_main_shadow_per_iteration": balance_log} if balance_log is not None else None),
        "passes": passes,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "call": "lh_render_frame_host (C ABI)"},
        "gpu_launches": int(launches * args.steps),
        "gpu_launches_per_frame": int(launches),
        "clocks": clocks,
        "wall_s_timed_region": round(wall, 3),
    }
    print(json.dumps(line), flush=True)
    ctx.close()
    if dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--tex-size", type=int, default=1024, help="SURVEY 8d: 1024^2 BC7 textures with full mip chains")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"], help="multi-GPU: peer-memory kernels or whole-buffer NCCL")
    ap.add_argument("--shadow-split", default="instances", choices=["instances", "cascades"])
    ap.add_argument("--partition", default="stripes", choices=["stripes", "range"], help="multi-GPU: how instances are shared out")
    ap.add_argument("--visible", type=float, default=0.0, help="culling-ratio sweep: move the camera until this share of the instances survives the instance cull (0: the preset's camera)")
    ap.add_argument("--balance", type=int, default=8, help="multi-GPU, stripes: untimed frames that move the stripe bounds towards equal per-rank pass times (0: equal-width interleaved stripes)")
    ap.add_argument("--equal-rows", action="store_true", help="multi-GPU: keep the even row slices (no row balancing)")
    ap.add_argument("--emulate-rank", default="", help="profiling aid, one GPU: do the local work of rank R of N, given as R/N (no exchange)")
    ap.add_argument("--stripes-per-rank", type=int, default=0, help="multi-GPU: interleaved screen stripes per GPU (0: sortlast.STRIPES_PER_RANK)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
