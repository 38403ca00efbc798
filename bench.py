#!/usr/bin/env python
"""bench.py — headline benchmark of the volume ray-march hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1], the one the metric is quoted on): synthetic BraTS-shaped
4-modality 240x240x155 fp32 volume, 1024x1024 perspective (fov 70 deg), step 0.5 voxel,
256-entry LUT transfer function, early ray termination at T <= 0.01 and occupancy-brick
empty-space skipping.  One STEP = one orbit batch of `--views` frames (default 8,
theta_k = 25 deg + k*360/V, phi = 80 deg).

Metric: ray samples/s — the number of sample slots the ORACLE's definition of the frame
evaluates (sum over rays of ceil((t1-t0)/dt) truncated by early termination; counted once by an
untimed counting pass, so skipping cannot inflate it) divided by device time.  frames/s rides
along.  `roofline` is computed from the samples the march kernel ACTUALLY evaluates (trilinear
fetches really issued), see DESIGN.md §measurement.

Prints ONE JSON line (rank 0).  Timing: CUDA events on the launching stream, barrier +
synchronize on both sides, max over ranks, L2 flushed between timed steps.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

DIMS = (240, 240, 155)
NCH = 4
IMG = 1024
TF_N = 256
METRIC = "ray_samples_per_sec"
UNIT = "samples/s"
WORKLOAD = ("cfg2: synthetic BraTS 4-modality 240x240x155 fp32, 1024x1024 perspective fov70, step 0.5 voxel, "
            "256-entry LUT TF, ERT 0.01, occupancy-brick skipping")


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _scene(views: int, theta0_deg: float = 25.0):
    import numpy as np
    from mri_raytracer_b200 import Camera, OrbitalCamera, RenderParams
    from mri_raytracer_b200.synth import world_box
    vs, vmin = world_box(DIMS)
    ext = vs * np.asarray(DIMS, dtype=np.float32)
    cam = OrbitalCamera(initial_radius=3.0, initial_theta=math.radians(theta0_deg), initial_phi=math.radians(80.0))
    cam.set_fov_degrees(70.0)
    cam.target = (vmin + 0.5 * ext).astype(np.float32)
    cam.radius = float(np.linalg.norm(ext) * 0.8)            # frame_volume, brats_viewer.py:320-324
    cams = []
    th0 = cam.theta
    for k in range(views):
        cam.theta = th0 + 2.0 * math.pi * k / views
        cams.append(Camera.from_orbital(cam))
    P = RenderParams(imageSize=(IMG, IMG), dims=DIMS, voxelSize=tuple(float(v) for v in vs),
                     volMin=tuple(float(v) for v in vmin), stepSize=float(np.float32(0.5) * vs[0]),
                     skipEmpty=1, tfMode=1)
    return P, cams


def _config(views: int, world: int):
    """The workload both arms declare (identical dict): one STEP = `views` full 1024x1024 frames of the
    orbit theta_k = 25 deg + k*360/(views*world) per GPU."""
    return {"workload": WORKLOAD, "views_per_step_per_gpu": views, "views_per_step_total": views * world,
            "image": f"{IMG}x{IMG}", "volume": f"{NCH}x{DIMS[2]}x{DIMS[1]}x{DIMS[0]} fp32", "tf_entries": TF_N,
            "orbit": "theta_k = 25 deg + k*360/views_per_step_total, phi = 80 deg, fov 70 deg, radius 0.8*|extent|",
            "l2": "flushed (256 MiB write) between timed steps; volume 142.8 MB > 126 MB L2"}


class NvmlSampler:
    """SM clock / throttle reasons DURING the timed region, polled through NVML every ~2 ms (the
    timed region of a default run is tens of ms: too short for `nvidia-smi -lms`)."""

    def __init__(self, index: int):
        self.index, self.rows, self.stop_flag, self.thread, self.ok = index, [], False, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.ok = True
        except Exception:
            self.ok = False

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((sm, rs))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.ok:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()

    def stop(self):
        if not self.ok:
            return None
        self.stop_flag = True
        self.thread.join(timeout=1.0)
        nv = self.nv
        sm = sorted(r[0] for r in self.rows)
        bits = 0
        for _, r in self.rows:
            bits |= r
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80}
        reasons = sorted(k for k, m in names.items() if bits & m)
        try:
            smax = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:
            smax = None
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": reasons,
                "samples": len(sm), "source": "NVML polled every 2 ms during the timed region"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------ CPU arm
def cpu_reference_sample(views, threads: int, kind: str = "c", stride: int = 1, total_views: int = 8):
    """The reference CPU path.  The reference ships no CPU renderer (its arithmetic exists only as
    Slang shaders), so this is the oracle port: `kind="c"` = oracle/oracle_c.c (scalar C, POSIX
    threads over all host cores — the faster of the two, hence the baseline we quote);
    `kind="torch"` = oracle/oracle_torch.py (BASELINE.md section 4's CPU-torch path).  Bounded
    sample: views `views` (indices into the step's orbit of `total_views`), every `stride`-th pixel in
    x and y.  Returns (callable -> samples taken, text)."""
    import torch
    from mri_raytracer_b200.synth import make_brats_like, ramp_tf
    vol = make_brats_like(NCH, DIMS, seed=0)
    tf = ramp_tf(TF_N)
    P, cams = _scene(total_views)
    ys, xs = torch.meshgrid(torch.arange(0, IMG, stride), torch.arange(0, IMG, stride), indexing="ij")
    px, py = xs.reshape(-1), ys.reshape(-1)
    views = list(views)
    what = (f"views {views} of the {total_views}-view step, " + ("every pixel" if stride == 1 else f"every {stride}th pixel in x and y")
            + f" ({px.numel() * len(views)} of {IMG * IMG * total_views} rays)")
    if kind == "c":
        from oracle import oracle_c
        voln, tfn, pxn, pyn = vol.numpy(), tf.numpy(), px.numpy(), py.numpy()

        def run():
            tot = 0
            for v in views:
                _, aux = oracle_c.render(voln, P.with_camera(cams[v]), tf=tfn, pixels=(pxn, pyn), return_aux=True, threads=threads)
                tot += int(aux["n_taken"].sum())
            return tot
        return run, what + f", scalar C oracle, {threads} POSIX threads, fp32"
    from oracle import oracle_torch as O
    torch.set_num_threads(threads)

    def run():
        tot = 0
        for v in views:
            _, aux = O.render(vol, P.with_camera(cams[v]), tf=tf, pixels=(px, py), return_aux=True)
            tot += int(aux["n_taken"].sum())
        return tot
    return run, what + f", CPU-torch oracle, {threads} threads, fp32"


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path (the oracle port: the
    reference has no CPU renderer and no C sources to compile) on all host threads, on the SAME
    config as the GPU arm.  One step = the step's own views at full resolution (every ray the GPU arm
    counts at N=1); `--ref-views` bounds it (default: all of them)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    V = args.views
    nv = V if args.ref_views <= 0 else min(V, args.ref_views)
    run, sample = cpu_reference_sample(range(nv), threads=threads, kind="c", total_views=V * args.gpus)
    for _ in range(max(args.warmup, 1)):
        run()
    step_s, tot = [], 0
    for _ in range(args.steps):
        t0 = time.perf_counter()
        tot += run()
        step_s.append(time.perf_counter() - t0)
    dt = sum(step_s)
    val = tot / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": _config(V, args.gpus),
        "frames_per_sec": nv * args.steps / dt,
        "step_ms": [1e3 * x for x in step_s],
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ GPU arm
def _profile_numbers():
    """Counters of the dominant kernel from the committed ncu digest of this round (profiles/)."""
    out = {}
    p = ROOT / "profiles" / "r02_fwd_batch8.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())[0]
            out["issue_active_pct"] = float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"].split()[0])
            out["l1_data_stage_wavefronts_pct"] = float(d["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"].split()[0])
            out["warp_instructions"] = float(d["smsp__inst_executed.sum"].split()[0])
            out["source"] = "profiles/r02_fwd_batch8.json (ncu --set full of the same batched launch)"
        except Exception:
            pass
    return out


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from mri_raytracer_b200 import api, build, tiles
    from mri_raytracer_b200 import dist as mdist
    from mri_raytracer_b200.synth import make_brats_like, ramp_tf

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the render path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    V = args.views                      # views per GPU per step (weak scaling: per-GPU work is fixed)
    VT = V * world
    P, cams_all = _scene(VT)
    mine = cams_all[rank * V:(rank + 1) * V]
    vol_host = make_brats_like(NCH, DIMS, seed=0).pin_memory()
    tf_host = ramp_tf(TF_N)
    vol = vol_host.to(dev, non_blocking=True)
    tf = tf_host.to(dev)
    volume = api.Volume(vol, fold=not args.no_fold)
    W = H = IMG
    fb = mdist.PeerFramebuffer(VT, H, W, dev, owners=args.owners, partition=args.partition) if world > 1 else None

    # ---- untimed counting pass: the oracle-defined sample count (each rank counts a block of views)
    taken = evaluated = clip = 0
    per_view_eval = []
    for c in mine:
        _, _, counts = api.render_aux(volume, c, tf, P)
        s = counts.sum(dim=(0, 1)).tolist()
        clip += s[0]; taken += s[1]; evaluated += s[2]
        per_view_eval.append(s[2])
    torch.cuda.synchronize()
    if world > 1:                               # whole-job totals
        tot = torch.tensor([taken, evaluated, clip], dtype=torch.float64, device=dev)
        dist.all_reduce(tot)
        taken, evaluated, clip = (int(x) for x in tot.tolist())

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    frames = torch.empty((V, H, W, 4), dtype=torch.float32, device=dev)
    kern_ev = []
    ev_pool = []                                 # timing events for the march, created (recorded once) outside the timed steps
    for _ in range(args.steps):
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record(); b_.record()
        ev_pool.append((a_, b_))
    last = [None]

    def step(record_kernels: bool):
        """One orbit batch through the public API; frames land in `frames` (N=1) or in the owners'
        peer-mapped framebuffers (N>1)."""
        volume.invalidate()      # every step re-folds the modalities + rebuilds the occupancy grid
        if world == 1 and not (args.per_view or args.no_fold):
            # the public call: the volume is stale, so fold + occupancy + layout, classify, spans and ONE march
            # launch (grid.y = view) are queued by one library call (mrt_render_views_refold); the two events
            # are re-recorded by the library around the spans + march launches
            evs = None
            if record_kernels:
                evs = ev_pool[len(kern_ev)]
                kern_ev.append(evs)
            api.render_views(volume, mine, tf, P, out=frames, march_events=evs)
        elif world == 1:
            Pv = P.with_camera(mine[0])
            packed, Ce, Pe = volume.prepared(Pv)                      # fold + occupancy build
            bits = volume.skip_levels(Pv, tf)                         # classify (camera independent)
            if record_kernels:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
            if args.per_view:
                for v, c in enumerate(mine):
                    api.render_forward(Pe.with_camera(c), packed, Ce, tf, bits, out=frames[v])
            else:
                volume.march_batch(Pe, mine, packed, Ce, tf, bits, out=frames)   # spans (tiny) + ONE march launch, grid.y = view
            if record_kernels:
                b.record(); kern_ev.append((a, b))
        else:
            fb.render(volume, cams_all, tf, P)      # this rank's tile rows of EVERY view, stored into the owners' frames over NVLink
            last[0] = fb.finish()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, nsteps):
        ms = []
        for _ in range(nsteps):
            flush.fill_(1)                                            # L2 flush, outside the timed events
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            barrier()
            ms.append(a.elapsed_time(b))
        t = torch.tensor(ms, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)                  # per step: the slowest rank
        return t.tolist()

    for _ in range(max(args.warmup, 3)):
        step(False)
    barrier()

    sampler = NvmlSampler(local)
    if not sampler.ok:
        sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    step_ms = timed(lambda: step(True), args.steps)
    clocks = sampler.stop() if rank == 0 else None
    tot_s = sum(step_ms) / 1e3
    value = taken * args.steps / tot_s
    med_ms = sorted(step_ms)[len(step_ms) // 2]

    # ---- N > 1: verify the gathered frames (outside the timed region), and strong scaling of ONE 8-view batch
    verified = None
    strong = None
    if world > 1:
        own = fb.owned_views()
        okl = True
        for v in sorted({own.start, own.stop - 1}) if len(own) else []:
            okl &= bool(torch.equal(last[0][v - own.start], api.render(volume, cams_all[v], tf, P)))
        okt = torch.tensor([1 if okl else 0], device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        verified = bool(okt.item())
        Ps, cams_s = _scene(V)                                         # the N=1 step: V views, FIXED total work
        fbs = mdist.PeerFramebuffer(V, H, W, dev, owners=args.owners, partition="tiles")

        def strong_step():
            volume.invalidate()
            fbs.render(volume, cams_s, tf, Ps)
            last[0] = fbs.finish()

        def single_step():                                            # the same batch on this GPU alone
            volume.invalidate()
            api.render_views(volume, cams_s, tf, Ps, out=frames)
        for _ in range(3):
            strong_step(); single_step()
        ms_n = timed(strong_step, args.steps)
        barrier()
        ms_1 = []
        for _ in range(args.steps):
            flush.fill_(1); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); single_step(); b.record(); torch.cuda.synchronize()
            ms_1.append(a.elapsed_time(b))
        own_s = fbs.owned_views()
        oks = True
        for v in own_s:
            oks &= bool(torch.equal(last[0][v - own_s.start], api.render(volume, cams_s[v], tf, Ps)))
        okt = torch.tensor([1 if oks else 0], device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        m1, mn = sorted(ms_1)[len(ms_1) // 2], sorted(ms_n)[len(ms_n) // 2]
        # the same batch with a STATIC volume (the reference's frame loop: the volume is uploaded once,
        # brats_viewer.py:219-230, and only the camera moves): the fold / occupancy / layout are cached,
        # a step is classify + spans + march
        def strong_static():
            fbs.render(volume, cams_s, tf, Ps)
            last[0] = fbs.finish()

        def single_static():
            api.render_views(volume, cams_s, tf, Ps, out=frames)
        for _ in range(3):
            strong_static(); single_static()
        ms_ns = timed(strong_static, args.steps)
        barrier()
        ms_1s = []
        for _ in range(args.steps):
            flush.fill_(1); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); single_static(); b.record(); torch.cuda.synchronize()
            ms_1s.append(a.elapsed_time(b))
        m1s, mns = sorted(ms_1s)[len(ms_1s) // 2], sorted(ms_ns)[len(ms_ns) // 2]
        strong = {"what": f"ONE {V}-view cfg2 batch (the N=1 step) split over {world} GPUs by interleaved tile rows, frames striped over the owners",
                  "ms_per_step_1gpu_same_run": m1, "ms_per_step": mn, "speedup": m1 / mn, "efficiency": m1 / mn / world,
                  "frames_verified": bool(okt.item()), "step_ms": ms_n,
                  "static_volume": {"what": "the same batch with the fold / occupancy / sampler layout cached (volume static, camera moving)",
                                    "ms_per_step_1gpu_same_run": m1s, "ms_per_step": mns, "speedup": m1s / mns,
                                    "efficiency": m1s / mns / world},
                  "limiter": "the modality fold + occupancy + quad layout (one pass over the 143 MB planar volume, ~0.06 ms) and "
                             "classify + spans are replicated on every rank, plus one symmetric-memory barrier and ~6 launches; only "
                             "the march (~0.56 of ~0.66 ms at N=1) divides by N.  Sharding the fold would need an all-gather of the "
                             "folded volume that costs as much as folding it locally"}
        del fbs

    # ---- the same step replayed as ONE CUDA graph (N = 1): what the step costs without the host launch path.
    # Not the headline: the cameras are kernel parameters, so a replay renders the captured views again.
    graph_rec = None
    if world == 1 and not args.per_view:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step(False)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                step(False)
            g.replay(); torch.cuda.synchronize()
            ref_frames = frames.clone()
            gms = timed(g.replay, args.steps)
            step(False); torch.cuda.synchronize()
            graph_rec = {"ms_per_step": sum(gms) / len(gms), "ms_per_step_median": sorted(gms)[len(gms) // 2],
                         "frames_equal_eager": bool(torch.equal(ref_frames, frames)),
                         "what": "the timed step (fold + occupancy + layout, classify, spans, one batched march) captured once and "
                                 "replayed with the same L2 flush and synchronisation between replays; eager - graph = the host "
                                 "launch path (Python + ctypes + 4 launches) exposed by the per-step synchronisation"}
        except Exception as e:                  # reported, never hidden
            graph_rec = {"error": f"{type(e).__name__}: {e}"}

    # ---- roofline of the dominant kernel (march), from live CUDA-event launch durations
    roof = None
    if kern_ev:
        durs = [a.elapsed_time(b) for a, b in kern_ev]
        avg_ms = sum(durs) / len(durs)                                # one launch = the whole batch of V views
        avg_eval = float(sum(per_view_eval))
        kch = 1 if volume.fold else NCH                               # channels the march kernel gathers
        bytes_per_sample = 32 * kch                                   # 8 corners x 4 B x C
        achieved = avg_eval * bytes_per_sample / (avg_ms * 1e-3) / 1e9
        peak, which = _peaks()
        traffic = None
        tp = ROOT / "profiles" / "traffic.json"
        if tp.exists():
            traffic = json.loads(tp.read_text()).get("mrt_fwd_kernel_dram_bytes_per_launch")
        prof = _profile_numbers()
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": which,
                "limiter": "issue slots + L1 data-stage wavefronts (NOT HBM: the folded volume is L1/L2-resident, DRAM traffic ~2 % of peak)",
                "kernel": f"mrt_fwd_kernel<{kch},false,true,false>",
                "avg_launch_ms": avg_ms, "bytes_per_sample": bytes_per_sample,
                "achieved_at_survey_128B_per_sample": avg_eval * 32 * NCH / (avg_ms * 1e-3) / 1e9,
                "evaluated_samples_per_launch": avg_eval,
                "evaluated_samples_per_sec": avg_eval / (avg_ms * 1e-3),
                "nominal_samples_per_launch": taken, "views_per_launch": 1 if args.per_view else V,
                "issue_active_pct": prof.get("issue_active_pct"),
                "l1_data_stage_wavefronts_pct": prof.get("l1_data_stage_wavefronts_pct"),
                "instr_per_slot": (prof["warp_instructions"] / (avg_eval / 32.0)) if "warp_instructions" in prof else None,
                "counters_source": prof.get("source"),
                "note": "`bound` keeps the bench contract's two-valued field (memory-side vs tensor-side roofline); `limiter` names "
                        "what actually binds. achieved = bytes the march kernel's own gathers request (8 corners x 4 B x channels "
                        "it reads; 1 channel after the modality fold) x samples whose fetches were really issued / CUDA-event "
                        "launch time; `frac` is that over the HBM copy peak and is a reference line only. `value` counts the "
                        "oracle-defined (nominal) slots, which include slots skipped as provably empty: the evaluated rate is "
                        "evaluated_samples_per_sec. instr_per_slot = warp instructions per 32 evaluated lane-slots"}
        if traffic:
            # what really crosses the HBM interface (ncu dram__bytes of the same launch / this run's launch time): `frac`
            # above can exceed 1 because the requested bytes are served from L1 (86 % hit) and L2, not from DRAM
            roof["dram_gbs"] = float(traffic) / (avg_ms * 1e-3) / 1e9
            roof["dram_frac"] = roof["dram_gbs"] / peak

    # ---- same-run gather ceilings: random 32-byte-sector gathers over an
    # L2-resident (32 MiB) and an HBM-resident (4 GiB) buffer, mrt_gather_probe
    if roof is not None and not args.no_probe:
        from mri_raytracer_b200._lib import lib, check
        chk = torch.zeros(2, dtype=torch.float32, device=dev)

        def ceiling(nbytes, n_gathers):
            buf = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
            st = torch.cuda.current_stream().cuda_stream
            per_thread = (n_gathers // (148 * 8 * 256) + 7) // 8 * 8
            real = per_thread * 148 * 8 * 256
            best = None
            for i in range(4):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                check(lib().mrt_gather_probe(buf.data_ptr(), nbytes, n_gathers, 17 + i, chk.data_ptr(), st), "gather_probe")
                b.record(); torch.cuda.synchronize()
                ms = a.elapsed_time(b)
                if i > 0:
                    best = ms if best is None else min(best, ms)
            del buf
            return real * 32 / (best * 1e-3) / 1e9
        l2c = ceiling(32 << 20, 1 << 27)
        hbc = ceiling(4 << 30, 1 << 27)
        roof["gather_ceiling"] = {
            "l2_resident_32MiB_gbs": l2c, "hbm_resident_4GiB_gbs": hbc, "unit": "GB/s of 32-byte sectors",
            "frac_of_l2_ceiling": roof["achieved"] / l2c,
            "note": "the folded volume (35.9 MB) is L2-resident and 96 % of its sectors hit in L1, so the march's "
                    "requested-byte rate may exceed the L2 random-gather ceiling"}

    # ---- e2e: host buffers in, host frames out, through the C ABI's host-buffer entry
    # (mrt_host_pipeline_*, api.HostPipeline).  The volume is RESIDENT, uploaded once before the
    # timed region exactly as the reference does at load time (brats_viewer.py:219-230); a step's
    # inputs — cameras, params incl. the modality weights, TF — go up from pinned host memory every
    # step, the fold + occupancy + classify + ONE batched march run, and the V frames come down
    # sparse (each view's bounding rectangle of non-background tiles; the pipeline keeps the rest of
    # the host frame at the background).  Steps are triple-buffered.  The round-1 variant that also
    # uploads the 143 MB volume every step is measured next to it.
    e2e = None
    if not args.no_e2e:
        vol_np = vol_host.numpy()
        tf_np = tf_host.numpy()
        from mri_raytracer_b200.hostmem import bind_to_gpu_numa
        numa = bind_to_gpu_numa(local)          # before the pinned frames are allocated (first touch)
        outs = [torch.empty((V, H, W, 4), dtype=torch.float32).pin_memory() for _ in range(3)]
        local_frames = None
        if world == 1:
            local_frames = frames.cpu()
        else:
            local_frames = api.render_views(volume, mine, tf, P).cpu()

        def run_pipe(resident: bool):
            pipe = api.HostPipeline(NCH, DIMS, (W, H), max_views=V, max_tf=TF_N, depth=3)
            if resident:
                pipe.set_volume(vol_np)
            vin = None if resident else vol_np
            for i in range(3):
                pipe.wait(pipe.submit(vin, mine, P, tf_np, outs[i % 3].numpy(), fresh=True))
                assert torch.equal(outs[i % 3], local_frames), "host pipeline frames differ from the device-side batch"
            up, down, _ = pipe.last_bytes()
            ks = max(6, min(args.steps, 20))
            barrier()
            t0 = time.perf_counter()
            tickets = [pipe.submit(vin, mine, P, tf_np, outs[i % 3].numpy()) for i in range(ks)]
            pipe.wait(tickets[-1])
            torch.cuda.synchronize()
            t_e2e = time.perf_counter() - t0
            lat = []
            for i in range(3):
                t0 = time.perf_counter()
                pipe.wait(pipe.submit(vin, mine, P, tf_np, outs[i % 3].numpy()))
                lat.append(time.perf_counter() - t0)
            assert torch.equal(outs[0], local_frames) and torch.equal(outs[2], local_frames)
            pipe.close()
            if world > 1:
                tt = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                t_e2e = float(tt)
            return {"value": taken * ks / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(up) * world,
                    "d2h_bytes_per_step": int(down) * world, "ms_per_step": 1e3 * t_e2e / ks,
                    "frames_per_sec": VT * ks / t_e2e, "steps": ks, "single_step_latency_ms": 1e3 * sorted(lat)[1]}
        e2e = run_pipe(True)
        e2e["numa_node_rank0"] = numa
        # what the host lets through: every rank copies the bytes its step downloads, as ONE contiguous
        # device -> pinned-host copy on its own PCIe link, all ranks at once (no kernels involved)
        nb = max(1, int(e2e["d2h_bytes_per_step"]) // world)
        dsrc = torch.empty(nb, dtype=torch.uint8, device=dev)
        hdst = outs[0].view(torch.uint8).reshape(-1)[:nb]
        for _ in range(2):
            hdst.copy_(dsrc, non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(10):
            hdst.copy_(dsrc, non_blocking=True)
        torch.cuda.synchronize()
        tcp = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([tcp], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            tcp = float(tt)
        e2e["d2h_ceiling"] = {"ms_per_step_copy_only": 1e3 * tcp / 10, "aggregate_gbs": nb * world * 10 / tcp / 1e9,
                              "what": "plain contiguous D2H copies of the same byte count from all ranks at once: the floor "
                                      "the host's PCIe / IOMMU sets for ms_per_step"}
        e2e["what"] = ("mrt_host_pipeline (C ABI, host buffers), one per GPU: volume resident (uploaded once, as the reference does "
                       "at load time); per step cameras + params + TF H2D from pinned host memory, modality fold + occupancy, "
                       "classify, spans, ONE batched march of V views whose in-span tiles are stored straight into the pinned "
                       "(device-mapped) host frames over PCIe — no staging copy; the host frame outside the spans is kept at the "
                       "background by damage tracking; frames verified bit-identical to the device-side batch; steps "
                       "triple-buffered, prepare stream at high priority; wall clock over all steps (max over ranks), "
                       "synchronize on both sides; bytes are whole-job totals")
        e2e["with_volume_upload_every_step"] = run_pipe(False)

    # ---- CPU baseline (rank 0, N = 1 only): the oracle on a bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        run, sample = cpu_reference_sample(range(V), threads=threads, kind="c", total_views=V)
        run_t, sample_t = cpu_reference_sample([0], threads=threads, kind="torch", stride=4, total_views=V)
        t0 = time.perf_counter()
        n = run()
        dt = time.perf_counter() - t0
        t0 = time.perf_counter()
        nt_ = run_t()
        dtt = time.perf_counter() - t0
        cpu = {"value": n / dt, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample, "seconds": dt,
               "torch_oracle": {"value": nt_ / dtt, "sample": sample_t, "seconds": dtt}}

    line = None
    if rank == 0:
        cfg = _config(V, world)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * tot_s / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": cfg,
            "partition": ("single GPU" if world == 1 else
                          f"image space, {fb.partition} (interleaved tile rows of every view per rank), volume replicated; frames "
                          f"owned {fb.owners} over the ranks and written by peer (NVLink) stores from inside the march kernel; tiles "
                          "outside the projected footprints of the active bricks are not sent but filled by the owner"),
            "ms_per_step_median": med_ms, "value_at_median_step": taken / (med_ms * 1e-3), "step_ms": step_ms,
            "frames_per_sec": VT * args.steps / tot_s,
            "samples_per_step": {"nominal_taken": taken, "clip": clip, "evaluated": evaluated},
            "gathered_frames_verified": verified, "strong": strong, "graph_replay": graph_rec,
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
            # our kernels inside the timed region, whole job: per rank fold+occupancy, classify, spans (init + the
            # per-brick union), march (+ the owners' background fill at N > 1)
            "gpu_launches": ((V + 2 if args.per_view else 5) + (1 if world > 1 else 0)) * world * args.steps,
        }
    # ---- the other BASELINE configs (outside every timed region above), attached to the same line
    if not args.no_configs:
        del frames, flush
        torch.cuda.empty_cache()
        other = _other_configs(rank, world, dev, line)
        if line is not None:
            line["other_configs"] = other
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _other_configs(rank, world, dev, line):
    """BASELINE.json's other configs, each measured by its own tool function (tools/bench_configs.py,
    tools/bench_cfg4.py, tools/bench_cfg5.py) and attached to the bench line so that they are on the
    driver's record: cfg1 (+ u8 storage), cfg3 (forward + backward, with the backward's roofline), INR
    inference (tensor roofline) and cfg4 at N = 1; cfg4 strong-scaled and cfg5 (brick-sharded fp16,
    sort-last over NVLink) at N > 1.  Not part of `value`: each record names its own unit.  A
    watchdog prints the headline line alone if this leg hangs (a peer that died inside a collective)."""
    import argparse as _ap
    import threading
    import torch
    sys.path.insert(0, str(ROOT / "tools")); sys.path.insert(0, str(ROOT / "tests"))

    def bail():
        if line is not None:
            line["other_configs"] = {"error": "watchdog: the secondary configs did not finish in time"}
            print(json.dumps(line), flush=True)
        os._exit(0)
    dog = threading.Timer(240.0, bail)
    dog.daemon = True
    dog.start()
    out = {}

    def attempt(name, fn):
        t0 = time.perf_counter()
        try:
            rec = fn()
            rec["wall_s"] = time.perf_counter() - t0
            out[name] = rec
        except Exception as e:                  # reported in the line, never hidden
            out[name] = {"error": f"{type(e).__name__}: {e}"}
        torch.cuda.empty_cache()

    if world == 1:
        import bench_configs as BC
        BC.WITH_ORACLE = False
        for name in ("cfg1", "cfg1_u8", "cfg3", "cfg4", "inr"):
            attempt(name, getattr(BC, name))
    else:
        import bench_cfg4
        import bench_cfg5
        # cfg4 as BASELINE states it: 64 views at 2048^2 over 512^3, strong-scaled over the ranks
        attempt("cfg4", lambda: bench_cfg4.run(_ap.Namespace(dim=512, img=2048, views=64, reps=5, partition="tiles", owners="striped"),
                                               rank, world, dev))
        # cfg5: 1024^3 fp16 voxels per GPU, the shard grid of the world size: 2048^3 (BASELINE's size) on 8 GPUs,
        # 1024 x 2048 x 2048 on 4, 1024 x 1024 x 2048 on 2 (2048^3 itself needs >= 4 shards: < 2^32 voxels per shard)
        from mri_raytracer_b200 import dist as mdist
        dims5 = tuple(1024 * g for g in mdist.shard_grid(world))
        attempt("cfg5", lambda: bench_cfg5.run(_ap.Namespace(dims=dims5, dim=0, img=4096, views=4, reps=3, emulate=0, nccl=False, check=False),
                                               rank, world, dev))
    dog.cancel()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--views", type=int, default=8, help="frames per step (orbit batch)")
    ap.add_argument("--partition", default="tiles", choices=["tiles", "views"], help="multi-GPU image-space partition")
    ap.add_argument("--owners", default="striped", choices=["striped", "root"], help="multi-GPU: where the frames live")
    ap.add_argument("--per-view", action="store_true", help="one march launch per view instead of one per batch")
    ap.add_argument("--ref-views", type=int, default=0, help="--impl reference: views per step (0 = all)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg")
    ap.add_argument("--no-probe", action="store_true", help="skip the gather-ceiling probe")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU-oracle baseline leg")
    ap.add_argument("--no-configs", action="store_true", help="skip the secondary configs attached as `other_configs`")
    ap.add_argument("--no-fold", action="store_true", help="blend modalities per sample (float4 gathers) instead of folding")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
