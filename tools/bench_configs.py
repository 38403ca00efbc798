"""Secondary measurements: the other BASELINE.json configs on one GPU (not the bench.py headline).

    python tools/bench_configs.py cfg1 cfg3 cfg4

cfg1: 1x240x240x155, 512^2 orthographic, step 0.5 voxel, 256-entry LUT (the reference's CPU-runnable case)
cfg3: differentiable rendering, forward+backward over a 256^3 volume and TF at 512^2
cfg4: orbit views at 2048^2 over a 512^3 volume (8 of the 64 views per timing batch on one GPU)
Each line: device time (CUDA events, pipelined launches), sample counts, and an oracle parity
spot check on a strided subset of pixels (forward) / a small-scene gradient check (cfg3).
"""
import json
import math
import sys
import time
from dataclasses import replace
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402

from mri_raytracer_b200 import api  # noqa: E402
from mri_raytracer_b200.synth import make_brats_like, ramp_tf  # noqa: E402
from scenes import framed_params  # noqa: E402


# bench.py attaches these records to its JSON line with WITH_ORACLE = False: the CPU oracle is then not
# imported at all (parity lives in tests/; here it is a spot check for the stand-alone tool)
WITH_ORACLE = True


def glorot_mlp(rng, in_dim, hidden, out_dim):
    """Glorot-uniform weights + small random biases in the reference's parameter layout
    (a list of {"W": [in,out], "b": [out]}, inr/inr/model.py:26-40)."""
    import numpy as np
    dims = [in_dim] + list(hidden) + [out_dim]
    params = []
    for a, b in zip(dims[:-1], dims[1:]):
        lim = math.sqrt(6.0 / (a + b))
        params.append({"W": rng.uniform(-lim, lim, size=(a, b)).astype(np.float32),
                       "b": rng.normal(scale=0.1, size=(b,)).astype(np.float32)})
    return params


def timeit(fn, n=5, warm=2, reps=4):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / reps)
    ts.sort()
    return ts[len(ts) // 2]


def parity_subset(vol_cpu, P, tf_cpu, img, stride):
    from oracle import oracle_c
    W, H = P.imageSize
    ys, xs = torch.meshgrid(torch.arange(0, H, stride), torch.arange(0, W, stride), indexing="ij")
    px, py = xs.reshape(-1).numpy(), ys.reshape(-1).numpy()
    ref = oracle_c.render(vol_cpu.numpy(), P, tf=None if tf_cpu is None else tf_cpu.numpy(), pixels=(px, py), threads=16)
    got = img.cpu().numpy()[py, px]
    d = abs(got - ref).max(axis=-1)
    return float(d.max()), int((d > 1e-4).sum()), int(px.size)


def cfg1():
    dims = (240, 240, 155)
    vol = make_brats_like(1, dims, seed=0)
    tf = ramp_tf(256)
    P = replace(framed_params(dims, 512, 512, ortho=True), tfMode=1)
    V = api.Volume(vol.cuda())
    tfd = tf.cuda()
    img, T, counts = api.render_aux(V, None, tfd, P)
    c = counts.sum(dim=(0, 1)).tolist()
    ms = timeit(lambda: api.render(V, None, tfd, P), reps=10)
    rec = dict(cfg="cfg1", ms_per_frame=ms, fps=1e3 / ms, samples_taken=c[1], samples_evaluated=c[2],
               gsamples_per_s=c[1] / ms / 1e6)
    if WITH_ORACLE:
        mx, nbad, npx = parity_subset(vol, P, tf, img, 4)
        rec.update(parity_max_abs=mx, parity_pixels_over_1e4=nbad, parity_pixels=npx)
    return rec


def cfg1_u8():
    """cfg1's shape with the volume stored as bytes (8 B gathered per sample instead of 32)."""
    dims = (240, 240, 155)
    vol = make_brats_like(1, dims, seed=0)
    tf = ramp_tf(256)
    P = replace(framed_params(dims, 512, 512, ortho=True), tfMode=1)
    u8 = (vol * 255.0).round().clamp(0, 255).to(torch.uint8)
    V8 = api.Volume(u8.cuda())
    V32 = api.Volume((u8.float() / 255.0).cuda())
    tfd = tf.cuda()
    _, _, counts = api.render_aux(V8, None, tfd, P)
    c = counts.sum(dim=(0, 1)).tolist()
    ms8 = timeit(lambda: api.render(V8, None, tfd, P), reps=10)
    ms32 = timeit(lambda: api.render(V32, None, tfd, P), reps=10)
    diff = float((api.render(V8, None, tfd, P) - api.render(V32, None, tfd, P)).abs().max())
    return dict(cfg="cfg1_u8", ms_per_frame_u8=ms8, ms_per_frame_fp32_same_values=ms32, samples_taken=c[1], samples_evaluated=c[2],
                gsamples_per_s_u8=c[1] / ms8 / 1e6, max_abs_u8_vs_fp32=diff)


def cfg3():
    dims = (256, 256, 256)
    vol = make_brats_like(1, dims, seed=4)
    tf = ramp_tf(256, sigma_scale=20.0, cutoff=0.05)
    P = replace(framed_params(dims, 512, 512), tfMode=1)
    with torch.no_grad():
        target = api.render(api.Volume(vol.cuda()), None, (tf * torch.tensor([0.8, 1.0, 1.1, 1.3])).cuda(), P)
    v = vol.cuda().requires_grad_(True)
    t = tf.cuda().requires_grad_(True)

    mse = torch.nn.functional.mse_loss                      # mean((img - target)^2), BASELINE cfg3's loss, as one fused op

    def step():
        v.grad = None; t.grad = None
        loss = mse(api.render(v, None, t, P), target)
        loss.backward()
        return loss

    ms = timeit(step, reps=2)
    # the same step captured once in a CUDA graph and replayed (the step is ~25 short launches: eager
    # Python + ctypes launch overhead is a third of its wall time)
    ms_graph = None
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                step()
        torch.cuda.current_stream().wait_stream(side)
        v.grad = None; t.grad = None
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            loss_g = mse(api.render(v, None, t, P), target)
            loss_g.backward()
        gv_eager = None
        graph.replay(); torch.cuda.synchronize()
        gv_graph = v.grad.clone(); gt_graph = t.grad.clone()
        ms_graph = timeit(graph.replay, reps=4)
        v.grad = None; t.grad = None
        step(); torch.cuda.synchronize()
        graph_rel = float((gv_graph - v.grad).abs().max() / v.grad.abs().max()), float((gt_graph - t.grad).abs().max() / t.grad.abs().max())
    except Exception as e:                      # report, do not hide
        graph_rel = f"graph capture failed: {type(e).__name__}: {e}"
    # ---- the same step as ONE library call (mrt_train_step_mse): no autograd graph, no dL/dC tensor, buffer
    # clears + flat classification + loss on a side stream
    one = {}
    try:
        ts = api.TrainStep(P, n_views=1, tf_entries=tf.shape[0])
        vd, tdd = v.detach(), t.detach()
        v.grad = None; t.grad = None
        loss_e = step(); torch.cuda.synchronize()
        loss_1, img_1, dv_1, dt_1 = ts(vd, tdd, target); torch.cuda.synchronize()
        one["grad_rel_vs_autograd"] = [float((dv_1 - v.grad).abs().max() / v.grad.abs().max()),
                                       float((dt_1 - t.grad).abs().max() / t.grad.abs().max())]
        one["loss_rel_vs_autograd"] = abs(float(loss_1) - float(loss_e)) / abs(float(loss_e))
        one["ms"] = timeit(lambda: ts(vd, tdd, target), reps=2)
        one["ms_pipelined_8"] = timeit(lambda: ts(vd, tdd, target), reps=8)
        try:
            g1 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1):
                ts(vd, tdd, target)
            g1.replay(); torch.cuda.synchronize()
            one["ms_cuda_graph"] = timeit(g1.replay, reps=4)
        except Exception as e:
            one["ms_cuda_graph"] = f"graph capture failed: {type(e).__name__}: {e}"
    except Exception as e:                      # report, do not hide
        one["error"] = f"{type(e).__name__}: {e}"
    Vv = api.Volume(vol.cuda())
    _, _, counts = api.render_aux(Vv, None, tf.cuda(), P)
    c = counts.sum(dim=(0, 1)).tolist()
    ms_fwd = timeit(lambda: api.render(Vv, None, t.detach(), P), reps=4)
    # ---- the two kernels of the step on their own (CUDA events), and the backward's roofline record
    td = t.detach()
    packed, mm = api.fold_volume_occupancy(vol.cuda(), P)
    Pe = api.folded_params(P)
    bits = api.classify_bricks(Pe, mm, 1, td)
    flat = api.classify_bricks(Pe, mm, 1, td, flat=True)
    img, ck = api.render_forward_ckpt(Pe, None, packed, 1, td, bits)
    g = (2.0 / img.numel()) * (img - target)
    ms_fwd_ckpt = timeit(lambda: api.render_forward_ckpt(Pe, None, packed, 1, td, bits), reps=4)
    stats = torch.zeros(2, dtype=torch.int64, device="cuda")
    api.render_backward(Pe, packed, 1, td, None, None, img, g, flat_levels=flat, minmax=mm, ckpt=ck, stats=stats)
    shaded, tasks = (int(x) for x in stats.tolist())
    ms_bwd = timeit(lambda: api.render_backward(Pe, packed, 1, td, None, None, img, g, flat_levels=flat, minmax=mm, ckpt=ck), reps=4)
    ms_bwd_whole = timeit(lambda: api.render_backward(Pe, packed, 1, td, None, None, img, g, flat_levels=flat, minmax=mm), reps=4)
    import json as _json
    peaks = ROOT / "MEASURED_PEAKS.json"
    peak = float(_json.loads(peaks.read_text())["hbm_gbs"]) if peaks.exists() else 6650.0
    bytes_per_sample = 64                       # 32 B re-read of the 8 corners + 8 x 4 B of atomic read-modify-write
    achieved = shaded * bytes_per_sample / (ms_bwd * 1e-3) / 1e9
    prof = ROOT / "profiles" / "r02_bwd_cfg3.json"
    issue = None
    if prof.exists():
        try:
            issue = float(_json.loads(prof.read_text())[0]["smsp__issue_active.avg.pct_of_peak_sustained_active"].split()[0])
        except Exception:
            issue = None
    roof = {"kernel": "mrt_bwd_kernel<1,false,true,false> (segment-parallel)", "bound": "issue slots + L2 reductions (see profiles/r02_bwd_cfg3.json)",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks.exists() else "fallback 6.65 TB/s",
            "bytes_per_sample": bytes_per_sample, "shaded_samples_per_launch": shaded, "warp_tasks_per_launch": tasks,
            "segments": ck.nseg, "slots_per_segment": ck.seg_slots, "avg_launch_ms": ms_bwd,
            "issue_active_pct_ncu": issue,
            "note": "bytes = samples the backward really shades x (32 B gathered + 32 B reduced); ms = the whole mrt_render_backward "
                    "call (dvol memset, task list, march, dL/dtf reduce). The volume and its gradient are L2-resident, so HBM is the "
                    "reference line SURVEY 8(d) asks for, not what binds"}
    # gradient parity on a small scene (the oracle's autograd cannot hold 256^3 x 512^2)
    rel_v = rel_t = None
    if WITH_ORACLE:
        from scenes import small_scene
        from oracle import oracle_torch as O
        sv, _, sP = small_scene(C=1, dims=(32, 32, 32), W=48, H=48, seed=4)
        sP = replace(sP, tfMode=1)
        stf = ramp_tf(64, sigma_scale=20.0, cutoff=0.05)
        a = sv.clone().requires_grad_(True); b = stf.clone().requires_grad_(True)
        O.render(a, sP, tf=b).square().mean().backward()
        ga = sv.cuda().requires_grad_(True); gb = stf.cuda().requires_grad_(True)
        api.render(ga, None, gb, sP).square().mean().backward()
        rel_v = float((ga.grad.cpu() - a.grad).abs().max() / a.grad.abs().max())
        rel_t = float((gb.grad.cpu() - b.grad).abs().max() / b.grad.abs().max())
    return dict(cfg="cfg3", ms_fwd_bwd=ms, ms_train_step_one_call=one.get("ms"), one_call=one, ms_fwd_bwd_cuda_graph=ms_graph, graph_vs_eager_grad_rel=graph_rel, ms_fwd_only=ms_fwd, ms_forward_ckpt_kernel=ms_fwd_ckpt, ms_backward_call=ms_bwd,
                ms_backward_whole_ray=ms_bwd_whole, steps_per_s=1e3 / ms, samples_taken=c[1],
                gsamples_per_s_fwd_bwd=c[1] / ms / 1e6, grad_rel_volume=rel_v, grad_rel_tf=rel_t, roofline=roof,
                note="fwd+bwd through the autograd API incl. layout + occupancy build, classify, checkpointing march, adjoint, unfold")


def cfg4():
    from mri_raytracer_b200 import OrbitalCamera, orbit_views
    import numpy as np
    dims = (512, 512, 512)
    vol = make_brats_like(1, dims, seed=5, device="cuda")
    tf = ramp_tf(256).cuda()
    P = replace(framed_params(dims, 2048, 2048, theta_deg=0.0), tfMode=1)
    V = api.Volume(vol)
    cam = V.frame_camera(OrbitalCamera(initial_phi=math.radians(80.0)))
    cam.set_fov_degrees(70.0)
    cams = orbit_views(cam, 64)[::8]
    taken = 0
    for c in cams:
        _, _, counts = api.render_aux(V, c, tf, P)
        taken += int(counts[..., 1].sum())
    out = torch.empty((len(cams), 2048, 2048, 4), device="cuda")

    def batch():
        api.render_views(V, cams, tf, P, out=out)                 # one classify + one batched march
    ms = timeit(batch, reps=1)
    return dict(cfg="cfg4", views_timed=len(cams), ms_per_view=ms / len(cams), fps=len(cams) * 1e3 / ms,
                samples_taken_per_view=taken / len(cams), gsamples_per_s=taken / ms / 1e6,
                note="8 of the 64 orbit views (every 8th) on one GPU; volume generated on device")


def inr():
    """INR predict_volume over a BraTS-sized case with the reference's network
    (31 -> 64 x 4 -> 4, inr/interactive.ipynb cell 1): the tcgen05 tensor-core kernel, the fp32 FFMA
    kernel (its parity reference), and the numpy oracle on a 1/64 sample as the CPU leg."""
    import json as _json
    import numpy as np
    from mri_raytracer_b200 import volume as mvol
    dims = (240, 240, 155)
    X, Y, Z = dims
    rng = np.random.default_rng(0)
    params = glorot_mlp(rng, 3 + 3 * 2 * 4 + 4, [64, 64, 64, 64], 4)       # in_dim = coords + 6k Fourier + M (inr/inr/train.py)
    mods = mvol.zscore_modalities(make_brats_like(4, dims, seed=0, device="cuda"))
    wdev = api.inr_upload_params(params, mods.device)          # the network is uploaded once, like the volume
    ms = timeit(lambda: api.inr_predict(mods, wdev, 4, impl="tensor"), reps=4)
    ms_ffma = timeit(lambda: api.inr_predict(mods, wdev, 4, impl="ffma"), reps=2)
    lt, gt = api.inr_predict(mods, params, 4, return_logits=True, impl="tensor")
    lf, gf = api.inr_predict(mods, params, 4, return_logits=True, impl="ffma")
    top2 = torch.sort(gf, dim=-1).values[..., -2:]
    clear = (top2[..., 1] - top2[..., 0]) > 1e-3
    nvox = X * Y * Z
    flop = 2.0 * (31 * 64 + 3 * 64 * 64 + 64 * 4) * nvox                      # the network's own multiply-adds
    flop_issued = 2.0 * 3 * (32 * 64 + 3 * 64 * 64 + 64 * 16) * nvox + 2.0 * 8 * (4 * 64 + 16) * nvox   # 3-term tf32 split, padded, + bias steps
    peaks = ROOT / "MEASURED_PEAKS.json"
    bf16 = float(_json.loads(peaks.read_text())["bf16_tflops"]) if peaks.exists() else 1590.0
    roof = {"bound": "tensor", "kernel": "mrt_inr_tc_kernel (tcgen05.mma kind::tf32, A from TMEM)",
            "achieved": flop_issued / ms / 1e9, "peak": bf16 / 2.0, "unit": "TFLOP/s", "frac": flop_issued / ms / 1e9 / (bf16 / 2.0),
            "peak_source": ("MEASURED_PEAKS.json bf16_tflops / 2" if peaks.exists() else "fallback 1.59 PFLOP/s / 2")
                           + " (tf32 runs at half the bf16 rate; no tf32 peak is measured on this pool)",
            "useful_tflops": flop / ms / 1e9, "issued_tflops": flop_issued / ms / 1e9, "traffic": None,
            "limiter": "the per-layer hand-off between the epilogue warps and the tensor core (two tile slots fit the 512 TMEM "
                       "columns): MMA-only the kernel takes 0.96 ms, the empty hand-off skeleton 0.47 ms (DESIGN.md)",
            "note": "achieved counts the tensor-core work really issued (every product as 3 tf32 MMAs, K padded 31->32, classes "
                    "padded 4->16, one bias step per layer); useful_tflops counts the network's own 29 kFLOP per voxel"}
    rec = dict(cfg="inr_predict", dims=dims, ms=ms, ms_fp32_ffma_kernel=ms_ffma, gvoxels_per_s=nvox / ms / 1e6,
               max_logit_diff_vs_ffma=float((gt - gf).abs().max()),
               labels_equal_where_top2_gap_over_1e3=bool((lt == lf)[clear].all()), label_agreement=float((lt == lf).float().mean()),
               roofline=roof)
    if WITH_ORACLE:
        from oracle import oracle_inr as I
        sub = mods[:, ::4, ::4, ::4].cpu().numpy().transpose(0, 3, 2, 1).copy()
        t0 = time.perf_counter()
        I.predict_volume(params, sub, 4)
        cpu_s = time.perf_counter() - t0
        rec.update(cpu_numpy_oracle_voxels_per_s=sub[0].size / cpu_s, cpu_sample="every 4th voxel per axis (1/64 of the case)",
                   speedup_vs_numpy=(nvox / (ms * 1e-3)) / (sub[0].size / cpu_s))
    return rec


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg1", "cfg3", "cfg4"]
    for w in which:
        t0 = time.time()
        r = dict(cfg1=cfg1, cfg1_u8=cfg1_u8, cfg3=cfg3, cfg4=cfg4, inr=inr)[w]()
        r["wall_s"] = time.time() - t0
        print(json.dumps(r), flush=True)
