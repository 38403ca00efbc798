"""Multi-GPU correctness check (torchrun, 2+ GPUs): the distributed framebuffer (tile and view
partitions, striped and root owners, several batches back to back without host syncs), the NCCL
gather paths, sort-last through NCCL and through peer memory, data-parallel differentiable
rendering and differentiable sort-last (sharded fp32 / fp16 backward) — all against single-GPU renders.  Prints one line per check and ALL OK / SOME FAILED."""
import os, sys
from dataclasses import replace
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch, torch.distributed as dist
from mri_raytracer_b200 import api, dist as mdist
from mri_raytracer_b200.synth import make_brats_like, ramp_tf
from scenes import framed_params
import bench

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok_all = True

def report(name, ok, extra=""):
    global ok_all
    t = torch.tensor([1 if ok else 0], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN)
    ok_all &= bool(t.item())
    if rank == 0:
        print(f"{name}: {'OK' if t.item() else 'FAIL'} {extra}", flush=True)

# 1. image space: distributed framebuffer, every partition x owner layout, 3 batches back to back
V = 2 * world
P, _ = bench._scene(V)
P = replace(P, imageSize=(256, 256))
vol = make_brats_like(4, bench.DIMS, seed=0, device=dev); tf = ramp_tf(256).to(dev)
volume = api.Volume(vol)
for partition in ("tiles", "views"):
    for owners in ("striped", "root"):
        fb = mdist.PeerFramebuffer(V, 256, 256, dev, owners=owners, partition=partition)
        kept, refs = [], []
        for b in range(3):                         # no host sync between batches: the double buffer must hold
            _, cams = bench._scene(V, theta0_deg=25.0 + 40.0 * b)
            fb.render(volume, cams, tf, P)
            kept.append(fb.finish())
            refs.append(cams)
        ok = True
        own = fb.owned_views()
        for b in (1, 2):                            # batch 0's buffer was legitimately reused by batch 2
            ref = api.render_views(volume, refs[b], tf, P)
            ok = ok and bool(torch.equal(kept[b], ref[own.start:own.stop]))
        report(f"framebuffer partition={partition} owners={owners} (p2p={fb.p2p})", ok)
        del fb
for mode in ("views", "tiles"):
    _, cams = bench._scene(V)
    got = mdist.render_views(volume, cams, tf, P, mode=mode)
    ref = api.render_views(volume, cams, tf, P)
    report(f"image-space NCCL all_gather mode={mode}", bool(torch.equal(got, ref)))

# 2. sort-last: NCCL exchange and peer exchange vs the unsharded render
dims = (72, 60, 52)
volc = make_brats_like(1, dims, seed=3, device=dev)
Ps = replace(framed_params(dims, 200, 136, theta_deg=33.0, phi_deg=64.0), tfMode=1, ertThreshold=1e-6, bgColor=(0.1, 0.2, 0.3))
tfs = ramp_tf(64, sigma_scale=10.0, cutoff=0.1).to(dev)
grid = mdist.shard_grid(world)
lo, hi, _ = mdist.shard_box(dims, grid, rank)
for half in (False, True):
    src = volc.half() if half else volc
    full = api.render(api.Volume(src), None, tfs, Ps)
    sv = api.Volume(mdist.slice_shard(src, lo, hi), shard=(lo, hi), global_dims=dims)
    a = mdist.render_sort_last(sv, None, tfs, Ps, grid)
    report(f"sort-last NCCL (half={half}) vs unsharded", float((a - full).abs().max()) <= 2e-5, f"max {float((a - full).abs().max()):.2e}")
    ex = mdist.PeerSortLast(136, 200, dev)
    for it in range(3):
        b = ex.render(sv, None, tfs, Ps, grid)
    report(f"sort-last peer exchange (half={half}, p2p={ex.p2p}) == NCCL path", bool(torch.equal(a, b.clone())))
# 3. differentiable rendering, data-parallel over tiles: gradients after all_reduce == single-GPU gradients
dimg = (64, 56, 48)
vg = make_brats_like(2, dimg, seed=9, device=dev)
Pg = replace(framed_params(dimg, 120, 88), tfMode=1)
tfg = ramp_tf(64, sigma_scale=20.0, cutoff=0.1).to(dev)
wgt = torch.rand((88, 120, 4), generator=torch.Generator().manual_seed(2)).to(dev)
v1 = vg.clone().requires_grad_(True); t1 = tfg.clone().requires_grad_(True)
(api.render(v1, None, t1, Pg) * wgt).sum().backward()
v2 = vg.clone().requires_grad_(True); t2 = tfg.clone().requires_grad_(True)
img = mdist.render_differentiable(v2, None, t2, Pg)
(img * wgt).sum().backward()
mdist.allreduce_gradients([v2, t2])
rv = float((v2.grad - v1.grad).abs().max() / v1.grad.abs().max()); rt = float((t2.grad - t1.grad).abs().max() / t1.grad.abs().max())
report("differentiable tiles + all_reduce gradients == single GPU", rv <= 1e-5 and rt <= 1e-5, f"rel dvol {rv:.1e} dtf {rt:.1e}")
# 4. differentiable sort-last (cfg5 trains): every rank differentiates its own sub-box; the strips' losses add
#    up to the frame's, dL/d(sub) stays on the rank that stores the voxels, dL/dtf is all-reduced
grid = mdist.shard_grid(world)
for storage in (None, torch.float16):
    vfull = volc if storage is None else volc.half().float()
    lo, hi, _ = mdist.shard_box(dims, grid, rank)
    sub = mdist.slice_shard(vfull, lo, hi).clone().requires_grad_(True)
    t3 = tfs.clone().requires_grad_(True)
    wg = torch.rand((136, 200, 4), generator=torch.Generator().manual_seed(4)).to(dev)
    strip, row0 = mdist.render_sort_last_differentiable(sub, None, t3, Ps, grid, storage=storage)
    valid = max(0, min(strip.shape[0], 136 - row0))
    (strip[:valid] * wg[row0:row0 + valid]).sum().backward()
    mdist.allreduce_gradients([t3])
    vr = vfull.clone().requires_grad_(True); tr = tfs.clone().requires_grad_(True)
    (api.render(vr, None, tr, Ps) * wg).sum().backward()
    # assemble the global gradient from the shards' (halo voxels are shared: SUM)
    gfull = torch.zeros_like(vfull)
    gfull[:, lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1] = sub.grad.float()
    dist.all_reduce(gfull)
    rv = float((gfull - vr.grad).abs().max() / vr.grad.abs().max()); rt = float((t3.grad - tr.grad).abs().max() / tr.grad.abs().max())
    report(f"differentiable sort-last (storage={'f16' if storage else 'f32'}) gradients == single GPU", rv <= 1e-3 and rt <= 1e-3,
           f"rel dvol {rv:.1e} dtf {rt:.1e}")
if rank == 0:
    print("ALL OK" if ok_all else "SOME FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok_all else 1)
