"""BASELINE config 4: a batch of 64 orbit views at 2048x2048 over a 512^3 fp32 volume, image-space
partition across 1/2/4/8 B200 with the framebuffer gathered on rank 0.

    python tools/bench_cfg4.py                                   # one GPU
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_cfg4.py

STRONG scaling: the 64 views are fixed; rank r renders tile rows ty % R == r of EVERY view (or whole
views with --partition views) in ONE batched launch and stores them straight into the peer-mapped
frame of each view's owner GPU (dist.PeerFramebuffer; owners striped over the ranks, or all on rank 0
with --owners root; tiles outside the projected active-brick box are not sent, the owner fills them).
Prints one JSON line: ms per 64-view batch (CUDA events, max over ranks), views/s, nominal samples/s,
and a bit-exact check of owned views against local renders on every rank.
"""
import argparse
import json
import math
import os
import sys
from dataclasses import replace
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from mri_raytracer_b200 import OrbitalCamera, api, orbit_views  # noqa: E402
from mri_raytracer_b200 import dist as mdist  # noqa: E402
from mri_raytracer_b200.synth import make_brats_like, ramp_tf  # noqa: E402
from scenes import framed_params  # noqa: E402


def run(args, rank, world, dev):
    """The measurement itself; the process group (world > 1) is the caller's.  Returns the record on every rank."""
    if args.views % world:
        raise ValueError("--views must be divisible by the number of GPUs")
    dims = (args.dim,) * 3
    vol = make_brats_like(1, dims, seed=5, device=dev)
    tf = ramp_tf(256).to(dev)
    P = replace(framed_params(dims, args.img, args.img, theta_deg=0.0), tfMode=1)
    V = api.Volume(vol)
    cam = V.frame_camera(OrbitalCamera(initial_phi=math.radians(80.0)))
    cam.set_fov_degrees(70.0)
    cams_all = orbit_views(cam, args.views)
    Vloc = args.views // world
    mine = cams_all[rank * Vloc:(rank + 1) * Vloc]

    # nominal sample count of the whole batch (each rank counts its own views)
    taken = 0
    for c in mine:
        _, _, counts = api.render_aux(V, c, tf, P)
        taken += int(counts[..., 1].sum())
    tot = torch.tensor([taken], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot)
    taken = int(tot)

    fb = mdist.PeerFramebuffer(args.views, args.img, args.img, dev, owners=args.owners, partition=args.partition) if world > 1 else None
    frames = torch.empty((Vloc, args.img, args.img, 4), device=dev) if world == 1 else None
    got_holder = [None]

    def batch():
        if world == 1:
            api.render_views(V, mine, tf, P, out=frames)
        else:
            fb.render(V, cams_all, tf, P)
            got_holder[0] = fb.finish()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(2):
        batch()
    barrier()
    ms = []
    for _ in range(args.reps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); batch(); b.record()
        barrier()
        ms.append(a.elapsed_time(b))
    t = torch.tensor([sorted(ms)[len(ms) // 2]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_batch = float(t)
    # every owner checks its first and last owned view against a local render (bit-exact)
    okl = True
    if world == 1:
        for v in (0, args.views // 2 + 1, args.views - 1):
            okl &= bool(torch.equal(frames[v], api.render(V, cams_all[v], tf, P)))
    else:
        own = fb.owned_views()
        for v in sorted({own.start, own.stop - 1}) if len(own) else []:
            okl &= bool(torch.equal(got_holder[0][v - own.start], api.render(V, cams_all[v], tf, P)))
    okt = torch.tensor([1 if okl else 0], device=dev)
    if world > 1:
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    return dict(cfg="cfg4", dims=dims, image=args.img, views=args.views, n_gpus=world, scaling="strong",
                gather=("none (single GPU)" if world == 1 else
                        f"peer (NVLink) stores from inside the march, partition={fb.partition}, owners={fb.owners}"),
                ms_per_batch=ms_batch, views_per_s=args.views * 1e3 / ms_batch,
                ms_per_view=ms_batch / args.views, nominal_samples_per_batch=taken,
                gsamples_per_s=taken / ms_batch / 1e6, gathered_views_equal_local_renders=bool(okt.item()),
                reps=args.reps)

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--img", type=int, default=2048)
    ap.add_argument("--views", type=int, default=64)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--partition", default="tiles", choices=["tiles", "views"])
    ap.add_argument("--owners", default="striped", choices=["striped", "root"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rec = run(args, rank, world, dev)
    if rank == 0:
        print(json.dumps(rec), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
