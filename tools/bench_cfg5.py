"""BASELINE config 5: a brick-sharded fp16 volume, sort-last front-to-back compositing over NVLink.

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_cfg5.py [--dim 2048] [--img 4096]
    python tools/bench_cfg5.py --emulate 8 --dim 512 --img 1024      # all shards on ONE GPU (reduced size)

Every rank generates its own sub-box (cells + one halo voxel) of the synthetic volume on its GPU
(synth.make_brats_like_box: a pure function of the global voxel index), packs it as fp16, marches
every ray through its sub-box only and exchanges / composites through peer memory
(dist.PeerSortLast; NCCL all_to_all + all_gather when symmetric memory is unavailable or with
--nccl).  Prints one JSON line: ms per frame (max over ranks, CUDA events), nominal ray samples
(sum over rays of the clip count n, identical on every rank) per second, and a parity check of a
pixel subset against the C oracle on a small configuration (--parity).
"""
import argparse
import json
import math
import os
import sys
import time
from dataclasses import replace
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from mri_raytracer_b200 import Camera, OrbitalCamera, RenderParams, api, orbit_views  # noqa: E402
from mri_raytracer_b200 import dist as mdist  # noqa: E402
from mri_raytracer_b200.synth import make_brats_like_box, ramp_tf, world_box  # noqa: E402


def run(args, rank, world, dev):
    """The measurement itself; the process group (world > 1) is the caller's.  Returns the record on every rank."""
    R = args.emulate or world
    grid = mdist.shard_grid(R)
    dims = tuple(args.dims) if getattr(args, "dims", None) else (args.dim, args.dim, args.dim)
    vs, vmin = world_box(dims)
    P = RenderParams(imageSize=(args.img, args.img), dims=dims, voxelSize=tuple(float(v) for v in vs),
                     volMin=tuple(float(v) for v in vmin), stepSize=float(np.float32(0.5) * vs[0]), skipEmpty=1,
                     tfMode=1, ertThreshold=1e-4)
    cam = OrbitalCamera(initial_radius=3.0, initial_theta=math.radians(25.0), initial_phi=math.radians(80.0))
    cam.set_fov_degrees(70.0)
    ext = vs * np.asarray(dims, dtype=np.float32)
    cam.target = (vmin + 0.5 * ext).astype(np.float32); cam.radius = float(np.linalg.norm(ext) * 0.8)
    cams = orbit_views(cam, args.views)
    tf = ramp_tf(256).to(dev)

    t0 = time.perf_counter()
    my = range(R) if args.emulate else [rank]
    vols = []
    for r in my:
        lo, hi, _ = mdist.shard_box(dims, grid, r)
        sub = make_brats_like_box(dims, lo, hi, seed=5, device=dev, dtype=torch.float16)
        vols.append(api.Volume(sub, shard=(lo, hi), global_dims=dims))
        del sub
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t0

    ex = mdist.PeerSortLast(args.img, args.img, dev, emulate=args.emulate)
    use_peer = ex.p2p and not args.nccl

    def frame(c):
        if args.emulate:
            return ex.render(vols, c, tf, P, grid)
        if use_peer:
            return ex.render(vols[0], c, tf, P, grid)
        return mdist.render_sort_last(vols[0], c, tf, P, grid)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # nominal sample count: sum over rays of the clip count n (same on every rank) + what this rank evaluated
    counts = torch.zeros((args.img, args.img, 4), dtype=torch.int32, device=dev)
    vols[0].forward(P.with_camera(cams[0]), tf, out_counts=counts)
    n_clip = int(counts[..., 0].sum()); n_eval = int(counts[..., 2].sum())
    del counts
    for c in cams[:2]:
        frame(c)
    barrier()
    ms = []
    for _ in range(args.reps):
        for c in cams:
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); frame(c); b.record()
            barrier()
            ms.append(a.elapsed_time(b))
    t = torch.tensor([sum(ms) / len(ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_frame = float(t)
    rec = dict(cfg="cfg5", dims=dims, image=args.img, shards=R, grid=grid, dtype="f16",
               exchange=("emulated on one GPU" if args.emulate else ("peer stores (NVLink symmetric memory)" if use_peer else "NCCL all_to_all + all_gather")),
               ms_per_frame=ms_frame, fps=1e3 / ms_frame, nominal_clip_samples_per_frame=n_clip,
               gsamples_per_s=n_clip / ms_frame / 1e6, evaluated_by_rank0=n_eval,
               shard_gb=vols[0].packed.numel() * 2 / 1e9, gen_pack_s=t_gen, views=args.views, reps=args.reps)
    if args.check:
        lo0 = (0, 0, 0); hi0 = tuple(d - 1 for d in dims)
        full = api.Volume(make_brats_like_box(dims, lo0, hi0, seed=5, device=dev, dtype=torch.float16))
        ref = api.render(full, cams[0], tf, P)
        got = frame(cams[0])
        rec["max_abs_vs_unsharded"] = float((got - ref).abs().max())
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=2048)
    ap.add_argument("--img", type=int, default=4096)
    ap.add_argument("--views", type=int, default=4)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--emulate", type=int, default=0, help="run all R shards on one GPU")
    ap.add_argument("--nccl", action="store_true", help="use the NCCL all_to_all/all_gather exchange")
    ap.add_argument("--check", action="store_true", help="compare with the unsharded render (small sizes only)")
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
