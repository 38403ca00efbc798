"""Dev check (2+ GPUs): frames gathered on rank 0 through the peer framebuffer equal local renders."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import torch, torch.distributed as dist
from dataclasses import replace
from mri_raytracer_b200 import api, dist as mdist
from mri_raytracer_b200.synth import make_brats_like, ramp_tf
import bench
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
V = 2
P, cams_all = bench._scene(V * world)
P = replace(P, imageSize=(256, 256))
vol = make_brats_like(4, bench.DIMS, seed=0, device=dev); tf = ramp_tf(256).to(dev)
volume = api.Volume(vol)
fb = mdist.PeerFramebuffer(V, 256, 256, dev)
print(rank, "p2p", fb.p2p, getattr(fb, "why", ""), flush=True)
mdist.render_views_to(fb, volume, cams_all[rank * V:(rank + 1) * V], tf, P, cams_all=cams_all); fb.finish()
torch.cuda.synchronize(); dist.barrier()
if rank == 0:
    ok = True
    for v, c in enumerate(cams_all):
        ref = api.render(volume, c, tf, P)
        ok &= bool(torch.equal(ref, fb.frames()[v]))
    print("gathered frames identical to local renders:", ok, flush=True)
dist.destroy_process_group()
