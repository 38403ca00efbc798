"""Profiling target (ncu): the bench's cfg2 orbit batch — fold, occupancy, classify, ONE batched
march launch — repeated a few times.  `python tools/prof_fwd.py [views] [reps]`."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import bench
from mri_raytracer_b200 import api
from mri_raytracer_b200.synth import make_brats_like, ramp_tf

def main():
    V = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    P, cams = bench._scene(V)
    vol = make_brats_like(bench.NCH, bench.DIMS, seed=0).cuda()
    tf = ramp_tf(bench.TF_N).cuda()
    volume = api.Volume(vol)
    frames = torch.empty((V, bench.IMG, bench.IMG, 4), device="cuda")
    for _ in range(reps):
        api.render_views(volume, cams, tf, P, out=frames)
    torch.cuda.synchronize()
    print("ok", float(frames.sum()))

if __name__ == "__main__":
    main()
