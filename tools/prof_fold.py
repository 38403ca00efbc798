"""Profiling target (ncu): the bench step's prepare phase — fold + occupancy + quad layout, classify —
re-run every repetition (volume.invalidate()), then the batched march.  `python tools/prof_fold.py [reps]`."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import bench
from mri_raytracer_b200 import api
from mri_raytracer_b200.synth import make_brats_like, ramp_tf

def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    P, cams = bench._scene(8)
    vol = make_brats_like(bench.NCH, bench.DIMS, seed=0).cuda()
    tf = ramp_tf(bench.TF_N).cuda()
    volume = api.Volume(vol)
    frames = torch.empty((8, bench.IMG, bench.IMG, 4), device="cuda")
    for _ in range(reps):
        volume.invalidate()
        api.render_views(volume, cams, tf, P, out=frames)
    torch.cuda.synchronize()
    print("ok", float(frames.sum()))

if __name__ == "__main__":
    main()
