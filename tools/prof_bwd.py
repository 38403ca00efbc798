"""Profiling target (ncu): the segment-parallel backward kernel at cfg3 (256^3, 512^2), a few launches.
argv: [n_launches] [seg_slots] [both|dvol|dtf]"""
import sys
from dataclasses import replace
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
from mri_raytracer_b200 import api
from mri_raytracer_b200.synth import make_brats_like, ramp_tf
from scenes import framed_params
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
S = int(sys.argv[2]) if len(sys.argv) > 2 else 32
what = sys.argv[3] if len(sys.argv) > 3 else "both"
dims = (256, 256, 256)
vol = make_brats_like(1, dims, seed=4, device="cuda")
tf = ramp_tf(256, sigma_scale=20.0, cutoff=0.05).cuda()
P = replace(framed_params(dims, 512, 512), tfMode=1)
packed = api.pack_volume(vol)
mm = api.build_occupancy(packed, 1, dims)
bits = api.classify_bricks(P, mm, 1, tf)
flat = api.classify_bricks(P, mm, 1, tf, flat=True)
out, ck = api.render_forward_ckpt(P, None, packed, 1, tf, bits, seg_slots=S)
g = torch.rand_like(out)
for _ in range(n):
    dv, dt = api.render_backward(P, packed, 1, tf, None, None, out, g, flat_levels=flat, minmax=mm, ckpt=ck if S > 0 else None,
                                 want_dvol=what != "dtf", want_dtf=what != "dvol")
torch.cuda.synchronize()
print("ok", None if dv is None else float(dv.abs().sum()), None if dt is None else float(dt.abs().sum()))
