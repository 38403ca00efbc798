"""dev: time the backward kernel alone at cfg3 with dL/dvolume only, dL/dtf only, and both."""
import sys
from dataclasses import replace
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
from mri_raytracer_b200 import api
from mri_raytracer_b200.synth import make_brats_like, ramp_tf
from scenes import framed_params
dims = (256, 256, 256)
vol = make_brats_like(1, dims, seed=4, device="cuda")
tf = ramp_tf(256, sigma_scale=20.0, cutoff=0.05).cuda()
P = replace(framed_params(dims, 512, 512), tfMode=1)
packed = api.pack_volume(vol)
mm = api.build_occupancy(packed, 1, dims)
bits = api.classify_bricks(P, mm, 1, tf)
flat = api.classify_bricks(P, mm, 1, tf, flat=True)
out = api.render_forward(P, packed, 1, tf, bits)
g = torch.rand_like(out)
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for name, kw in (("both", {}), ("dvol only", dict(want_dtf=False)), ("dtf only", dict(want_dvol=False))):
    print(name, "ms", round(t(lambda: api.render_backward(P, packed, 1, tf, None, None, out, g, flat_levels=flat, minmax=mm, **kw)), 4))
print("forward ms", round(t(lambda: api.render_forward(P, packed, 1, tf, bits, out=out)), 4))
