"""Profiling target (ncu launch list): three cfg3 training steps (256^3, 512^2) through the autograd API."""
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
vol = make_brats_like(1, dims, seed=4)
tf = ramp_tf(256, sigma_scale=20.0, cutoff=0.05)
P = replace(framed_params(dims, 512, 512), tfMode=1)
with torch.no_grad():
    target = api.render(api.Volume(vol.cuda()), None, (tf * torch.tensor([0.8, 1.0, 1.1, 1.3])).cuda(), P)
v = vol.cuda().requires_grad_(True); t = tf.cuda().requires_grad_(True)
for i in range(3):
    v.grad = None; t.grad = None
    torch.cuda.synchronize()
    if i == 2:
        torch.cuda.nvtx.range_push("step")
    loss = ((api.render(v, None, t, P) - target) ** 2).mean()
    loss.backward()
    torch.cuda.synchronize()
    if i == 2:
        torch.cuda.nvtx.range_pop()
print("ok", float(loss))
