"""Profiling target (ncu): cfg3 forward+backward (256^3, 512^2, MSE loss) a few times."""
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
with torch.no_grad():
    target = api.render(api.Volume(vol), None, (tf * torch.tensor([0.8, 1.0, 1.1, 1.3], device="cuda")), P)
v = vol.requires_grad_(True); t = tf.requires_grad_(True)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    v.grad = None; t.grad = None
    ((api.render(v, None, t, P) - target) ** 2).mean().backward()
torch.cuda.synchronize()
print("ok", float(v.grad.abs().sum()), float(t.grad.abs().sum()))
