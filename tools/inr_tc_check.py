"""dev: tensor-core INR kernel vs the fp32 FFMA kernel (logit difference, label agreement) and timings."""
import sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
from mri_raytracer_b200 import api, volume as mvol
from mri_raytracer_b200.synth import make_brats_like
from oracle import oracle_inr as I
dims = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (131, 37, 29)
X, Y, Z = dims
rng = np.random.default_rng(11)
params = I.init_mlp(rng, I.input_dim(4, 4), [64, 64, 64, 64], 4)
for p in params:
    p["b"] = rng.normal(scale=0.2, size=p["b"].shape).astype(np.float32)
mods = mvol.zscore_modalities(make_brats_like(4, dims, seed=0, device="cuda"))
lf, gf = api.inr_predict(mods, params, 4, return_logits=True, impl="ffma")
torch.cuda.synchronize()
print("ffma done", flush=True)
lt, gt = api.inr_predict(mods, params, 4, return_logits=True, impl="tensor")
torch.cuda.synchronize()
d = (gt - gf).abs()
print(json.dumps(dict(dims=dims, max_logit_diff=float(d.max()), mean_logit_diff=float(d.mean()), label_agree=float((lt == lf).float().mean()),
                      logit_scale=float(gf.abs().max()))), flush=True)
if float(d.max()) > 1e-3:
    idx = torch.nonzero(d.amax(-1) > 1e-3)[:8]
    print("first bad voxels (z,y,x):", idx.tolist(), flush=True)
    for z, y, x in idx.tolist()[:3]:
        print(" ffma", gf[z, y, x].tolist(), "tc", gt[z, y, x].tolist(), flush=True)
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
print(json.dumps(dict(ms_tensor=t(lambda: api.inr_predict(mods, params, 4, impl="tensor")),
                      ms_ffma=t(lambda: api.inr_predict(mods, params, 4, impl="ffma")))), flush=True)
