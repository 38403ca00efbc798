"""Profiling target (ncu): the tensor-core INR kernel on a slab of a BraTS-sized case."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
from mri_raytracer_b200 import api, volume as mvol
from mri_raytracer_b200.synth import make_brats_like
from oracle import oracle_inr as I
dims = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (240, 240, 32)
impl = sys.argv[4] if len(sys.argv) > 4 else "tensor"
rng = np.random.default_rng(11)
params = I.init_mlp(rng, I.input_dim(4, 4), [64, 64, 64, 64], 4)
mods = mvol.zscore_modalities(make_brats_like(4, dims, seed=0, device="cuda"))
for _ in range(3):
    lab = api.inr_predict(mods, params, 4, impl=impl)
torch.cuda.synchronize()
print("ok", int(lab.sum()))
