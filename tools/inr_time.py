import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from mri_raytracer_b200 import api, volume as mvol
from mri_raytracer_b200.synth import make_brats_like
from oracle import oracle_inr as I
dims=(240,240,155)
rng=np.random.default_rng(11)
params=I.init_mlp(rng, I.input_dim(4,4), [64,64,64,64], 4)
mods=mvol.zscore_modalities(make_brats_like(4, dims, seed=0, device="cuda"))
def t(fn,n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/n
print("ms", t(lambda: api.inr_predict(mods, params, 4, impl="tensor")))
