"""dev: time the backward kernel alone at cfg3 (256^3, 512^2): whole-ray vs segment-parallel (several
segment lengths), dL/dvolume only / dL/dtf only / both, plus the checkpointing forward.
MRT_BWD_HIST=0 in the environment switches the shared-memory dL/dtf histogram off (L2 reductions)."""
import json
import os
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


def t(fn, n=6, reps=4):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / reps)
    ts.sort()
    return ts[len(ts) // 2]


res = {"hist": os.environ.get("MRT_BWD_HIST", "1")}
res["forward_ms"] = t(lambda: api.render_forward(P, packed, 1, tf, bits, out=out))
kw = dict(flat_levels=flat, minmax=mm)
for name, extra in (("both", {}), ("dvol", dict(want_dtf=False)), ("dtf", dict(want_dvol=False))):
    res[f"whole_ray_{name}_ms"] = t(lambda: api.render_backward(P, packed, 1, tf, None, None, out, g, **kw, **extra))
for S in (16, 32, 64):
    img, ck = api.render_forward_ckpt(P, None, packed, 1, tf, bits, seg_slots=S)
    assert torch.equal(img, out)
    res[f"S{S}_forward_ckpt_ms"] = t(lambda: api.render_forward_ckpt(P, None, packed, 1, tf, bits, seg_slots=S))
    for name, extra in (("both", {}), ("dvol", dict(want_dtf=False)), ("dtf", dict(want_dvol=False))):
        res[f"S{S}_{name}_ms"] = t(lambda: api.render_backward(P, packed, 1, tf, None, None, out, g, ckpt=ck, **kw, **extra))
    stats = torch.zeros(2, dtype=torch.int64, device="cuda")
    api.render_backward(P, packed, 1, tf, None, None, out, g, ckpt=ck, stats=stats, **kw)
    res[f"S{S}_nseg"] = ck.nseg
    res[f"S{S}_shaded_slots"], res[f"S{S}_tasks"] = (int(x) for x in stats.tolist())
print(json.dumps(res), flush=True)
