"""dev: time mrt_fill_masked_tiles and the sparse vs dense batched march on one GPU."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch, bench
from dataclasses import replace
from mri_raytracer_b200 import api
from mri_raytracer_b200.synth import make_brats_like, ramp_tf
V = 64
P, cams = bench._scene(V)
vol = make_brats_like(4, bench.DIMS, seed=0).cuda(); tf = ramp_tf(256).cuda()
volume = api.Volume(vol)
out = torch.empty((V, 1024, 1024, 4), device="cuda")
mask = torch.zeros(api.sparse_mask_bytes(1024, 1024, V), dtype=torch.uint8, device="cuda")
Pm = replace(P, tfMode=1)
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
print("dense batch 64 views ms", t(lambda: api.render_views(volume, cams, tf, P, out=out)))
print("sparse batch 64 views ms", t(lambda: volume.forward_batch_sparse(Pm, cams, tf, out.data_ptr(), mask.data_ptr())))
print("culled fraction", float(mask.float().mean()))
print("fill 64 views ms", t(lambda: api.fill_masked_tiles(Pm.with_camera(cams[0]), mask, V, out)))
ref = api.render_views(volume, cams, tf, P)
volume.forward_batch_sparse(Pm, cams, tf, out.data_ptr(), mask.data_ptr()); api.fill_masked_tiles(Pm.with_camera(cams[0]), mask, V, out)
print("equal", bool(torch.equal(ref, out)))
