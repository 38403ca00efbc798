"""CUDA-event timing of the fold + occupancy passes: cfg2 (4 channels -> quad layout) and cfg3 (1 channel, 256^3)."""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import bench
from mri_raytracer_b200 import api
from mri_raytracer_b200.synth import make_brats_like
from scenes import framed_params

def timeit(fn, n=20):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]

P2, _ = bench._scene(8)
v2 = make_brats_like(4, bench.DIMS, seed=0).cuda()
q = mm = None
def f2():
    global q, mm
    q, mm = api.fold_volume_occupancy_quad(v2, P2, q, mm)
d3 = (256, 256, 256)
v3 = make_brats_like(1, d3, seed=4).cuda()
P3 = framed_params(d3, 512, 512)
print(json.dumps({"fold_occ_quad_cfg2_ms": timeit(f2), "fold_occ_cfg3_ms": timeit(lambda: api.fold_volume_occupancy(v3, P3)),
                  "fold_occ_c4_scalar_cfg2_ms": timeit(lambda: api.fold_volume_occupancy(v2, P2))}))
