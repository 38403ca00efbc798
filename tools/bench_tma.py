"""Staged-brick (TMA + shared memory) march vs the direct-gather march, same scene, same image.
  python tools/bench_tma.py [cfg2|cfg4] [--once]     (--once: one launch of each variant, for ncu)
cfg2: BraTS-shaped 240x240x155 folded to one channel (36 MB, L2-resident), 1024^2, one view.
cfg4: 512^3 single channel (537 MB > L2), 2048^2, one view.
Prints one JSON line: per variant the median launch time (L2 flushed between launches), how many slots
went through the shared-memory stage, and max |image - direct image|."""
import json
import sys
from dataclasses import replace
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
from mri_raytracer_b200 import api
from mri_raytracer_b200.synth import make_brats_like, ramp_tf
from scenes import framed_params

cfg = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "cfg2"
once = "--once" in sys.argv
if cfg == "cfg2":
    dims, W, C = (240, 240, 155), 1024, 4
else:
    dims, W, C = (512, 512, 512), 2048, 1
vol = make_brats_like(C, dims, seed=0 if cfg == "cfg2" else 5, device="cuda")
tf = ramp_tf(256).cuda()
P = replace(framed_params(dims, W, W, theta_deg=25.0 if cfg == "cfg2" else 0.0), tfMode=1)
Vs = api.Volume(vol, quad=False)
Vq = api.Volume(vol, quad=True)
packed, Cn, Pe = Vs.prepared(P)
bits = Vs.skip_levels(P, tf)
quad, _, Pq = Vq.prepared(P)
out = torch.empty((W, W, 4), device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, n=7):
    if once:
        fn(); torch.cuda.synchronize(); return None
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.fill_(1); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


res = {"config": cfg, "dims": dims, "image": W, "volume_MB": packed.numel() * 4 / 1e6}
ref = api.render_forward(Pe, packed, Cn, tf, bits).clone()
res["direct_scalar_ms"] = timed(lambda: api.render_forward(Pe, packed, Cn, tf, bits, out=out))
res["direct_quad_ms"] = timed(lambda: api.render_forward(Pq, quad, 1, tf, bits, out=out))
assert torch.equal(out, ref)
for box, tile in ((8, 8), (16, 8), (8, 16), (16, 16)):
    stats = torch.zeros(4, dtype=torch.int64, device="cuda")
    img = api.render_forward_tma(Pe, packed, tf, bits, box_edge=box, tile=tile, stats=stats)
    st = [int(x) for x in stats.tolist()]
    key = f"tma_box{box}_tile{tile}"
    res[key + "_ms"] = timed(lambda: api.render_forward_tma(Pe, packed, tf, bits, box_edge=box, tile=tile, out=out))
    res[key] = dict(max_abs_vs_direct=float((img - ref).abs().max()), staged_slots=st[0], direct_slots=st[1], boxes_staged=st[2],
                    overflowed_lists=st[3], staged_MB=st[2] * (12 * 81 if box == 8 else 20 * 289) * 4 / 1e6)
print(json.dumps(res), flush=True)
