"""Device time of mrt_view_spans at cfg2 for 8 and 64 views (the N = 8 weak-scaling step computes the spans of all
64 views on every rank).  `python tools/time_spans.py`"""
import json
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import bench
from mri_raytracer_b200 import api
from mri_raytracer_b200.synth import make_brats_like, ramp_tf


def main():
    out = {}
    vol = make_brats_like(bench.NCH, bench.DIMS, seed=0).cuda()
    tf = ramp_tf(bench.TF_N).cuda()
    volume = api.Volume(vol)
    for V in (8, 64):
        P, cams = bench._scene(V)
        packed, Cn, Pe = volume.prepared(P)
        bits = volume.skip_levels(P, tf)
        arr = api._camera_array(cams)
        spans = api.view_spans(Pe, arr, Cn, bits)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                api.view_spans(Pe, arr, Cn, bits, out=spans)
            b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) / 10)
        ts.sort()
        r = spans.cpu()
        tiles = int(((r[..., 1] | 7) - (r[..., 0] & ~7) + 1).clamp_min(0).sum() // 8)
        out[f"views_{V}"] = {"ms": ts[len(ts) // 2], "in_span_tiles_per_view": tiles / V}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
