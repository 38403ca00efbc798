"""Stage times of ONE mrt_train_step_mse call at cfg3 (256^3 volume, 512^2, 256-entry LUT), from timing
events the library records on the caller's stream after every stage (mrt_debug_train_trace).

    python tools/time_train_step.py            # side stream on
    MRT_TRAIN_NO_SIDE=1 python tools/time_train_step.py
"""
import ctypes as C
import json
import os
import sys
from dataclasses import replace
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402

from mri_raytracer_b200 import api  # noqa: E402
from mri_raytracer_b200._lib import lib  # noqa: E402
from mri_raytracer_b200.synth import make_brats_like, ramp_tf  # noqa: E402
from scenes import framed_params  # noqa: E402


def main():
    dims = (256, 256, 256)
    vol = make_brats_like(1, dims, seed=4).cuda()
    tf = ramp_tf(256, sigma_scale=20.0, cutoff=0.05).cuda()
    P = replace(framed_params(dims, 512, 512), tfMode=1)
    with torch.no_grad():
        target = api.render(api.Volume(vol), None, tf * torch.tensor([0.8, 1.0, 1.1, 1.3], device="cuda"), P)
    ts = api.TrainStep(P, n_views=1, tf_entries=256)
    for _ in range(3):
        ts(vol, tf, target)
    torch.cuda.synchronize()
    trace = lib().mrt_debug_train_trace
    rows = []
    for _ in range(5):
        trace(None)
        ts(vol, tf, target)
        ms = (C.c_float * 6)()
        assert trace(C.cast(ms, C.c_void_p)) == 0
        rows.append([float(x) for x in ms])
    rows.sort(key=sum)
    med = rows[len(rows) // 2]
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(8):
        ts(vol, tf, target)
    b.record(); torch.cuda.synchronize()
    print(json.dumps({"side_stream": os.environ.get("MRT_TRAIN_NO_SIDE") is None,
                      "stages_ms": dict(zip(["fold_occ", "classify", "march_ckpt", "adjoint", "fold_adjoint", "join"], med)),
                      "sum_ms": sum(med), "ms_per_step_pipelined_8": a.elapsed_time(b) / 8}))


if __name__ == "__main__":
    main()
