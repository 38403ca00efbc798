"""Dev-only timing probe (not the official bench): cfg2 frame time, with/without skipping."""
import sys, time, json
from dataclasses import replace
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import torch
from mri_raytracer_b200 import api
from mri_raytracer_b200.synth import make_brats_like, ramp_tf
from scenes import framed_params

def timeit(fn, n=5, warm=3, reps=10):
    """median over n of (reps back-to-back calls)/reps: launches pipeline, CPU overhead hidden"""
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): fn()
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / reps)
    ts.sort()
    return ts[len(ts) // 2]

def main():
    C = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    dims = (240, 240, 155)
    vol = make_brats_like(C, dims, seed=0, device="cuda")
    fold = (len(sys.argv) > 3 and sys.argv[3] == "fold")
    V = api.Volume(vol, fold=fold)
    tf = ramp_tf(256).cuda()
    P = framed_params(dims, W, W)
    for skip in (0, 1):
        Ps = replace(P, skipEmpty=skip)
        img, T, counts = api.render_aux(V, None, tf, Ps)
        c = counts.sum(dim=(0, 1)).tolist()
        ms = timeit(lambda: api.render(V, None, tf, Ps))
        print(json.dumps(dict(C=C, W=W, fold=fold, skip=skip, ms=ms, n_clip=c[0], n_taken=c[1], n_eval=c[2], n_seg=c[3],
                              gsamples_taken_s=c[1] / ms / 1e6, gsamples_eval_s=c[2] / ms / 1e6,
                              alg_GBs=c[2] * 32 * C / ms / 1e6)))
    Pi = replace(P, skipEmpty=1)
    ms = timeit(lambda: api.render(V, None, None, Pi))
    print("brats-intensity TF, skip=1: ms", ms)

if __name__ == "__main__":
    main()
