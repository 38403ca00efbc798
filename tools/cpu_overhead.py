"""dev: host-side (Python + ctypes + launch) time per step of the batched render paths vs their GPU time."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import bench
from mri_raytracer_b200 import api, dist as mdist
from mri_raytracer_b200.synth import make_brats_like, ramp_tf

V = 8
P, cams = bench._scene(V)
vol = make_brats_like(bench.NCH, bench.DIMS, seed=0).cuda()
tf = ramp_tf(bench.TF_N).cuda()
volume = api.Volume(vol)
frames = torch.empty((V, bench.IMG, bench.IMG, 4), device="cuda")
fb = mdist.PeerFramebuffer(V, bench.IMG, bench.IMG, torch.device("cuda", 0))

def measure(name, fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(n): fn()
    t1 = time.perf_counter(); b.record(); torch.cuda.synchronize()
    print(f"{name}: host {1e3 * (t1 - t0) / n:.3f} ms/step, gpu {a.elapsed_time(b) / n:.3f} ms/step", flush=True)

measure("render_views static", lambda: api.render_views(volume, cams, tf, P, out=frames))
def refold():
    volume.invalidate(); api.render_views(volume, cams, tf, P, out=frames)
measure("render_views refold", refold)
def fbstep():
    fb.render(volume, cams, tf, P); fb.finish()
measure("PeerFramebuffer static (1 rank)", fbstep)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(200): fbstep()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(45)
