"""Multi-GPU host logic on CPU: world_size-2 gloo process group (127.0.0.1), with duck-typed
stand-ins for the CUDA renderer objects (tests/dist_fakes.py: renders come from the oracle).  Checks
that the image-space partitions ('views' and 'tiles') + the single all_gather reproduce the
single-process image exactly, for image sizes that do not divide evenly."""
import os
import socket
import sys
from dataclasses import replace
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mode, W, H, V, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mri_raytracer_b200 import dist as mdist, orbit_views, OrbitalCamera
        from mri_raytracer_b200.synth import ramp_tf
        from scenes import small_scene
        vol, _, P = small_scene(C=2, dims=(20, 18, 16), W=W, H=H, seed=7)
        tf = ramp_tf(32, sigma_scale=20.0, cutoff=0.1)
        cam = OrbitalCamera(initial_radius=float(np.linalg.norm(np.asarray(P.eye))), initial_phi=1.3, initial_theta=0.4)
        cam.set_fov_degrees(70.0)
        cams = orbit_views(cam, V)
        from dist_fakes import OracleVolume
        img = mdist.render_views(OracleVolume(vol), cams, tf, replace(P, tfMode=1), mode=mode, device="cpu")
        if rank == 0:
            ret.put(img.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode,W,H,V", [("views", 20, 13, 2), ("tiles", 21, 27, 1), ("tiles", 16, 8, 2)])
def test_image_space_partition_world2(mode, W, H, V):
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, W, H, V, ret)) for r in range(2)]
    for p in procs:
        p.start()
    got = ret.get()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # single-process reference
    from mri_raytracer_b200 import orbit_views, OrbitalCamera
    from mri_raytracer_b200.synth import ramp_tf
    from oracle import oracle_torch as O
    from scenes import small_scene
    vol, _, P = small_scene(C=2, dims=(20, 18, 16), W=W, H=H, seed=7)
    tf = ramp_tf(32, sigma_scale=20.0, cutoff=0.1)
    cam = OrbitalCamera(initial_radius=float(np.linalg.norm(np.asarray(P.eye))), initial_phi=1.3, initial_theta=0.4)
    cam.set_fov_degrees(70.0)
    for v, c in enumerate(orbit_views(cam, V)):
        want = O.render(vol, replace(P.with_camera(c), tfMode=1), tf=tf).numpy()
        assert np.array_equal(got[v], want), (mode, v)


def _sl_worker(rank, world, port, W, H, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mri_raytracer_b200 import dist as mdist, RenderParams
        P = RenderParams(imageSize=(W, H), dims=(32, 32, 32), voxelSize=(0.05, 0.05, 0.05), volMin=(-0.8, -0.8, -0.8),
                         eye=(3.0, 2.0, 1.0), bgColor=(0.1, 0.2, 0.3), alphaMode=1)
        g = torch.Generator().manual_seed(100 + rank)
        partial = torch.rand(H, W, 4, generator=g)
        import dist_fakes
        mdist.composite_over = dist_fakes.composite_over_torch            # the product composite is CUDA-only
        img = mdist.render_sort_last(dist_fakes.OracleVolume(None, partial=partial), None, None, P, (2, 1, 1))
        if rank == 0:
            ret.put(img.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("W,H", [(16, 16), (21, 27)])
def test_sort_last_exchange_world2(W, H):
    """all_to_all of image strips + ordered composite + all_gather == compositing whole partials."""
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_sl_worker, args=(r, 2, port, W, H, ret)) for r in range(2)]
    for p in procs:
        p.start()
    got = ret.get()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    from mri_raytracer_b200 import dist as mdist, RenderParams
    P = RenderParams(imageSize=(W, H), dims=(32, 32, 32), voxelSize=(0.05, 0.05, 0.05), volMin=(-0.8, -0.8, -0.8),
                     eye=(3.0, 2.0, 1.0), bgColor=(0.1, 0.2, 0.3), alphaMode=1)
    parts = torch.stack([torch.rand(H, W, 4, generator=torch.Generator().manual_seed(100 + r)).reshape(-1, 4)
                         for r in range(2)])
    order = mdist.visibility_order(np.array(P.eye, dtype=np.float64), P, (2, 1, 1))
    assert order == [1, 0]                                     # eye on the +x side: the +x half is in front
    from dist_fakes import composite_over_torch
    want = composite_over_torch(parts, order, P.bgColor, 1).reshape(H, W, 4).numpy()
    assert np.array_equal(got, want)


def test_sort_last_helpers():
    from mri_raytracer_b200 import dist as mdist, RenderParams
    for R in (1, 2, 4, 8):
        g = mdist.shard_grid(R)
        assert g[0] * g[1] * g[2] == R
        dims = (33, 20, 17)
        covered = np.zeros([d - 1 for d in dims], dtype=np.int32)
        for r in range(R):
            lo, hi, _ = mdist.shard_box(dims, g, r)
            covered[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] += 1
        assert covered.min() == 1 and covered.max() == 1          # cells partitioned exactly once
    assert mdist.shard_grid(8) == (2, 2, 2)
    P = RenderParams(dims=(32, 32, 32), voxelSize=(0.05, 0.05, 0.05), volMin=(-0.8, -0.8, -0.8))
    order = mdist.visibility_order(np.array([3.0, 2.0, 1.0]), P, (2, 2, 2))
    assert sorted(order) == list(range(8))
    assert order[0] == 7 and order[-1] == 0                       # eye beyond the +x,+y,+z corner
    # ordered 'over' compositing is associative: compositing halves equals compositing all
    g = torch.Generator().manual_seed(0)
    parts = torch.rand(4, 50, 4, generator=g)
    from dist_fakes import composite_over_torch
    full = composite_over_torch(parts, [2, 0, 3, 1], (0.1, 0.2, 0.3))
    C = torch.zeros(50, 3); T = torch.ones(50)
    for k in [2, 0, 3, 1]:
        C = C + T[:, None] * parts[k, :, :3]; T = T * parts[k, :, 3]
    assert torch.allclose(full[:, :3], C + torch.tensor([0.1, 0.2, 0.3])) and torch.all(full[:, 3] == 1)


@pytest.mark.gpu
def test_composite_kernel_matches_torch(cuda):
    import ctypes as C
    from mri_raytracer_b200 import dist as mdist
    from mri_raytracer_b200._lib import lib, check
    g = torch.Generator().manual_seed(1)
    K, n = 8, 4099
    parts = torch.rand(K, n, 4, generator=g)
    order = torch.tensor([3, 1, 7, 0, 2, 6, 5, 4], dtype=torch.int32)
    from dist_fakes import composite_over_torch
    want = composite_over_torch(parts, order.tolist(), (0.05, 0.1, 0.2), alpha_mode=1)
    p, o = parts.cuda().contiguous(), order.cuda()
    bg = torch.tensor([0.05, 0.1, 0.2]).numpy()
    out = torch.empty(n, 4, device="cuda")
    check(lib().mrt_composite_over(p.data_ptr(), K, o.data_ptr(), n, bg.ctypes.data, 1, out.data_ptr(),
                                   torch.cuda.current_stream().cuda_stream))
    assert (out.cpu() - want).abs().max() <= 1e-6


# ----------------------------------------------------------------------------- differentiable, data-parallel tiles
def _grad_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mri_raytracer_b200 import dist as mdist
        from mri_raytracer_b200.synth import ramp_tf
        from scenes import small_scene
        vol, _, P = small_scene(C=2, dims=(14, 12, 10), W=19, H=13, seed=5)
        tf = ramp_tf(16, sigma_scale=20.0, cutoff=0.1)
        v = vol.clone().requires_grad_(True); t = tf.clone().requires_grad_(True)
        import dist_fakes
        from mri_raytracer_b200 import api
        api.render = dist_fakes.oracle_render_part                         # stands in for the CUDA renderer
        img = mdist.render_differentiable(v, None, t, P)
        target = torch.linspace(0, 1, img.numel()).reshape(img.shape)
        ((img - target) ** 2).mean().backward()
        mdist.allreduce_gradients([v, t])
        if rank == 0:
            ret.put((img.detach().numpy(), v.grad.numpy(), t.grad.numpy()))
    finally:
        dist.destroy_process_group()


def test_differentiable_tiles_world2_gradients_match_single_process():
    from oracle import oracle_torch as O
    from mri_raytracer_b200.synth import ramp_tf
    from scenes import small_scene
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    img, gv, gt = ret.get()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    vol, _, P = small_scene(C=2, dims=(14, 12, 10), W=19, H=13, seed=5)
    tf = ramp_tf(16, sigma_scale=20.0, cutoff=0.1)
    v = vol.clone().requires_grad_(True); t = tf.clone().requires_grad_(True)
    ref = O.render(v, replace(P, tfMode=1), tf=t)
    target = torch.linspace(0, 1, ref.numel()).reshape(ref.shape)
    ((ref - target) ** 2).mean().backward()
    assert np.array_equal(img, ref.detach().numpy())
    assert np.abs(gv - v.grad.numpy()).max() <= 1e-6 * max(1.0, float(v.grad.abs().max()))
    assert np.abs(gt - t.grad.numpy()).max() <= 1e-5 * max(1.0, float(t.grad.abs().max()))


def _sl_grad_worker(rank, world, port, W, H, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mri_raytracer_b200 import api, dist as mdist, RenderParams
        P = RenderParams(imageSize=(W, H), dims=(32, 32, 32), voxelSize=(0.05, 0.05, 0.05), volMin=(-0.8, -0.8, -0.8),
                         eye=(3.0, 2.0, 1.0), bgColor=(0.1, 0.2, 0.3), alphaMode=1)
        w = torch.randn(H, W, 4, generator=torch.Generator().manual_seed(200 + rank)).requires_grad_(True)
        # stands in for the CUDA shard renderer: a differentiable partial that depends on this rank's parameters only
        api.render_shard = lambda sub, shard, gdims, cam, tf, Pp, storage=None: torch.sigmoid(sub)
        strip, row0 = mdist.render_sort_last_differentiable(w, None, None, P, (2, 1, 1))
        rows = strip.shape[0]
        valid = max(0, min(rows, H - row0))
        wmap = torch.linspace(0.5, 1.5, H * W * 4).reshape(H, W, 4)
        loss = (strip[:valid] * wmap[row0:row0 + valid]).sum()
        loss.backward()
        ret.put((rank, float(loss), w.grad.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("W,H", [(16, 16), (21, 27)])
def test_differentiable_sort_last_world2(W, H):
    """Differentiable all_to_all of strips + tensor-op composite: the ranks' strip losses add up to the
    whole-frame loss and every rank ends up with d(total loss)/d(its own partial's parameters)."""
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_sl_grad_worker, args=(r, 2, port, W, H, ret)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict()
    for _ in range(2):
        r, l, g = ret.get()
        got[r] = (l, g)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    from mri_raytracer_b200 import dist as mdist, RenderParams
    from dist_fakes import composite_over_torch
    P = RenderParams(imageSize=(W, H), dims=(32, 32, 32), voxelSize=(0.05, 0.05, 0.05), volMin=(-0.8, -0.8, -0.8),
                     eye=(3.0, 2.0, 1.0), bgColor=(0.1, 0.2, 0.3), alphaMode=1)
    ws = [torch.randn(H, W, 4, generator=torch.Generator().manual_seed(200 + r)).requires_grad_(True) for r in range(2)]
    order = mdist.visibility_order(np.array(P.eye, dtype=np.float64), P, (2, 1, 1))
    img = composite_over_torch(torch.stack([torch.sigmoid(w).reshape(-1, 4) for w in ws]), order, P.bgColor, 1).reshape(H, W, 4)
    wmap = torch.linspace(0.5, 1.5, H * W * 4).reshape(H, W, 4)
    loss = (img * wmap).sum()
    loss.backward()
    assert abs(got[0][0] + got[1][0] - float(loss)) <= 1e-4 * abs(float(loss))
    for r in range(2):
        assert np.abs(got[r][1] - ws[r].grad.numpy()).max() <= 1e-5


def test_composite_over_differentiable_matches_the_loop():
    from mri_raytracer_b200 import dist as mdist
    from dist_fakes import composite_over_torch
    parts = torch.rand(5, 37, 4, generator=torch.Generator().manual_seed(3), dtype=torch.float64)
    order = [3, 0, 4, 1, 2]
    for am in (0, 1):
        a = mdist.composite_over_differentiable(parts, order, (0.1, 0.2, 0.3), am)
        b = composite_over_torch(parts, order, (0.1, 0.2, 0.3), am)
        assert torch.allclose(a, b, atol=1e-12)


# ----------------------------------------------------------------------------- PeerFramebuffer fallback (no symmetric memory)
def _fb_worker(rank, world, port, partition, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import warnings
        from mri_raytracer_b200 import dist as mdist, orbit_views, OrbitalCamera
        from mri_raytracer_b200.synth import ramp_tf
        from scenes import small_scene
        from dist_fakes import OracleVolume
        W, H, V = 19, 13, 4
        vol, _, P = small_scene(C=1, dims=(16, 14, 12), W=W, H=H, seed=11)
        tf = ramp_tf(16, sigma_scale=20.0, cutoff=0.1)
        cam = OrbitalCamera(initial_radius=float(np.linalg.norm(np.asarray(P.eye))), initial_phi=1.2, initial_theta=0.1)
        cam.set_fov_degrees(70.0)
        cams = orbit_views(cam, V)
        try:
            mdist.PeerFramebuffer(V, H, W, "cpu", partition=partition)
            loud = False
        except RuntimeError as e:                            # no symmetric memory on CPU: must be loud by default
            loud = "symmetric" in str(e)
        with warnings.catch_warnings(record=True) as rec:
            warnings.simplefilter("always")
            fb = mdist.PeerFramebuffer(V, H, W, "cpu", partition=partition, allow_nccl_fallback=True)
        assert loud and not fb.p2p and any("falling back" in str(w.message) for w in rec)
        assert list(fb.owned_views()) == list(range(V))
        fb.render(OracleVolume(vol), cams, tf, P)
        frames = fb.finish()
        if rank == 0:
            ret.put(frames.numpy().copy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("partition", ["views", "tiles"])
def test_peer_framebuffer_nccl_fallback_is_loud_and_exact_world2(partition):
    from oracle import oracle_torch as O
    from mri_raytracer_b200 import orbit_views, OrbitalCamera
    from mri_raytracer_b200.synth import ramp_tf
    from scenes import small_scene
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_fb_worker, args=(r, 2, port, partition, ret)) for r in range(2)]
    for p in procs:
        p.start()
    got = ret.get()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    W, H = 19, 13
    vol, _, P = small_scene(C=1, dims=(16, 14, 12), W=W, H=H, seed=11)
    tf = ramp_tf(16, sigma_scale=20.0, cutoff=0.1)
    cam = OrbitalCamera(initial_radius=float(np.linalg.norm(np.asarray(P.eye))), initial_phi=1.2, initial_theta=0.1)
    cam.set_fov_degrees(70.0)
    cams = orbit_views(cam, 4)
    assert got.shape == (4, H, W, 4)
    for v, c in enumerate(cams):
        ref = O.render(vol, replace(P.with_camera(c), tfMode=1), tf=tf).numpy()
        assert np.array_equal(got[v], ref), f"view {v}"


def test_framebuffer_ownership_maps():
    """Integer maps of the distributed framebuffer: striped owners partition the views, every view has
    exactly one (owner, slot); interleaved tile rows partition the tile rows."""
    from mri_raytracer_b200 import tiles
    for V in (1, 3, 8, 13, 64):
        for R in (1, 2, 4, 8):
            per = (V + R - 1) // R
            seen = set()
            for v in range(V):
                o, slot = v // per, v % per
                assert 0 <= o < R and 0 <= slot < per
                seen.add((o, slot))
            assert len(seen) == V
    for H in (1, 8, 13, 64, 1024, 2050):
        ty = tiles.tiles_y(H)
        for R in (1, 2, 3, 4, 8):
            rows = sorted(t for r in range(R) for t in tiles.interleaved_rows(ty, r, R))
            assert rows == list(range(ty))
