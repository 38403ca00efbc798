"""Forward parity: CUDA kernels (through the C ABI) vs the CPU-torch oracle — max-abs 1e-4."""
import math
from dataclasses import replace

import numpy as np
import pytest
import torch

from mri_raytracer_b200 import api, RenderParams
from mri_raytracer_b200.synth import ramp_tf
from scenes import small_scene
from parity import check_forward, O

pytestmark = pytest.mark.gpu


def _render_both(vol, P, tf, lab=None, prd=None, dev="cuda", fold=True):
    V = api.Volume(vol.to(dev), labels=None if lab is None else lab.to(dev), preds=None if prd is None else prd.to(dev),
                   fold=fold)
    tfd = None if tf is None else tf.to(dev)
    img = api.render(V, None, tfd, P)
    img2, T, counts = api.render_aux(V, None, tfd, P)
    torch.cuda.synchronize()
    return img, img2, T, counts


@pytest.mark.parametrize("C", [1, 2, 3, 4])
@pytest.mark.parametrize("use_tf", [False, True])
@pytest.mark.parametrize("fold", [False, True])
def test_forward_matches_oracle(cuda, C, use_tf, fold):
    if fold and C == 1:
        pytest.skip("folding is the identity for one modality")
    vol, _, P = small_scene(C=C, dims=(40, 36, 28), W=72, H=56, seed=C)
    tf = ramp_tf(64) if use_tf else None
    P = replace(P, intensityAlpha=25.0, volWeight=(1.0, 0.5, 2.0, 0.75), volEnabled=(1, 1, 0 if C == 4 else 1, 1))
    img, img2, T, counts = _render_both(vol, P, tf, fold=fold)
    assert torch.equal(img, img2), "fast and counting kernel variants must agree bit-for-bit"
    stats, ref, aux = check_forward(img, counts, vol, replace(P, tfMode=int(use_tf)), tf)
    assert stats["max_abs"] <= 1e-4
    assert stats["samples_taken"] == int(aux["n_taken"].sum()) or stats["n_flip"] > 0
    assert (T.cpu() - aux["T"]).abs().max() <= 1e-4 or stats["n_flip"] > 0


@pytest.mark.parametrize("ortho", [False, True])
def test_skipping_is_exact(cuda, ortho):
    vol, _, P = small_scene(C=4, dims=(64, 48, 40), W=96, H=80, seed=3, ortho=ortho)
    tf = ramp_tf(256)
    a, _, _, ca = _render_both(vol, replace(P, skipEmpty=1), tf)
    b, _, _, cb = _render_both(vol, replace(P, skipEmpty=0), tf)
    assert torch.equal(a, b), "empty-space skipping changed the image"
    ca, cb = ca.cpu(), cb.cpu()
    assert torch.equal(ca[..., :2], cb[..., :2])
    assert int(ca[..., 2].sum()) < int(cb[..., 2].sum()), "skipping skipped nothing"
    check_forward(a, ca, vol, replace(P, tfMode=1), tf)


def test_labels_overlay(cuda):
    vol, lab, P = small_scene(C=4, dims=(48, 40, 32), W=64, H=64, seed=5, labels=True)
    assert int((lab > 0).sum()) > 0
    prd = torch.roll(lab, shifts=2, dims=2).contiguous()
    # sample label positions avoid exact .5 ties with overwhelming probability
    P = replace(P, showSeg=1, showPred=1, intensityAlpha=5.0)
    img, img2, T, counts = _render_both(vol, P, None, lab, prd)
    assert torch.equal(img, img2)
    stats, _, _ = check_forward(img, counts, vol, P, None, lab.long(), prd.long())
    assert stats["max_abs"] <= 1e-4


def test_accumulate_mode_and_gamma(cuda):
    vol, _, P = small_scene(C=1, dims=(32, 32, 32), W=40, H=40, seed=2)
    P = replace(P, tMode="accumulate", gamma=1.7, intensityAlpha=10.0)
    V = api.Volume(vol.cuda())
    img = api.render(V, None, None, P).cpu()
    ref = O.render(vol, P)
    # powf differs between libm and CUDA by a few ulp; still far inside the tolerance
    assert (img - ref).abs().max() <= 1e-4


def test_miss_inside_and_degenerate_rays(cuda):
    vol, _, P = small_scene(C=1, dims=(24, 24, 24), W=33, H=17, seed=1)
    tf = ramp_tf(32)
    cases = [
        replace(P, eye=(5.0, 5.0, 5.0), W=(1.0, 0.0, 0.0)),                       # every ray misses
        replace(P, eye=(0.0, 0.0, 0.0)),                                           # eye inside the box
        replace(P, eye=(0.0, 0.0, -3.0), U=(1, 0, 0), V=(0, 1, 0), W=(0, 0, 1), ortho=1, orthoHalfHeight=0.8),  # d.x=d.y=0
        replace(P, nearT=1.9, farT=2.3),                                           # near/far clamp
        replace(P, bgColor=(0.2, 0.3, 0.4), alphaMode=1),
        replace(P, maxSteps=7),
    ]
    for Pc in cases:
        img, img2, T, counts = _render_both(vol, Pc, tf)
        stats, _, _ = check_forward(img, counts, vol, replace(Pc, tfMode=1), tf)
        assert stats["max_abs"] <= 1e-4


def test_ragged_image_sizes_and_tile_ranges(cuda):
    vol, _, P0 = small_scene(C=2, dims=(24, 20, 16), W=8, H=8, seed=4)
    V = api.Volume(vol.cuda())
    for (W, H) in [(1, 1), (7, 3), (9, 17), (70, 33)]:
        P = replace(P0, imageSize=(W, H))
        full = api.render(V, None, None, P)
        ref = O.render(vol, P)
        assert (full.cpu() - ref).abs().max() <= 1e-4
        # render in two disjoint tile ranges: union must equal the full frame bit-for-bit
        from mri_raytracer_b200 import tiles
        nt = tiles.tile_count(W, H)
        out = torch.full((H, W, 4), -7.0, device="cuda")
        for r in range(2):
            V.forward(replace(P, tfMode=0), None, out=out, tile_range=tiles.rank_tile_range(nt, r, 2))
        assert torch.equal(out, full)


@pytest.mark.parametrize("ortho", [False, True])
def test_batched_views_equal_single_frames(cuda, ortho):
    """mrt_render_forward_batch: view v of one batched launch == the single-frame call, bit for bit
    (incl. a batch larger than one launch's 64 views, tile ranges and the per-ray counters)."""
    from mri_raytracer_b200 import Camera, OrbitalCamera, orbit_views, tiles
    vol, lab, P = small_scene(C=4, dims=(40, 36, 28), W=41, H=27, seed=8, labels=True, ortho=ortho)
    tf = ramp_tf(64).cuda()
    V = api.Volume(vol.cuda(), labels=lab.cuda())
    P = replace(P, showSeg=1)
    cam = V.frame_camera(OrbitalCamera(initial_radius=3.0, initial_theta=0.3, initial_phi=1.2))
    cam.set_fov_degrees(70.0)
    cams = orbit_views(cam, 67, ortho=ortho)
    batch = api.render_views(V, cams, tf, P)
    assert batch.shape == (67, 27, 41, 4)
    for v in (0, 1, 31, 63, 64, 66):
        single = api.render(V, cams[v], tf, P)
        assert torch.equal(batch[v], single), f"view {v} differs from the single-frame render"
    ref = O.render(vol, replace(P.with_camera(cams[5]), tfMode=1), tf=tf.cpu(), labels=lab.long())
    assert (batch[5].cpu() - ref).abs().max() <= 1e-4
    # two disjoint tile ranges of the batch == the whole batch
    nt = tiles.tile_count(41, 27)
    out = torch.full((3, 27, 41, 4), -7.0, device="cuda")
    for r in range(2):
        api.render_views(V, cams[:3], tf, P, out=out, tile_range=tiles.rank_tile_range(nt, r, 2))
    assert torch.equal(out, batch[:3])
    with pytest.raises(ValueError):
        api.render_views(V, [cams[0], replace(cams[1], fovY=0.5)], tf, P)


@pytest.mark.parametrize("C,dims", [(4, (250, 37, 21)), (3, (40, 33, 17)), (2, (8, 8, 8)), (1, (505, 10, 9))])
def test_fused_fold_occupancy_equals_two_pass(cuda, C, dims):
    """mrt_fold_volume_occupancy_f32 == mrt_fold_volume_f32 + mrt_build_occupancy, bit for bit
    (ragged dims, more than one 248-column chunk, every channel count)."""
    X, Y, Z = dims
    g = torch.Generator().manual_seed(C)
    vol = (torch.rand((C, Z, Y, X), generator=g) - 0.3).cuda()
    P = RenderParams(imageSize=(8, 8), dims=dims, volWeight=(1.0, -0.5, 2.0, 0.75), volEnabled=(1, 1, 0 if C == 4 else 1, 1))
    folded_a = api.fold_volume(vol, P)
    mm_a = api.build_occupancy(folded_a, 1, dims)
    folded_b, mm_b = api.fold_volume_occupancy(vol, P)
    assert torch.equal(api.unpack_volume(folded_a, 1, dims), api.unpack_volume(folded_b, 1, dims))
    assert torch.equal(mm_a, mm_b)
    # independent check of the brick ranges: brick b covers voxels [8b, 8b+8] per axis
    f = api.unpack_volume(folded_b, 1, dims)[0].cpu()
    nbx, nby = (X + 7) // 8, (Y + 7) // 8
    for b in (0, mm_b.shape[0] // 2, mm_b.shape[0] - 1):
        bx, by, bz = b % nbx, (b // nbx) % nby, b // (nbx * nby)
        blk = f[8 * bz:8 * bz + 9, 8 * by:8 * by + 9, 8 * bx:8 * bx + 9]
        assert float(mm_b[b, 0, 0]) == float(blk.min()) and float(mm_b[b, 0, 1]) == float(blk.max())


@pytest.mark.parametrize("ortho,use_tf", [(False, True), (True, False)])
def test_fp16_volume_matches_oracle_on_rounded_values(cuda, ortho, use_tf):
    """fp16 storage (BASELINE config 5): corners are widened to fp32 on load, interpolation stays
    fp32 — so the image equals the oracle's on the fp16-rounded volume to the usual 1e-4, the
    per-ray sample counts stay bit-exact, and skipping stays exact."""
    vol, _, P = small_scene(C=1, dims=(67, 36, 29), W=72, H=56, seed=11, ortho=ortho)
    tf = ramp_tf(64) if use_tf else None
    P = replace(P, intensityAlpha=25.0)
    volh = vol.half()
    V = api.Volume(volh.cuda())
    assert V.half and V.packed.dtype == torch.float16
    assert torch.equal(api.unpack_volume_f16(V.packed, V.dims).cpu(), volh)
    tfd = None if tf is None else tf.cuda()
    img, T, counts = api.render_aux(V, None, tfd, P)
    stats, ref, aux = check_forward(img, counts, volh.float(), replace(P, tfMode=int(use_tf)), tf)
    assert stats["max_abs"] <= 1e-4
    fast = api.render(V, None, tfd, P)
    assert torch.equal(fast, img)
    dense = api.render(V, None, tfd, replace(P, skipEmpty=0))
    assert torch.equal(dense, img), "empty-space skipping changed an fp16 render"
    # fp32 render of the same (rounded) values: same arithmetic after the load
    img32 = api.render(api.Volume(volh.float().cuda()), None, tfd, P)
    assert torch.equal(img32, img)
    with pytest.raises(ValueError):
        api.Volume(torch.zeros((2, 8, 8, 8), dtype=torch.float16, device="cuda"))


@pytest.mark.parametrize("W,H,alpha,ortho", [(200, 136, 0, False), (41, 27, 1, False), (96, 80, 0, True)])
def test_sparse_batch_plus_fill_equals_dense_batch(cuda, W, H, alpha, ortho):
    """mrt_view_spans + mrt_render_forward_batch_sparse + mrt_fill_outside_spans (the sparse
    framebuffer gather, here into a local buffer full of NaNs) == mrt_render_forward_batch, bit for
    bit; some tiles really are skipped; a camera inside the volume degrades to the full frame."""
    from mri_raytracer_b200 import Camera, OrbitalCamera, orbit_views
    vol, _, P = small_scene(C=4, dims=(40, 36, 28), W=W, H=H, seed=8, ortho=ortho)
    P = replace(P, bgColor=(0.1, 0.2, 0.3), alphaMode=alpha)
    tf = ramp_tf(64).cuda()
    V = api.Volume(vol.cuda())
    cam = V.frame_camera(OrbitalCamera(initial_radius=3.0, initial_theta=0.3, initial_phi=1.2))
    cam.radius *= 2.0
    cam.set_fov_degrees(70.0)
    cams = orbit_views(cam, 5, ortho=ortho)
    if not ortho:
        cams[4] = replace(cams[4], eye=np.asarray([0.02, 0.01, -0.03], dtype=np.float32))    # inside the box
    dense = api.render_views(V, cams, tf, P)
    out = torch.full((5, H, W, 4), float("nan"), device="cuda")
    Pm = replace(P, tfMode=1)
    packed, Cn, Pe, bits = V.sparse_plan(Pm, cams, tf)
    spans = api.view_spans(Pe, cams, Cn, bits)
    api.render_forward_batch_sparse(Pe, cams, packed, Cn, tf, bits, out.data_ptr(), spans)
    r = spans.cpu()
    assert bool(torch.isnan(out[0]).any()), "nothing was skipped: the test scene must have empty borders"
    skipped = float(torch.isnan(out[0, ..., 0]).float().mean())
    assert skipped > 0.3, skipped
    if not ortho:
        assert bool((r[4, :, 0] == 0).all()) and bool((r[4, :, 1] == W - 1).all()) and not bool(torch.isnan(out[4]).any())
    api.fill_outside_spans(Pe, spans, out)
    assert torch.equal(out, dense)
    assert V.sparse_plan(replace(Pm, gamma=1.5), cams, tf) is None
    # the same buffer reused by the next batch (the camera moved): the DELTA fill only writes the tiles the
    # previous spans covered and the new ones do not, and the frame is still the dense one bit for bit
    cam.theta += 0.35; cam.phi -= 0.1
    cams2 = orbit_views(cam, 5, ortho=ortho)
    dense2 = api.render_views(V, cams2, tf, P)
    spans2 = api.view_spans(Pe, cams2, Cn, bits)
    assert not torch.equal(spans2, spans)
    api.render_forward_batch_sparse(Pe, cams2, packed, Cn, tf, bits, out.data_ptr(), spans2)
    api.fill_outside_spans(Pe, spans2, out, prev_spans=spans)
    assert torch.equal(out, dense2)


def test_refold_when_weights_change(cuda):
    vol, _, P = small_scene(C=4, dims=(32, 28, 24), W=40, H=32, seed=9)
    V = api.Volume(vol.cuda())
    a = api.render(V, None, None, replace(P, intensityAlpha=8.0)).cpu()
    P2 = replace(P, intensityAlpha=8.0, volWeight=(0.2, 3.0, 1.0, 0.0), volEnabled=(1, 1, 0, 1))
    b = api.render(V, None, None, P2).cpu()
    assert (a - b).abs().max() > 1e-3
    assert (b - O.render(vol, P2)).abs().max() <= 1e-4
    assert (api.render(V, None, None, replace(P, intensityAlpha=8.0)).cpu() - a).abs().max() == 0.0


def test_render_host_entry_point(cuda):
    vol, _, P = small_scene(C=4, dims=(32, 28, 24), W=40, H=32, seed=6)
    tf = ramp_tf(128)
    V = api.Volume(vol.cuda(), fold=False)
    dev_img = api.render(V, None, tf.cuda(), P).cpu()
    host_img = api.render_host(vol.numpy(), P, tf.numpy())
    assert np.array_equal(host_img, dev_img.numpy())


@pytest.mark.parametrize("C,use_tf", [(4, True), (1, False)])
def test_host_pipeline_equals_device_renders(cuda, C, use_tf):
    """mrt_host_pipeline_*: several queued steps (different volumes / cameras per step, more steps
    than slots) give bit for bit the frames of the device-side API."""
    from mri_raytracer_b200 import OrbitalCamera, orbit_views
    dims, W, H, V = (40, 36, 28), 41, 27, 3
    _, _, P = small_scene(C=C, dims=dims, W=W, H=H, seed=1)
    P = replace(P, intensityAlpha=9.0, volWeight=(1.0, 0.5, 2.0, 0.75))
    tf = ramp_tf(64) if use_tf else None
    pipe = api.HostPipeline(C, dims, (W, H), max_views=V, max_tf=64, depth=2)
    vols, outs, tickets, cams_per_step = [], [], [], []
    for step in range(5):
        vol, _, _ = small_scene(C=C, dims=dims, W=W, H=H, seed=20 + step)
        vh = vol.pin_memory()
        oh = torch.empty((V, H, W, 4)).pin_memory()
        Vd = api.Volume(vol.cuda())
        cam = Vd.frame_camera(OrbitalCamera(initial_radius=3.0, initial_theta=0.4 * step, initial_phi=1.3))
        cam.set_fov_degrees(70.0)
        cams = orbit_views(cam, V)
        tickets.append(pipe.submit(vh.numpy(), cams, P, None if tf is None else tf.numpy(), oh.numpy()))
        vols.append(Vd); outs.append(oh); cams_per_step.append(cams)
    for step in (4, 0, 2, 1, 3):
        pipe.wait(tickets[step])
        ref = api.render_views(vols[step], cams_per_step[step], None if tf is None else tf.cuda(), P)
        assert torch.equal(outs[step], ref.cpu()), f"step {step} differs"
    with pytest.raises(ValueError):
        pipe.submit(np.zeros((C, 2, 2, 2), np.float32), cams, P, None, outs[0].numpy())
    pipe.close()


@pytest.mark.parametrize("pinned", [True, False])
def test_host_pipeline_resident_volume_sparse_download_and_damage_tracking(cuda, pinned):
    """Resident volume (set_volume once, the reference's load-time upload) + sparse frame download:
    only each view's bounding rectangle of non-background tiles crosses PCIe, the pipeline keeps the
    rest of the host frame at the background.  Output arrays are REUSED across steps whose cameras
    move (so the footprint moves, shrinks, vanishes and the background colour changes): the host
    frames must equal the device-side renders bit for bit every time.  Page-locked outputs are written
    by the march kernel itself (zero-copy stores of the in-span tiles); pageable ones take the staged
    path (one strided copy of each view's bounding rectangle)."""
    from mri_raytracer_b200 import OrbitalCamera, orbit_views
    dims, W, H, V = (40, 36, 28), 75, 53, 2
    vol, _, P = small_scene(C=4, dims=dims, W=W, H=H, seed=3)
    tf = ramp_tf(64)
    Vd = api.Volume(vol.cuda())
    pipe = api.HostPipeline(4, dims, (W, H), max_views=V, max_tf=64, depth=3)
    vh = vol.pin_memory()
    pipe.set_volume(vh.numpy())
    outs = [torch.full((V, H, W, 4), 7.0) for _ in range(2)]                     # garbage: must be overwritten
    if pinned:
        outs = [o.pin_memory() for o in outs]
    total_d2h = 0
    for step in range(7):
        cam = Vd.frame_camera(OrbitalCamera(initial_radius=3.0, initial_theta=0.7 * step, initial_phi=1.3))
        cam.set_fov_degrees(70.0)
        cam.radius *= (1.0, 1.6, 0.8, 2.5, 1.0, 1.0, 1.2)[step]                   # the footprint grows and shrinks
        if step == 4:
            cam.target = cam.target + np.float32(50.0)                            # looks away from the volume
        cams = orbit_views(cam, V)
        Ps = replace(P, bgColor=(0.1, 0.2, 0.3) if step >= 5 else (0.0, 0.0, 0.0),
                     volWeight=(1.0, 0.5 + 0.1 * step, 2.0, 0.75))
        out = outs[step % 2]
        pipe.wait(pipe.submit(None, cams, Ps, tf.numpy(), out.numpy()))
        up, down, fill = pipe.last_bytes()
        total_d2h += down
        assert up < 4096 + 64 * 16                                                # cameras, params, TF: no volume
        ref = api.render_views(Vd, cams, tf.cuda(), Ps)
        assert torch.equal(out, ref.cpu()), f"step {step} differs"
    assert total_d2h < 7 * V * H * W * 16                                         # less than dense downloads
    pipe.forget(outs[0].numpy())
    pipe.close()


@pytest.mark.parametrize("R", [2, 3, 8])
def test_interleaved_tile_rows_scattered_to_per_view_frames_equal_the_batch(cuda, R):
    """mrt_render_forward_batch_scatter: the image-space tile partition with the gather fused into the
    march.  R 'ranks' (emulated one after the other on this GPU) each render tile rows ty % R == r of
    every view and store them through a per-view pointer table into frames that live in two separate
    allocations (standing for two owner GPUs); the owners fill the background outside the spans.
    The union equals render_views bit for bit."""
    from mri_raytracer_b200 import OrbitalCamera, orbit_views, tiles
    dims, W, H, V = (48, 40, 36), 83, 61, 5
    vol, _, P = small_scene(C=4, dims=dims, W=W, H=H, seed=5)
    P = replace(P, tfMode=1, bgColor=(0.2, 0.1, 0.3))
    tf = ramp_tf(64).cuda()
    Vd = api.Volume(vol.cuda())
    cam = Vd.frame_camera(OrbitalCamera(initial_radius=3.0, initial_theta=0.3, initial_phi=1.2))
    cam.set_fov_degrees(70.0)
    cams = orbit_views(cam, V)
    ref = api.render_views(Vd, cams, tf, P)
    packed, Cn, Pe, bits = Vd.sparse_plan(P, cams, tf)
    owner_a = torch.full((3, H, W, 4), -1.0, device="cuda")          # views 0..2
    owner_b = torch.full((2, H, W, 4), -1.0, device="cuda")          # views 3..4
    frame = H * W * 16
    ptrs = torch.tensor([owner_a.data_ptr() + i * frame for i in range(3)] + [owner_b.data_ptr() + i * frame for i in range(2)],
                        dtype=torch.int64, device="cuda")
    spans = api.view_spans(Pe, cams, Cn, bits)
    for r in range(R):
        api.render_forward_batch_scatter(Pe, cams, packed, Cn, tf, bits, ptrs, spans, store_outside=False, row_mod=R, row_rem=r)
    api.fill_outside_spans(Pe, spans[:3].contiguous(), owner_a)
    api.fill_outside_spans(Pe, spans[3:].contiguous(), owner_b)
    assert torch.equal(torch.cat([owner_a, owner_b]), ref)
    assert sorted(t for r in range(R) for t in tiles.interleaved_rows(tiles.tiles_y(H), r, R)) == list(range(tiles.tiles_y(H)))


def test_peer_framebuffer_single_rank_double_buffer(cuda):
    """dist.PeerFramebuffer degenerates to a local double-buffered framebuffer on one rank: the same
    scatter + owner-fill code path as on N GPUs; frames of batch b stay intact while batch b+1 is
    rendered into the other buffer."""
    from mri_raytracer_b200 import dist as mdist, OrbitalCamera, orbit_views
    dims, W, H, V = (40, 36, 28), 64, 48, 3
    vol, _, P = small_scene(C=1, dims=dims, W=W, H=H, seed=9)
    tf = ramp_tf(64).cuda()
    Vd = api.Volume(vol.cuda())
    fb = mdist.PeerFramebuffer(V, H, W, "cuda")
    refs, got = [], []
    for b in range(3):
        cam = Vd.frame_camera(OrbitalCamera(initial_radius=3.0, initial_theta=0.5 * b, initial_phi=1.2))
        cam.set_fov_degrees(70.0)
        cams = orbit_views(cam, V)
        fb.render(Vd, cams, tf, P)
        got.append(fb.finish())
        refs.append(api.render_views(Vd, cams, tf, P))
        if b >= 1:
            assert torch.equal(got[b - 1], refs[b - 1])               # previous batch untouched by this one
    assert torch.equal(got[2], refs[2])


@pytest.mark.parametrize("ortho,use_tf", [(False, True), (True, True), (False, False)])
def test_u8_volume_through_the_main_marcher(cuda, ortho, use_tf):
    """1 byte per voxel (the reference's single-volume app stores bytes and reads value = byte/255,
    volume_render.slang:33-38; scripts/volumeRendering/app.py:145-158) through the SAME marcher as
    fp32 volumes — occupancy skipping, TF, ERT — gathering 8 B per sample: per-ray counts bit-exact and
    the image within 1e-4 of the oracle on the byte/255 volume (early-termination tie rule), within
    5e-6 of the fp32 kernel on that volume away from such ties, and bit-identical with skipping on and off."""
    vol, _, P = small_scene(C=1, dims=(52, 44, 36), W=72, H=56, seed=13, ortho=ortho)
    P = replace(P, tfMode=int(use_tf), intensityAlpha=8.0, bgColor=(0.05, 0.0, 0.1))
    tf = ramp_tf(64, sigma_scale=20.0, cutoff=0.1) if use_tf else None
    u8 = (vol * 255.0).round().clamp(0, 255).to(torch.uint8)
    as_f32 = u8.to(torch.float32) / 255.0
    tfd = None if tf is None else tf.cuda()
    V8 = api.Volume(u8.cuda())
    img, T, counts = api.render_aux(V8, None, tfd, P)
    stats, _, _ = check_forward(img, counts, as_f32, P, tf_cpu=tf)
    assert stats["max_abs"] <= 1e-4
    ref32 = api.render(api.Volume(as_f32.cuda()), None, tfd, P)
    d32 = (api.render(V8, None, tfd, P) - ref32).abs().amax(dim=-1)
    # (the /255 is applied after the interpolation instead of per corner: ~1e-7 per sample; a ray whose
    # transmittance lands within that of the early-termination threshold may take one sample more or less)
    assert float((d32 > 5e-6).float().mean()) <= 2e-3 and float(d32.max()) <= 2e-3
    assert torch.equal(api.render(V8, None, tfd, P), api.render(V8, None, tfd, replace(P, skipEmpty=0)))
    if use_tf:      # (the window/level TF with lo == 0 is "val > 0": the classifier's safety margin keeps air bricks active)
        assert stats["samples_evaluated"] < stats["samples_taken"]      # skipping really happens
    with pytest.raises(api._lib.MrtError):
        api.render_backward(replace(P, volDtype=2), V8.packed, 1, tfd, None, None, img, torch.ones_like(img))


@pytest.mark.parametrize("C,ortho,use_tf", [(1, False, True), (4, False, True), (2, True, False), (1, True, True)])
def test_quad_layout_is_bit_identical_to_the_scalar_layout(cuda, C, ortho, use_tf):
    """The 16 B/voxel quad sampler layout (mrt_pack_volume_quad, volDtype 3: two 16-byte loads per
    sample) holds the same eight corners as the scalar layout: images, transmittance and per-ray
    counts must be bit-identical, with skipping on and off, single frames and batches."""
    vol, _, P = small_scene(C=C, dims=(52, 44, 37), W=72, H=56, seed=21, ortho=ortho)
    P = replace(P, tfMode=int(use_tf), intensityAlpha=8.0, bgColor=(0.05, 0.0, 0.1))
    tfd = ramp_tf(64, sigma_scale=20.0, cutoff=0.1).cuda() if use_tf else None
    Vq, Vs = api.Volume(vol.cuda(), quad=True), api.Volume(vol.cuda(), quad=False)
    assert Vq.quad and not Vs.quad
    _, _, Pq = Vq.prepared(P)
    assert Pq.volDtype == 3
    for skip in (1, 0):
        a = api.render_aux(Vq, None, tfd, replace(P, skipEmpty=skip))
        b = api.render_aux(Vs, None, tfd, replace(P, skipEmpty=skip))
        for x, y in zip(a, b):
            assert torch.equal(x, y)
    assert torch.equal(api.render(Vq, None, tfd, P), api.render(Vs, None, tfd, P))
    img, _, counts = api.render_aux(Vq, None, tfd, P)
    stats, _, _ = check_forward(img, counts, vol, P, tf_cpu=None if tfd is None else tfd.cpu())
    assert stats["max_abs"] <= 1e-4
    # the far faces (x = X-1 / y = Y-1 are only ever read as the +1 neighbour) and a changed fold
    if C > 1:
        P2 = replace(P, volWeight=(0.2, 1.0, 0.5, 2.0))
        assert torch.equal(api.render(Vq, None, tfd, P2), api.render(Vs, None, tfd, P2))


def test_fp16_quad_layout_is_bit_identical_incl_shards(cuda):
    """volDtype 4 (quads of fp16 voxels, 8 B per element) against volDtype 1 on the same fp16 volume:
    whole volume and a sort-last sub-box (BASELINE config 5's storage), skipping on and off."""
    vol, _, P = small_scene(C=1, dims=(52, 44, 37), W=72, H=56, seed=22)
    P = replace(P, tfMode=1, bgColor=(0.05, 0.0, 0.1), ertThreshold=1e-4)
    tfd = ramp_tf(64, sigma_scale=20.0, cutoff=0.1).cuda()
    vh = vol.cuda().half()
    Vq, Vs = api.Volume(vh, quad=True), api.Volume(vh, quad=False)
    assert Vq.prepared(P)[2].volDtype == 4 and Vs.prepared(P)[2].volDtype == 1
    for skip in (1, 0):
        assert torch.equal(api.render(Vq, None, tfd, replace(P, skipEmpty=skip)), api.render(Vs, None, tfd, replace(P, skipEmpty=skip)))
    lo, hi = (8, 0, 16), (40, 43, 36)
    sub = vh[:, lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1].contiguous()
    Sq = api.Volume(sub, shard=(lo, hi), global_dims=(52, 44, 37), quad=True)
    Ss = api.Volume(sub, shard=(lo, hi), global_dims=(52, 44, 37), quad=False)
    a, b = Sq.forward(P, tfd), Ss.forward(P, tfd)
    assert torch.equal(a, b) and float(a[..., 3].min()) < 1.0


@pytest.mark.parametrize("box_edge,tile", [(8, 8), (16, 8), (8, 16), (16, 16)])
@pytest.mark.parametrize("ortho,use_tf", [(False, True), (True, True), (False, False)])
def test_staged_brick_tma_variant_renders_the_same_image(cuda, box_edge, tile, ortho, use_tf):
    """mrt_render_forward_tma (boxes staged through shared memory by 3-D TMA loads) against the direct
    gathers of mrt_render_forward: the same sampler arithmetic on the same voxels, so the images agree
    to fp32 contraction noise; nearly every slot must really come from the staged boxes."""
    vol, _, P = small_scene(C=1, dims=(60, 52, 44), W=88, H=72, seed=31, ortho=ortho)
    P = replace(P, tfMode=int(use_tf), intensityAlpha=8.0, bgColor=(0.05, 0.0, 0.1), alphaMode=1)
    tfd = ramp_tf(64, sigma_scale=20.0, cutoff=0.1).cuda() if use_tf else None
    V = api.Volume(vol.cuda(), quad=False)
    packed, Cn, Pe = V.prepared(P)
    bits = V.skip_levels(P, tfd)
    ref = api.render_forward(Pe, packed, Cn, tfd, bits)
    stats = torch.zeros(4, dtype=torch.int64, device="cuda")
    img = api.render_forward_tma(Pe, packed, tfd, bits, box_edge=box_edge, tile=tile, stats=stats)
    d = (img - ref).abs().amax(-1)
    assert float(d.max()) <= 2e-6, float(d.max())
    staged, direct, boxes, overflow = (int(x) for x in stats.tolist())
    assert staged > 0 and boxes > 0 and overflow == 0
    assert direct <= 0.02 * staged, (staged, direct)


def test_bad_arguments_raise(cuda):
    vol, _, P = small_scene(C=1, dims=(16, 16, 16), W=16, H=16)
    V = api.Volume(vol.cuda())
    with pytest.raises(RuntimeError):
        api.render(vol, None, None, P)                      # CPU tensor: no fallback
    with pytest.raises(ValueError):
        api.render(V, None, None, replace(P, ww=0.0))
    with pytest.raises(ValueError):
        api.render(V, None, None, replace(P, dims=(8, 8, 8)))
    with pytest.raises(api._lib.MrtError):
        api.render_forward(replace(P, fovY=4.0), V.packed, V.C)


def test_one_call_refold_step_equals_the_separate_calls(cuda):
    """mrt_render_views_refold (fold + occupancy + layout, classify, spans, march in ONE library call, taken
    by render_views whenever the folded volume is stale) gives the frames of the separate calls bit for
    bit, leaves the Volume's caches valid, and re-records the caller's timing events around the march."""
    import numpy as np
    from mri_raytracer_b200 import OrbitalCamera, orbit_views
    vol, _, P = small_scene(C=4, dims=(44, 37, 30), W=72, H=56, seed=17, theta_deg=20.0, phi_deg=70.0)
    P = replace(P, tfMode=1, volWeight=(1.0, 0.5, 2.0, 0.25))
    tf = ramp_tf(64, sigma_scale=12.0, cutoff=0.1).cuda()
    cam = OrbitalCamera(initial_radius=float(np.linalg.norm(np.asarray(P.eye))), initial_phi=1.2, initial_theta=0.3)
    cam.set_fov_degrees(70.0)
    cams = orbit_views(cam, 5)
    V = api.Volume(vol.cuda())
    ref = torch.stack([api.render(V, c, tf, P) for c in cams])          # separate calls (fold cached by the first)
    V.invalidate()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); e1.record()
    one = api.render_views(V, cams, tf, P, march_events=(e0, e1))       # stale -> the one-call step
    assert torch.equal(one, ref)
    torch.cuda.synchronize()
    assert e0.elapsed_time(e1) > 0.0
    again = api.render_views(V, cams, tf, P)                            # caches valid -> classify + march only
    assert torch.equal(again, ref)
    P2 = replace(P, volWeight=(0.3, 1.0, 1.0, 2.0))                     # weights changed -> stale again
    ref2 = torch.stack([api.render(api.Volume(vol.cuda()), c, tf, P2) for c in cams])
    assert torch.equal(api.render_views(V, cams, tf, P2), ref2)
    # the C form a non-Python host uses: stage 0 = the whole step in one call
    import ctypes as C
    from mri_raytracer_b200._lib import check, lib
    from mri_raytracer_b200 import tiles as _t
    X, Y, Z = 44, 37, 30
    quad = torch.empty((lib().mrt_packed_volume_bytes_quad(X, Y, Z) // 4,), device="cuda")
    mm = torch.empty((lib().mrt_brick_count(X, Y, Z), 1, 2), device="cuda")
    lv = torch.empty((lib().mrt_skip_levels_bytes(X, Y, Z),), dtype=torch.uint8, device="cuda")
    sp = torch.empty((5, _t.tiles_y(56), 2), dtype=torch.int32, device="cuda")
    out = torch.empty((5, 56, 72, 4), device="cuda")
    arr = api._camera_array(cams)
    s = P.with_projection_of(cams[0]).to_struct()
    planar = vol.cuda()
    check(lib().mrt_render_views_refold(C.byref(s), arr.ctypes.data, 5, planar.data_ptr(), 4, quad.data_ptr(), mm.data_ptr(),
                                        lv.data_ptr(), sp.data_ptr(), tf.data_ptr(), 64, out.data_ptr(), None, None, 0,
                                        torch.cuda.current_stream().cuda_stream), "render_views_refold")
    assert torch.equal(out, ref)
