"""mri_raytracer_b200 — B200-native volume ray-marcher (hand-written sm_100a CUDA behind a C ABI).

Scope: the reference's volume ray-march hot path (klukaszek/MRI-RayTracer:
inr/viewer/brats_rt.slang, scripts/volumeRendering/volume_render.slang, the orbital cameras
and the loader-side layout), plus the differentiable-rendering backward of
docs/DifferentiableRendering.md.  See DESIGN.md.
"""
from .camera import Camera, OrbitalCamera, OrbitalCameraYUp, orbit_views
from .params import RenderParams, SlabParams, default_label_lut
from . import tiles

__all__ = ["Camera", "OrbitalCamera", "OrbitalCameraYUp", "orbit_views", "RenderParams", "SlabParams",
           "default_label_lut", "tiles", "render", "render_views", "render_aux", "render_slab", "render_host", "HostPipeline", "Volume", "TrainStep"]


def __getattr__(name):
    # torch-dependent API is imported lazily so the pure-host pieces work without torch/CUDA
    if name in ("render", "inr_predict", "ray_gradients", "render_views", "render_aux", "render_slab", "render_host", "Volume", "TrainStep", "HostPipeline", "pack_volume", "unpack_volume",
                "render_forward", "render_backward", "build_occupancy", "classify_bricks", "tile_index_map",
                "build_label_occupancy"):
        from . import api
        return getattr(api, name)
    if name in ("make_brats_like", "ramp_tf", "world_box"):
        from . import synth
        return getattr(synth, name)
    raise AttributeError(name)
