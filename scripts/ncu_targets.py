"""Minimal launch sequences for the per-kernel ncu captures (profiles/r2_*): one instance of every kernel class.

    python scripts/ncu_targets.py c2      # Cornell 1080p, one 32-spp wave: mask, tile list, bounce, shadow, accumulate, resolve, LBVH (34 prims)
    python scripts/ncu_targets.py c4      # 1 M triangles, one 8-spp wave: LBVH build (bounds .. finalize + CUB sort), fused first bounce,
                                          # ray sort, persistent walk, shade stage, shadow
    python scripts/ncu_targets.py c3      # textured Whitted 1080p x 16 spp (f32) ; c1: CPU-semantics Whitted 320x240 (f64)
"""
import os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "path-tracing__ray-tracer_b200"))
import torch
from b200rt import renderer, scenes
from b200rt.cornell import CustomSceneBuilder
from b200rt.scene_api import RenderSettings

which = sys.argv[1]
random.seed(0)
b = CustomSceneBuilder(texture_dir=False)
if which == "c2":
    scene = b.build_scene(); cam = b.create_camera(1920 / 1080)
    r = renderer.B200PathTracer(precision="f32")
    r.render(scene, cam, RenderSettings(1920, 1080, 32, 8))
elif which == "c4":
    scene, b4 = scenes.heightfield_scene(); cam = b4.create_camera(1920 / 1080)
    r = renderer.B200PathTracer(precision="f32")
    r.render(scene, cam, RenderSettings(1920, 1080, 8, 4))
elif which == "c4big":                       # the bench's wave size (32 spp per wave): one walk-kernel launch at full occupancy of the queues
    scene, b4 = scenes.heightfield_scene(); cam = b4.create_camera(1920 / 1080)
    r = renderer.B200PathTracer(precision="f32")
    r.render(scene, cam, RenderSettings(1920, 1080, 32, 4))
elif which == "c3":
    scene = b.build_scene(); cam = b.create_camera(1920 / 1080)
    renderer.B200TextureRaytracer(precision="f32").render(scene, cam, RenderSettings(1920, 1080, 16, 6))
elif which == "c1":
    scene = b.build_scene(); cam = b.create_camera(320 / 240)
    renderer.B200WhittedRenderer(precision="f64", jitter_seed=None).trace(scene, cam, 320, 240, 4)
torch.cuda.synchronize()
print("done", which)
