"""Where the end-to-end render() time goes beyond the kernels (host packing, uploads, LBVH build, read-back)."""
import os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "path-tracing__ray-tracer_b200"))
import torch
from b200rt import renderer
from b200rt.cornell import CustomSceneBuilder
from b200rt.scene_api import RenderSettings
import cProfile, pstats

W, H, SPP, D = 1920, 1080, int(sys.argv[1]) if len(sys.argv) > 1 else 128, 8
random.seed(0); b = CustomSceneBuilder(texture_dir=False); scene = b.build_scene(); cam = b.create_camera(W / H)
r = renderer.B200PathTracer(precision="f32")
st = RenderSettings(W, H, SPP, D)
for _ in range(2):
    r._tex_cache.enabled = False
    r.render(scene, cam, st)
ts = []
for _ in range(5):
    r._tex_cache.enabled = False
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r.render(scene, cam, st)
    torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
print("render wall ms", [round(t * 1e3, 2) for t in ts], "kernel ms", round(r.last_stats["kernel_s"] * 1e3, 2))
pr = cProfile.Profile()
r._tex_cache.enabled = False
pr.enable(); r.render(scene, cam, st); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
