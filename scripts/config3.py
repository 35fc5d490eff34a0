"""BASELINE config 3: textured Cornell box, cuda_texture_raytracer semantics, 1920x1080, 256 spp (16x16 grid),
depth 6 — plus the golden setting 2000x1500 / 25 spp / depth 16.  One JSON line per setting."""
import json, os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "path-tracing__ray-tracer_b200"))
import numpy as np, torch
from b200rt import renderer
from b200rt.cornell import CustomSceneBuilder
from b200rt.scene_api import RenderSettings

random.seed(0); b = CustomSceneBuilder(texture_dir=False); scene = b.build_scene()
for (W, H, SPP, D) in ((1920, 1080, 256, 6), (2000, 1500, 25, 16)):
    cam = b.create_camera(W / H)
    out = {"config": f"textured Whitted {W}x{H} {SPP} spp depth {D}"}
    for prec in ("f32", "f64"):
        r = renderer.B200TextureRaytracer(precision=prec)
        r.render(scene, cam, RenderSettings(W, H, SPP, D))            # warm-up
        ks = []
        for _ in range(3):
            img = r.render(scene, cam, RenderSettings(W, H, SPP, D)); ks.append(r.last_stats["kernel_s"])
        k = float(np.median(ks)); prim = r.last_stats["primary"]
        out[prec] = {"kernel_ms": k * 1e3, "Mprimary_per_s": prim / k / 1e6, "Mrays_per_s_at_9.55_per_primary": prim * 9.55 / k / 1e6,
                     "e2e_ms": r.last_stats["wall_s"] * 1e3}
    print(json.dumps(out))
