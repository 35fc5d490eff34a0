"""BASELINE config 1 in float32 against the oracle: the fraction of pixels that differ (profiles/r2_c1_f32_fraction.json)."""
import json, os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "path-tracing__ray-tracer_b200"))
import numpy as np
from b200rt import renderer
from b200rt.cornell import CustomSceneBuilder
from oracle import cpu_oracle as O
random.seed(0); b = CustomSceneBuilder(texture_dir=False); scene = b.build_scene(); cam = b.create_camera(320 / 240)
ref = O.cpu_whitted(O.cpu_export(scene, cam), 320, 240, 4)["rgb"]
out = {}
for prec in ("f64", "f32"):
    rgb = renderer.B200WhittedRenderer(precision=prec, jitter_seed=None).trace(scene, cam, 320, 240, 4)
    d = np.abs(rgb - ref).max(axis=2)
    out[prec] = {"max_abs": float(d.max()), "frac_gt_1e-4": float((d > 1e-4).mean()), "frac_gt_1e-2": float((d > 1e-2).mean()),
                 "p99": float(np.quantile(d, 0.99)), "p99.9": float(np.quantile(d, 0.999))}
print(json.dumps(out))
