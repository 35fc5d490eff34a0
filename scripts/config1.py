"""BASELINE config 1: Cornell box from custom_scene_builder, Whitted ray tracer with cpu_renderer semantics,
320x240, 1 spp, depth 4 (the reference's own CPU-runnable case: 54.4 s in the pure-Python renderer).
Prints one JSON line: GPU kernel time (float64 parity instantiation and float32), the oracle port's time on the
host cores, and the max abs difference between the two (the 1e-4 bar of the north star)."""
import json, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "path-tracing__ray-tracer_b200"))
import numpy as np, torch
from b200rt import renderer
from b200rt.cornell import CustomSceneBuilder
from oracle import cpu_oracle as O

W, H, D = 320, 240, 4
random.seed(0); b = CustomSceneBuilder(texture_dir=False); scene = b.build_scene(); cam = b.create_camera(W / H)
out = {"config": "C1 Whitted (cpu_renderer semantics) 320x240 1 spp depth 4, pixel centres"}
t0 = time.perf_counter(); ref = O.cpu_whitted(O.cpu_export(scene, cam), W, H, D)["rgb"]; out["oracle_port_s"] = time.perf_counter() - t0
out["oracle_threads"] = O.num_threads()
for prec in ("f64", "f32"):
    r = renderer.B200WhittedRenderer(precision=prec, jitter_seed=None)
    r.trace(scene, cam, W, H, D)
    ts = []
    for _ in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter(); rgb = r.trace(scene, cam, W, H, D); ts.append(time.perf_counter() - t0)
    out[prec] = {"trace_call_ms": float(np.median(ts)) * 1e3, "max_abs_vs_oracle": float(np.abs(rgb - ref).max()),
                 "p999_abs_vs_oracle": float(np.quantile(np.abs(rgb - ref), 0.999))}
print(json.dumps(out))
