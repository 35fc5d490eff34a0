"""C4 (1 M-triangle height field): persistent walk kernel on quantised 32 B nodes (quant_walk=True) against the default
64 B nodes — (1) the float sums and counters of a 960x540 x 4 spp x depth 4 render must be IDENTICAL, (2) per-kernel times
of the 1080p x 64 spp x depth 4 step for both.  One JSON line per stage (flushed), so a cut-off run still reports."""
import ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "path-tracing__ray-tracer_b200"))
import numpy as np, torch
from b200rt import _lib, renderer, scenes
from b200rt.scene_api import RenderSettings

t00 = time.perf_counter()
lib = _lib.load()
scene, b = scenes.heightfield_scene()
cam = b.create_camera(1920 / 1080)
out = {}
acc = {}
for name, kw in (("binary", {}), ("quant", {"quant_walk": True})):
    r = renderer.B200PathTracer(precision="f32", seed=3, distributed=False, **kw)
    a, c = r.render_accum(scene, cam, RenderSettings(960, 540, 4, 4))
    acc[name] = (a, c)
    del r
same = bool(np.array_equal(acc["binary"][0], acc["quant"][0]) and np.array_equal(acc["binary"][1][:4], acc["quant"][1][:4]))
print(json.dumps({"stage": "equality", "sums_and_counters_identical": same,
                  "max_abs_diff": float(np.abs(acc["binary"][0] - acc["quant"][0]).max()),
                  "counters": [int(x) for x in acc["quant"][1][:4]], "elapsed_s": time.perf_counter() - t00}), flush=True)
ws = None
for name, kw in (("binary", {}), ("quant", {"quant_walk": True}), ("quant_counted", {"quant_walk": True, "count_tests": True})):
    r = renderer.B200PathTracer(precision="f32", distributed=False, **kw)
    if ws is not None:
        r._ws = ws
    counted = name.endswith("counted")
    st = r.prepare(scene, cam, RenderSettings(1920, 1080, 8 if counted else 64, 4))
    ws = r._ws
    r.accumulate(st); torch.cuda.synchronize()
    if counted:
        cc = st["counters"].cpu().numpy().astype(np.float64)
        rays = max(1.0, cc[1] - cc[0])
        print(json.dumps({"stage": name, "box_steps_per_ray": cc[8] / rays, "leaf_steps_per_ray": cc[9] / rays}), flush=True)
        continue
    st["counters"].zero_(); lib.b2rt_profile_enable(1)
    ms = (C.c_double * 8)(); nl = (C.c_int64 * 8)(); lib.b2rt_profile_read(ms, nl)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 2
    e0.record()
    for _ in range(steps):
        r.accumulate(st)
    e1.record(); torch.cuda.synchronize()
    lib.b2rt_profile_read(ms, nl); lib.b2rt_profile_enable(0)
    cnt = st["counters"].cpu().numpy(); dt = e0.elapsed_time(e1) * 1e-3
    print(json.dumps({"stage": name, "ms_per_step": dt / steps * 1e3, "mpaths_per_s": float(cnt[0]) / dt / 1e6,
                      "kernel_ms_per_step": {k: ms[i] / steps for i, k in enumerate(["raygen", "walk", "bounce0_and_shade", "shadow", "accumulate", "ray_sort"])},
                      "elapsed_s": time.perf_counter() - t00}), flush=True)
    del r, st
