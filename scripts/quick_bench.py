"""A/B timing helper (not the headline bench): python scripts/quick_bench.py key=value ...
keys: spp (128) W H depth wave_paths scan (64) top (512) steps (2) precision (f32)"""
import ctypes as C, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "path-tracing__ray-tracer_b200"))
import torch
from b200rt import _lib, renderer
from b200rt.cornell import CustomSceneBuilder
from b200rt.scene_api import RenderSettings

kw = dict(spp=128, W=1920, H=1080, depth=8, wave_paths=1 << 24, scan=64, top=512, steps=2, precision="f32", fused=1, hints=1, boxes=1, pwalk=0, surf=1)
for a in sys.argv[1:]:
    k, v = a.split("="); kw[k] = type(kw[k])(v)
random.seed(0); b = CustomSceneBuilder(texture_dir=False); scene = b.build_scene(); cam = b.create_camera(kw["W"] / kw["H"])
lib = _lib.load()
r = renderer.B200PathTracer(precision=kw["precision"], wave_paths=kw["wave_paths"], scan_max_prims=kw["scan"], top_nodes=kw["top"], fused=bool(kw["fused"]), occluder_hints=bool(kw["hints"]), scan_boxes=bool(kw["boxes"]), primary_walk=bool(kw["pwalk"]), surface_records=bool(kw["surf"]))
st = r.prepare(scene, cam, RenderSettings(kw["W"], kw["H"], kw["spp"], kw["depth"]))
r.accumulate(st); torch.cuda.synchronize()
st["counters"].zero_(); lib.b2rt_profile_enable(1)
from bench import ClockSampler
sampler = ClockSampler(0); sampler.start(); time.sleep(0.3)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(kw["steps"]):
    r.accumulate(st)
e1.record(); torch.cuda.synchronize()
clocks = sampler.stop()
tfl = C.c_double(0); lib.b2rt_fp32_peak(200000, C.byref(tfl), None)
ms = (C.c_double * 8)(); nl = (C.c_int64 * 8)(); lib.b2rt_profile_read(ms, nl); lib.b2rt_profile_enable(0)
cnt = st["counters"].cpu().numpy(); dt = e0.elapsed_time(e1) * 1e-3
print({k: kw[k] for k in ("spp", "wave_paths", "scan", "top", "precision", "fused", "hints", "boxes", "pwalk", "surf")}, "wave", st["wave"],
      "Mpaths/s %.1f  Mrays/s %.1f  rays/path %.3f" % (cnt[0] / dt / 1e6, (cnt[1] + cnt[2]) / dt / 1e6, (cnt[1] + cnt[2]) / cnt[0]),
      "kernel-seen MHz %.0f" % (cnt[11] / max(1, cnt[12]) * 1e3), "clocks", clocks, "fp32 peak %.1f TF" % tfl.value,
      "ms/step: " + " ".join(f"{n}={ms[i] / kw['steps']:.1f}" for i, n in enumerate(["raygen", "extend", "shade", "shadow", "accum"])))
