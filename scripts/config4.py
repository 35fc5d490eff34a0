"""BASELINE config 4: synthetic 1M-triangle mesh scene — LBVH build + traversal, 1920x1080, 64 spp, depth 4.
Prints one JSON line: LBVH build ms / Mtris/s, path-tracing Mpaths/s / Mrays/s, per-kernel ms, and a
BVH-vs-brute-force equality check on random rays."""
import ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "path-tracing__ray-tracer_b200"))
import numpy as np, torch
from b200rt import _lib, packer, renderer, scenes
from b200rt.device import DeviceScene, current_stream_ptr
from b200rt.scene_api import RenderSettings

kw = dict(W=1920, H=1080, spp=64, depth=4, steps=2, check=4096, top=512, sort=1, fwalk=0, wprim=0, rout=1, rot=1, wide=0, quant=0, morton=0)
for a in sys.argv[1:]:
    k, v = a.split("="); kw[k] = type(kw[k])(v)
lib = _lib.load()
dev = torch.device("cuda", 0)
t0 = time.perf_counter(); scene, b = scenes.heightfield_scene(); cam = b.create_camera(kw["W"] / kw["H"])
if kw["morton"]:                                              # faces of the mesh along a Morton curve (TriangleMesh.spatially_sorted)
    scene.objects = [o.spatially_sorted() if isinstance(o, packer.TriangleMesh) else o for o in scene.objects]
t_scene = time.perf_counter() - t0
t0 = time.perf_counter(); pk = packer.pack_scene(scene, "numba"); t_pack = time.perf_counter() - t0
ds = DeviceScene(pk, _lib.P_F32, dev, kw["top"], ray_origin_extent=50.0, rects_outside=bool(kw["rout"]), lbvh_rotations=bool(kw["rot"]),
                 wide_nodes=bool(kw["wide"]), quant_nodes=bool(kw["quant"]))
# ---- LBVH build timing (CUDA events around the whole build call, 5 repetitions)
n = pk.n_prims
need = C.c_size_t(0); lib.b2rt_lbvh_temp_bytes(n, C.byref(need))
temp = torch.empty(need.value, dtype=torch.uint8, device=dev)
meta = (C.c_int32 * 3)()
times = []
for _ in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.b2rt_lbvh_build(pk.n_rect, pk.n_sphere, pk.n_tri, ds.rect.data_ptr(), ds.sphere.data_ptr(), ds.tri.data_ptr(),
                                   C.c_float(ds.box_pad), ds.nodes.data_ptr(), ds.top.data_ptr(), kw["top"], meta,
                                   temp.data_ptr(), temp.numel(), current_stream_ptr(dev), (1 if ds.rects_outside else 0) | (0 if kw["rot"] else 2)), "lbvh")
    if ds.wide is not None:                                   # the 4-wide nodes are part of the build
        _lib.check(lib.b2rt_lbvh_widen(ds.nodes.data_ptr(), ds.top.data_ptr(), meta[0], meta[2], ds.wide.data_ptr(), ds.wide.numel(),
                                       current_stream_ptr(dev)), "widen")
    e1.record(); torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
build_ms = float(np.median(times[1:]))
# ---- BVH walk == brute-force scan on random rays
rng = np.random.default_rng(3)
m = kw["check"]
o = np.concatenate([np.tile([0, 0, 50.0], (m // 2, 1)), rng.uniform(-13, 13, (m // 2, 3)) * [1, 0.2, 1] + [0, -5, 0]])
d = rng.normal(size=(m, 3)); d[: m // 2, 2] = -np.abs(d[: m // 2, 2]) * 4; d /= np.linalg.norm(d, axis=1, keepdims=True)
a_ids, a_rec = renderer.trace_rays(scene, o, d, "numba", "f32", use_bvh=1, packed=pk)
b_ids, b_rec = renderer.trace_rays(scene, o, d, "numba", "f32", use_bvh=0, packed=pk)
same = bool(np.array_equal(a_ids, b_ids) and np.array_equal(a_rec[:, 0], b_rec[:, 0]))
# ---- path tracing throughput
r = renderer.B200PathTracer(precision="f32", top_nodes=kw["top"], sort_rays=bool(kw["sort"]), fused_walk=bool(kw["fwalk"]), walk_primary=bool(kw["wprim"]), rects_outside=bool(kw["rout"]), lbvh_rotations=bool(kw["rot"]), wide_walk=bool(kw["wide"]), quant_walk=bool(kw["quant"]))
st = r.prepare(scene, cam, RenderSettings(kw["W"], kw["H"], kw["spp"], kw["depth"]))
r.accumulate(st); torch.cuda.synchronize()
st["counters"].zero_(); lib.b2rt_profile_enable(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(kw["steps"]):
    r.accumulate(st)
e1.record(); torch.cuda.synchronize()
ms = (C.c_double * 8)(); nl = (C.c_int64 * 8)(); lib.b2rt_profile_read(ms, nl); lib.b2rt_profile_enable(0)
cnt = st["counters"].cpu().numpy(); dt = e0.elapsed_time(e1) * 1e-3
img = r.render(scene, cam, RenderSettings(kw["W"], kw["H"], min(kw["spp"], 16), kw["depth"]))
img.save(os.path.join(ROOT, "gpurun_out", "config4.png")) if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else None
print(json.dumps({
    "config": "C4 heightfield 1M triangles", "n_prims": int(n), "scene_build_s": t_scene, "pack_s": t_pack,
    "lbvh_build_ms": build_ms, "lbvh_mtris_per_s": n / build_ms / 1e3, "lbvh_builds_ms": times,
    "bvh_nodes": ds.n_internal, "bvh_top_nodes": ds.n_top, "bvh_equals_bruteforce": same, "rays_checked": m,
    "hit_fraction": float((a_ids >= 0).mean()),
    "mpaths_per_s": cnt[0] / dt / 1e6, "mrays_per_s": (cnt[1] + cnt[2]) / dt / 1e6, "rays_per_path": float((cnt[1] + cnt[2]) / cnt[0]),
    "ms_per_step": dt / kw["steps"] * 1e3, "wave": st["wave"],
    "kernel_ms_per_step": {k: ms[i] / kw["steps"] for i, k in enumerate(["raygen", "extend", "bounce", "shadow", "accumulate", "sort"])},
    "sort_rays": kw["sort"], "fused_walk": kw["fwalk"], "walk_primary": kw["wprim"], "rects_outside": kw["rout"], "lbvh_rotations": kw["rot"], "wide_walk": kw["wide"], "quant_walk": kw["quant"], "morton_face_order": kw["morton"],
}))
