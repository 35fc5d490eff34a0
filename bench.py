#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 path-tracing core.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-extras]

Workload at every N: BASELINE.json configs[1] — Cornell box path trace, 1920x1080, 1024 spp,
max depth 8 (glass-sphere caustics), synthetic textures of the reference's dimensions.
A "step" renders that whole frame once.  With N > 1 (torchrun, one rank per GPU) the 1024 samples
per pixel are split across ranks (strong scaling: total work fixed) and summed with one NCCL reduce.

One JSON line on rank 0:
  value          Mpaths/s, inputs resident in HBM, CUDA-event time of K steps, max over ranks
  e2e            the same through the public renderer API: render(scene, camera, settings) -> PIL image
                 (scene + texture H2D and image D2H inside the timed region); e2e_cold_ms = the FIRST render() of
                 the process (library load, host packing, pinned buffers, 13 GB workspace, first launches)
  mrays_per_s    rays actually TRACED (closest-hit + queued shadow rays); the shadow rays that the one-primitive
                 occluder hint answered are reported separately (mrays_answered_by_hint_per_s)
  roofline       dominant kernel (fused closest-hit + shade bounce kernel): algorithmic queue bytes / measured launch time
  hbm            whole step: algorithmic bytes from the device counters and DRAM bytes measured by ncu
  fp32           executed = counted intersection tests x canonical costs + shading; useful_reference_flops
                 = what the REFERENCE algorithm would spend on the same rays (1 070 flop/ray); both against the FMA peak
                 measured in this run
  cpu_baseline   the oracle port (C, float64, reference algorithm) on the host cores, bounded sample
  extra_configs  (N = 1) BASELINE configs 1, 3, 4 and the float64 parity instantiation of config 2, timed in the same
                 run, and the UNMODIFIED reference GPU renderer from baseline/_ref when that checkout travelled;
                 config 5 (4K x 4096 spp) at every N
  multi_gpu_parity  (N > 1) the NCCL-reduced 8-spp sums against the same 8 spp rendered by rank 0 alone
--impl reference times the CPU implementation instead (all host threads, bounded sample per step).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import random
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "path-tracing__ray-tracer_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

W, H, SPP, DEPTH = 1920, 1080, 1024, 8
WORKLOAD = "cornell_path_1920x1080_1024spp_depth8"
FLOPS_PER_RAY = 1070.0          # reference-algorithm intersection cost per ray (measured)
QUEUE_RECORD_BYTES = 48.0       # one ray-queue or shadow-queue record (3 float4 streams)
# canonical costs of SURVEY 8d (flops): ray-AABB slab test, ray-triangle, ray-sphere, ray-rectangle, shading per segment;
# a box record is a slab test in the box's own frame (+18 flops to take origin and direction there); a planar scan
# record or occluder-hint test is a rectangle test; camera-ray set-up 21
COST = {"slab": 24.0, "triangle": 45.0, "sphere": 28.0, "rect": 33.0, "box_record": 42.0, "shade": 150.0, "camera": 21.0}
# dram__bytes_read.sum + dram__bytes_write.sum from one `ncu --set full` capture of one 32-spp wave at 1080p
# (profiles/, file named in NCU_SOURCE): bounce launches, shadow launches, accumulate
NCU_SOURCE = "profiles/r2_ncu_full_c2.csv"
NCU_WAVE_PATHS = 1920 * 1080 * 32
NCU_BOUNCE_BYTES_PER_WAVE = 12.981e9      # 8 bounce launches: 6.958 GB read + 6.023 GB written
NCU_SHADOW_BYTES_PER_WAVE = 1.499e9       # 8 shadow launches
NCU_ACCUM_BYTES_PER_WAVE = 0.593e9        # accumulate (tiles that see nothing are never written or read)
NCU_TRAFFIC_BYTES_PER_LAUNCH = NCU_BOUNCE_BYTES_PER_WAVE / 8


# stdout carries exactly ONE line, the JSON result: everything else that libraries print to file descriptor 1 (NCCL's
# "NCCL version ..." banner at NCCL_DEBUG=WARN/VERSION, numba, ...) is sent to stderr by pointing fd 1 at fd 2 for the
# whole run and writing the result to the saved descriptor.
_RESULT_FD = None


def claim_stdout() -> None:
    """Called first thing in main(): from here on fd 1 is stderr, the result line goes to the saved descriptor."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj) -> None:
    line = (json.dumps(obj) + "\n").encode()
    os.write(_RESULT_FD if _RESULT_FD is not None else 1, line)


def log(*a) -> None:
    print("[bench]", *a, file=sys.stderr, flush=True)


def build_scene(aspect: float = W / H):
    from b200rt.cornell import CustomSceneBuilder
    random.seed(0)
    b = CustomSceneBuilder(texture_dir=False)
    return b.build_scene(), b.create_camera(aspect), b


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            try:
                pw.append(float(r[3]))
            except Exception:
                pass
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        pw.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "sm_mhz_min": sm[0] if sm else None, "power_w_median": pw[len(pw) // 2] if pw else None,
                "power_w_max": pw[-1] if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


def host_threads() -> int:
    """All host cores the process may use — NOT OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1 to every rank)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_baseline(scene, camera, budget_s: float = 15.0):
    """Oracle port (reference algorithm, float64, OpenMP) on a bounded sample of the same workload."""
    from oracle import cpu_oracle as O
    O.set_num_threads(host_threads())
    cores = O.num_threads()
    pk = O.nb_pack(scene, camera)
    w, h = W // 4, H // 4
    t0 = time.perf_counter()
    O.nb_path_trace(pk, w, h, 2, DEPTH, 0, want_stats=False)
    rate = w * h * 2 / (time.perf_counter() - t0)
    spp = int(max(4, min(4096, budget_s * rate / (w * h))))
    t0 = time.perf_counter()
    res = O.nb_path_trace(pk, w, h, spp, DEPTH, 0, want_stats=False)
    dt = time.perf_counter() - t0
    paths = w * h * spp
    rays = res["counters"]["closest_rays"] + res["counters"]["shadow_rays"]
    return {"value": paths / dt / 1e6, "unit": "Mpaths/s", "cores": cores, "kind": "port",
            "sample": f"{w}x{h} x {spp} spp of the depth-{DEPTH} Cornell path trace ({paths} paths, {dt:.1f} s)",
            "mrays_per_s": rays / dt / 1e6, "precision": "f64"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_oracle as O
    O.set_num_threads(host_threads())          # torchrun sets OMP_NUM_THREADS=1: the CPU arm uses every host core
    scene, camera, _ = build_scene()
    pk = O.nb_pack(scene, camera)
    cores = O.num_threads()
    w, h = W // 4, H // 4
    t0 = time.perf_counter()
    O.nb_path_trace(pk, w, h, 2, DEPTH, 0, want_stats=False)
    rate = w * h * 2 / (time.perf_counter() - t0)
    total_budget = 150.0
    per_step = total_budget / max(1, args.steps + args.warmup)
    spp = int(max(1, min(64, per_step * rate / (w * h))))
    for _ in range(args.warmup):
        O.nb_path_trace(pk, w, h, spp, DEPTH, 0, want_stats=False)
    t0 = time.perf_counter()
    for s in range(args.steps):
        O.nb_path_trace(pk, w, h, spp, DEPTH, s, want_stats=False)
    dt = time.perf_counter() - t0
    paths = w * h * spp * args.steps
    v = paths / dt / 1e6
    sample = f"each step = {w}x{h} x {spp} spp of the workload ({w * h * spp} paths)"
    emit(({
        "impl": "reference", "metric": "Mpaths/s", "value": v, "unit": "Mpaths/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "width": W, "height": H, "spp": SPP, "max_depth": DEPTH, "sample": sample},
        "cpu_baseline": {"value": v, "unit": "Mpaths/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------ extra configurations
def _events():
    import torch
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def extra_c2_f64(scene, camera, dev):
    """BASELINE config 2 in the float64 PARITY instantiation (-fmad=false, reference xorshift, exact replay of the
    reference sample for sample) at a reduced spp: what the bit-faithful mode costs."""
    import torch
    from b200rt import renderer
    from b200rt.scene_api import RenderSettings
    spp = 8
    r = renderer.B200PathTracer(precision="f64", rng="reference", device=dev, distributed=False)
    st = r.prepare(scene, camera, RenderSettings(W, H, spp, DEPTH))
    r.accumulate(st); torch.cuda.synchronize(dev)
    st["counters"].zero_()
    e0, e1 = _events()
    e0.record(); r.accumulate(st); e1.record(); torch.cuda.synchronize(dev)
    s = e0.elapsed_time(e1) * 1e-3
    c = st["counters"].cpu().numpy()
    return {"workload": f"cornell_path_{W}x{H}_{spp}spp_depth{DEPTH}", "dtype": "f64", "rng": "reference xorshift (exact replay)",
            "mpaths_per_s": float(c[0]) / s / 1e6, "mrays_per_s": float(c[1] + c[2] - c[5]) / s / 1e6, "ms": s * 1e3,
            "note": "parity instantiation, reduced spp (the rate does not depend on spp); headline config is f32"}


def extra_c3(scene, builder, dev):
    """BASELINE config 3: textured Cornell box, cuda_texture_raytracer semantics (the reference's DEFAULT renderer,
    main.py:27), 1920x1080, 256 spp (16x16 grid), depth 6."""
    import numpy as np
    from b200rt import renderer
    from b200rt.scene_api import RenderSettings
    cam = builder.create_camera(W / H)
    out = {"workload": "cornell_whitted_texture_1920x1080_256spp_depth6"}
    for prec, reps in (("f32", 3), ("f64", 1)):
        r = renderer.B200TextureRaytracer(precision=prec, device=dev)
        r.render(scene, cam, RenderSettings(W, H, 16, 6))                  # warm-up at 16 spp
        ks, ws = [], []
        for _ in range(reps):
            r.render(scene, cam, RenderSettings(W, H, 256, 6))
            ks.append(r.last_stats["kernel_s"]); ws.append(r.last_stats["wall_s"])
        k, prim = float(np.median(ks)), r.last_stats["primary"]
        out[prec] = {"kernel_ms": k * 1e3, "mprimary_per_s": prim / k / 1e6, "e2e_ms": float(np.median(ws)) * 1e3,
                     "mrays_per_s_at_9.55_scene_hits_per_primary": prim * 9.55 / k / 1e6}
    return out


def extra_c1(dev):
    """BASELINE config 1: cpu_raytracer semantics (CPURenderer._trace), 320x240, 1 spp, depth 4, float64."""
    import torch
    from b200rt import renderer
    scene, cam, _ = build_scene(320 / 240)
    r = renderer.B200WhittedRenderer(precision="f64", jitter_seed=None, device=dev)
    r.trace(scene, cam, 320, 240, 4)
    t0 = time.perf_counter()
    r.trace(scene, cam, 320, 240, 4)
    torch.cuda.synchronize(dev)
    return {"workload": "cornell_whitted_cpu_semantics_320x240_1spp_depth4", "dtype": "f64",
            "call_ms": (time.perf_counter() - t0) * 1e3, "note": "host packing + kernel + read-back per call"}


def extra_c4(dev, steps: int = 2):
    """BASELINE config 4: synthetic 1 M-triangle height field inside the Cornell walls, LBVH build + traversal,
    1920x1080, 64 spp, depth 4.  Box / leaf steps come from a separate counted pass (B2RT_PATH_COUNT_TESTS)."""
    import numpy as np
    import torch
    from b200rt import _lib, packer, renderer, scenes
    from b200rt.device import current_stream_ptr
    from b200rt.scene_api import RenderSettings
    lib = _lib.load()
    t0 = time.perf_counter()
    scene, b = scenes.heightfield_scene()
    cam = b.create_camera(W / H)
    host_s = time.perf_counter() - t0
    spp, depth = 64, 4
    r = renderer.B200PathTracer(precision="f32", device=dev, distributed=False)
    t0 = time.perf_counter()
    st = r.prepare(scene, cam, RenderSettings(W, H, spp, depth))
    torch.cuda.synchronize(dev)
    prepare_s = time.perf_counter() - t0
    ds = st["ds"]
    n = ds.packed.n_prims
    # LBVH build alone: CUDA events around the build call, median of 5
    need = C.c_size_t(0)
    lib.b2rt_lbvh_temp_bytes(n, C.byref(need))
    temp = torch.empty(need.value, dtype=torch.uint8, device=dev)
    meta = (C.c_int32 * 3)()
    times = []
    for _ in range(6):
        e0, e1 = _events()
        e0.record()
        _lib.check(lib.b2rt_lbvh_build(ds.packed.n_rect, ds.packed.n_sphere, ds.packed.n_tri, ds.rect.data_ptr(),
                                       ds.sphere.data_ptr(), ds.tri.data_ptr(), C.c_float(ds.box_pad),
                                       ds.nodes.data_ptr(), ds.top.data_ptr(), ds.n_top, meta, temp.data_ptr(),
                                       temp.numel(), current_stream_ptr(dev), 1 if ds.rects_outside else 0), "lbvh")
        if ds.wide is not None:                                         # the 4-wide nodes are part of the build
            _lib.check(lib.b2rt_lbvh_widen(ds.nodes.data_ptr(), ds.top.data_ptr(), meta[0], meta[2], ds.wide.data_ptr(),
                                           ds.wide.numel(), current_stream_ptr(dev)), "widen")
        e1.record(); torch.cuda.synchronize(dev)
        times.append(e0.elapsed_time(e1))
    build_ms = float(np.median(times[1:]))
    r.accumulate(st); torch.cuda.synchronize(dev)                      # warm-up
    st["counters"].zero_()
    lib.b2rt_profile_enable(1)
    ms = (C.c_double * 8)(); nl = (C.c_int64 * 8)()
    lib.b2rt_profile_read(ms, nl)                                       # drop the warm-up records
    e0, e1 = _events()
    e0.record()
    for _ in range(steps):
        r.accumulate(st)
    e1.record(); torch.cuda.synchronize(dev)
    _lib.check(lib.b2rt_profile_read(ms, nl), "b2rt_profile_read")
    lib.b2rt_profile_enable(0)
    s = e0.elapsed_time(e1) * 1e-3
    c = st["counters"].cpu().numpy().astype(np.float64)
    # counted pass (untimed): box / leaf steps of the persistent walk kernel
    rc = renderer.B200PathTracer(precision="f32", device=dev, distributed=False, count_tests=True)
    rc._ws = r._ws
    stc = rc.prepare(scene, cam, RenderSettings(W, H, 8, depth))
    rc.accumulate(stc); torch.cuda.synchronize(dev)
    cc = stc["counters"].cpu().numpy().astype(np.float64)
    walk_rays = max(1.0, cc[1] - cc[0])                                 # closest-hit rays of bounce >= 1
    nodes_per_ray, leaves_per_ray = cc[8] / walk_rays, cc[9] / walk_rays
    walk_ms = ms[1] / steps
    walk_rays_step = (c[1] - c[0]) / steps
    wide = ds.wide is not None                                          # 4-wide nodes: 128 B and four slab tests per box step
    node_bytes, node_slabs = (128.0, 4) if wide else (64.0, 2)
    out = {
        "workload": "heightfield_1M_triangles_1920x1080_64spp_depth4", "dtype": "f32", "n_prims": int(n),
        "lbvh_build_ms": build_ms, "lbvh_mtris_per_s": n / build_ms / 1e3,
        "mpaths_per_s": c[0] / s / 1e6, "mrays_per_s": (c[1] + c[2] - c[5]) / s / 1e6, "ms_per_step": s / steps * 1e3,
        "kernel_ms_per_step": {k: ms[i] / steps for i, k in enumerate(["raygen", "walk", "bounce0_and_shade", "shadow", "accumulate", "ray_sort"])},
        "walk_kernel": {"ms_per_step": walk_ms, "grays_per_s": walk_rays_step / (walk_ms * 1e-3) / 1e9 if walk_ms else None,
                        "box_steps_per_ray": nodes_per_ray, "leaf_steps_per_ray": leaves_per_ray,
                        "node_width": node_slabs,
                        "algorithmic_bytes_per_ray": nodes_per_ray * node_bytes + leaves_per_ray * 48.0 + 48.0 + 16.0,
                        "executed_tflops": walk_rays_step * (nodes_per_ray * node_slabs * COST["slab"] + leaves_per_ray * COST["triangle"]) / (walk_ms * 1e-3) / 1e12 if walk_ms else None,
                        "note": "box step = one %d B node (%d child boxes), leaf step = one 48 B triangle; + 48 B ray in, 16 B hit out" % (int(node_bytes), node_slabs)},
        "host_scene_build_s": host_s, "prepare_s_incl_pack_upload_lbvh": prepare_s,
    }
    del r, rc, st, stc
    return out


def extra_c4_node_formats(dev, steps: int = 2):
    """The persistent walk kernel of config 4 on its three node formats — binary 64 B (default), quantised 32 B
    (``quant_walk``), 4-wide 128 B (``wide_walk``) — and on the default format with the mesh's faces in Morton order: float
    sums of a 960x540 x 4 spp x depth 4 render compared bit for bit with the default, then the walk kernel's ms per
    1080p x 64 spp x depth 4 step (profiles/r2_c4_walk_kernel_analysis.md)."""
    import numpy as np
    import torch
    from b200rt import _lib, renderer, scenes
    from b200rt.scene_api import RenderSettings
    import copy
    from b200rt.packer import TriangleMesh
    lib = _lib.load()
    scene0, b = scenes.heightfield_scene()
    # the same mesh with its faces listed along a Morton curve (TriangleMesh.spatially_sorted): neighbours in space are
    # neighbours in the triangle records; closest hits are the same except for exact ties on shared edges
    scene1 = copy.copy(scene0)
    scene1.objects = [o.spatially_sorted() if isinstance(o, TriangleMesh) else o for o in scene0.objects]
    cam = b.create_camera(1920 / 1080)
    out, ref, ws = {}, None, None
    for name, kw, scene in (("binary_64B", {}, scene0), ("quantised_32B", {"quant_walk": True}, scene0),
                            ("wide_128B", {"wide_walk": True}, scene0), ("binary_64B_morton_face_order", {}, scene1)):
        r = renderer.B200PathTracer(precision="f32", seed=3, device=dev, distributed=False, **kw)
        acc, cnt = r.render_accum(scene, cam, RenderSettings(960, 540, 4, 4))
        if ref is None:
            ref = (acc, cnt)
        same = bool(np.array_equal(acc, ref[0]) and np.array_equal(cnt[:4], ref[1][:4]))
        differing = int((np.abs(acc[..., :3] - ref[0][..., :3]).max(axis=2) > 0).sum())
        if ws is not None:
            r._ws = ws
        st = r.prepare(scene, cam, RenderSettings(1920, 1080, 64, 4))
        ws = r._ws
        r.accumulate(st); torch.cuda.synchronize(dev)
        lib.b2rt_profile_enable(1)
        ms = (C.c_double * 8)(); nl = (C.c_int64 * 8)()
        lib.b2rt_profile_read(ms, nl)
        e0, e1 = _events()
        e0.record()
        for _ in range(steps):
            r.accumulate(st)
        e1.record(); torch.cuda.synchronize(dev)
        lib.b2rt_profile_read(ms, nl)
        lib.b2rt_profile_enable(0)
        out[name] = {"sums_identical_to_binary": same, "pixels_differing": differing, "walk_kernel_ms_per_step": ms[1] / steps,
                     "bounce0_and_shade_ms_per_step": ms[2] / steps, "ms_per_step": e0.elapsed_time(e1) / steps}
        del r, st
    return out


def extra_c5(scene, builder, dev, rank, world, local):
    """BASELINE config 5: 4K (3840x2160) Cornell box, 4096 spp, depth 8, spp split across the ranks + one NCCL reduce."""
    import torch
    import torch.distributed as td
    from b200rt import dist, renderer
    from b200rt.scene_api import RenderSettings
    W5, H5, S5 = 3840, 2160, 4096
    cam = builder.create_camera(W5 / H5)
    r = renderer.B200PathTracer(precision="f32", rng="pcg", seed=0, device=dev)
    warm = r.prepare(scene, cam, RenderSettings(W5, H5, 8 * world, DEPTH))
    r.accumulate(warm); torch.cuda.synchronize(dev)
    st = r.prepare(scene, cam, RenderSettings(W5, H5, S5, DEPTH))
    if world > 1:
        td.barrier(device_ids=[local])
    torch.cuda.synchronize(dev)
    e0, e1 = _events()
    e0.record()
    r.accumulate(st)
    r.finish(st)
    e1.record(); torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev)
    cnt = st["counters"].clone()
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX); td.all_reduce(cnt, op=td.ReduceOp.SUM)
    c, s = cnt.cpu().numpy(), float(t.item())
    del r, st, warm
    return {"workload": "cornell_path_3840x2160_4096spp_depth8", "n_gpus": world, "frame_s": s,
            "mpaths_per_s": float(c[0]) / s / 1e6, "mrays_per_s": float(c[1] + c[2] - c[5]) / s / 1e6, "steps": 1,
            "timing": "CUDA events around accumulate + NCCL reduce + resolve, max over ranks"}


def extra_reference_gpu(timeout_s: float = 300.0):
    """The UNMODIFIED reference GPU renderer (cuda_path_raytracer, numba.cuda JIT) from baseline/_ref on this GPU,
    in a subprocess (its own CUDA context; a numba failure cannot take the bench down)."""
    script = os.path.join(ROOT, "scripts", "reference_gpu.py")
    if not os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "main.py")):
        return {"unavailable": "baseline/_ref holds no reference checkout on this box"}
    try:
        p = subprocess.run([sys.executable, script, "16,256", "path-only"], capture_output=True, text=True, timeout=timeout_s)
        line = [l for l in p.stdout.splitlines() if l.startswith("{")]
        if not line:
            return {"unavailable": f"reference_gpu.py printed no result (rc={p.returncode}): {p.stderr[-300:]}"}
        d = json.loads(line[-1])
    except subprocess.TimeoutExpired:
        return {"unavailable": f"reference GPU renderer did not finish in {timeout_s:.0f} s"}
    except Exception as e:
        return {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    if "reference" not in d:
        return {"unavailable": d.get("reference_error", d.get("unavailable", "no reference result"))}
    ref = d["reference"]
    out = {"renderer": "cuda_path_raytracer (reference, numba.cuda JIT, unmodified, real JPEG textures)",
           "workload": f"cornell_path_{W}x{H}_depth{DEPTH}", "render_s": {k: v["render_s"] for k, v in ref.items()},
           "first_call_incl_jit_s": d.get("reference_first_call_incl_jit_s")}
    if "16" in ref and "256" in ref:
        # render() = host packing (constant per call) + kernel (linear in spp): the slope between two sample counts is
        # the kernel rate; each point is the faster of two calls
        slope = (ref["256"]["render_s"] - ref["16"]["render_s"]) / 240.0            # seconds per spp, kernel only
        out["kernel_mpaths_per_s"] = W * H / slope / 1e6 if slope > 0 else None
        out["e2e_mpaths_per_s_at_256spp"] = ref["256"]["Mpaths_per_s"]
        out["host_overhead_s_per_call"] = ref["16"]["render_s"] - 16 * slope
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--spp", type=int, default=SPP, help="debug only: a reduced spp makes the line invalid")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip extra_configs (C1, C3, C4, C5, f64, reference GPU)")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    t_process = time.perf_counter()
    import numpy as np
    import torch
    import torch.distributed as td

    from b200rt import _lib, dist, renderer
    from b200rt.scene_api import RenderSettings

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: b200rt has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        td.init_process_group("nccl", device_id=dev)
    spp = args.spp
    scene, camera, builder = build_scene()
    settings = RenderSettings(W, H, spp, DEPTH)

    def barrier():
        if world > 1:
            td.barrier(device_ids=[local])
        torch.cuda.synchronize(dev)

    # ---- the very first render() of this process: what a user's first call costs (library load, host packing of the
    # scene and the small-scene records, pinned texture block, 13 GB workspace allocation, first launches)
    e2e_cold = None
    if not args.no_e2e:
        barrier()
        t0 = time.perf_counter()
        r_cold = renderer.B200PathTracer(precision="f32", rng="pcg", seed=0, device=dev)
        img = r_cold.render(scene, camera, settings)
        barrier()
        e2e_cold = (time.perf_counter() - t0) * 1e3
        ws_shared = r_cold._ws
        del r_cold
    else:
        ws_shared = None
    lib = _lib.load()

    base = cpu_baseline(scene, camera) if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None

    # ---- N > 1: the NCCL-reduced image equals the single-GPU image (8 spp, same seeds) before anything is timed
    parity = None
    if world > 1:
        rs = renderer.B200PathTracer(precision="f32", rng="pcg", seed=0, device=dev)
        rs._ws = ws_shared
        acc_split, _ = rs.render_accum(scene, camera, RenderSettings(W, H, 8, DEPTH))
        img_split = rs.render(scene, camera, RenderSettings(W, H, 8, DEPTH))      # the fused reduce + resolve path
        if rank == 0:
            r1 = renderer.B200PathTracer(precision="f32", rng="pcg", seed=0, device=dev, distributed=False)
            r1._ws = ws_shared
            acc_one, _ = r1.render_accum(scene, camera, RenderSettings(W, H, 8, DEPTH))
            a, b = acc_split[..., :3].astype(np.float64), acc_one[..., :3].astype(np.float64)
            rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-3)
            img_one = r1.render(scene, camera, RenderSettings(W, H, 8, DEPTH))
            dimg = np.abs(np.asarray(img_split).astype(int) - np.asarray(img_one).astype(int))
            parity = {"spp": 8, "max_rel_diff": float(rel.max()),
                      "image_max_abs_diff_levels": int(dimg.max()), "image_bytes_differing": int((dimg > 0).sum()),
                      "fused_reduce_resolve": rs._symm is not None,
                      "ok": bool(np.allclose(a, b, rtol=1e-5, atol=1e-6) and dimg.max() <= 1),
                      "what": f"float sums of {world} ranks after ncclReduce vs the same 8 spp on rank 0 alone (rtol 1e-5); "
                              "8-bit image of the multi-GPU render() vs the single-GPU one (<= 1 level)"}
            del r1
        del rs
        barrier()

    r = renderer.B200PathTracer(precision="f32", rng="pcg", seed=0, device=dev)
    r._ws = ws_shared
    st = r.prepare(scene, camera, settings)                 # scene, textures and LBVH now resident in HBM

    def step():
        st["accum"].zero_()
        r.accumulate(st)
        r.finish(st)            # N > 1: fused reduce + resolve over peer memory (or one NCCL reduce + resolve on the root)
        r.frame_count += 1

    # clocks are sampled from the warm-up on (the same load): with 8 GPUs the timed region alone lasts < 100 ms
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    st["counters"].zero_()
    lib.b2rt_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    elapsed = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev)
    ms = (C.c_double * 8)(); nl = (C.c_int64 * 8)()
    _lib.check(lib.b2rt_profile_read(ms, nl), "b2rt_profile_read")
    lib.b2rt_profile_enable(0)
    cnt = st["counters"].clone()
    if world > 1:
        td.all_reduce(elapsed, op=td.ReduceOp.MAX)
        td.all_reduce(cnt, op=td.ReduceOp.SUM)
    elapsed_s = float(elapsed.item())
    cnt = cnt.cpu().numpy()
    paths, closest, shadow, by_hint = int(cnt[0]), int(cnt[1]), int(cnt[2]), int(cnt[5])
    assert paths == W * H * spp * args.steps, (paths, W * H * spp * args.steps)
    value = paths / elapsed_s / 1e6

    # ---- end to end through the public API (host scene in, PIL image out), texture upload included
    e2e = None
    if not args.no_e2e:
        r2 = renderer.B200PathTracer(precision="f32", rng="pcg", seed=0, device=dev)
        r2._ws = r._ws
        h2d = d2h = 0
        r2._tex_cache.enabled = False
        for _ in range(min(2, args.warmup)):                 # warm-up: pinned buffers, NCCL communicator
            r2.render(scene, camera, settings)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            r2._tex_cache.enabled = False                    # re-upload the textures every step
            img = r2.render(scene, camera, settings)
            h2d, d2h = r2.last_stats["h2d_bytes"], r2.last_stats["d2h_bytes"]
        barrier()
        t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            td.all_reduce(t_e2e, op=td.ReduceOp.MAX)
        e2e = {"value": W * H * spp * args.steps / float(t_e2e.item()) / 1e6, "unit": "Mpaths/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h if rank == 0 else 0),
               "ms_per_step": float(t_e2e.item()) / args.steps * 1e3,
               "e2e_cold_ms": e2e_cold,
               "e2e_cold_note": "first render() of the process (library load, host packing, pinned buffers, "
                                "workspace allocation, first launches) for the same 1024-spp frame"}
        if rank == 0:
            assert img is not None and img.size == (W, H)
        del r2

    # ---- this rank's numbers for the roofline objects (before the workspace is handed to the extra configurations)
    fused_reduce = st.get("symm") is not None
    c0 = st["counters"].cpu().numpy()
    n_box, n_loose, n_sph = int(st["ds"].struct.n_scan_boxes), int(st["ds"].struct.n_scan_loose), int(st["ds"].struct.n_sphere)
    wave = st["wave"]
    ws_gb = r._ws.numel() / 1e9

    extras = {}
    if not args.no_extras:
        del st
        r._ws = None
        ws_shared = None                                     # free the 13 GB wave state: C4 / C5 size their own
        torch.cuda.empty_cache()
        if rank == 0 and world == 1:
            for name, fn in (("c2_f64_parity", lambda: extra_c2_f64(scene, camera, dev)),
                             ("c3_whitted_texture", lambda: extra_c3(scene, builder, dev)),
                             ("c1_whitted_cpu_semantics", lambda: extra_c1(dev)),
                             ("c4_heightfield_1m_triangles", lambda: extra_c4(dev))):
                t0 = time.perf_counter()
                try:
                    extras[name] = fn()
                except Exception as e:                      # an extra must never take the headline line down
                    extras[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
                log(name, "%.1f s" % (time.perf_counter() - t0))
                torch.cuda.empty_cache()
        try:
            c5 = extra_c5(scene, builder, dev, rank, world, local)
        except Exception as e:
            c5 = {"error": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.empty_cache()
        if rank == 0:
            extras["c5_4k_4096spp"] = c5
        if rank == 0 and world == 1:
            t0 = time.perf_counter()
            extras["reference_gpu"] = extra_reference_gpu()
            log("reference_gpu", "%.1f s" % (time.perf_counter() - t0))
            ref = extras["reference_gpu"]
            if ref.get("kernel_mpaths_per_s"):
                ref["b200rt_over_reference_kernel"] = value / ref["kernel_mpaths_per_s"]
            t0 = time.perf_counter()                        # last: the opt-in node formats of the large-scene walk kernel
            try:
                extras["c4_walk_node_formats"] = extra_c4_node_formats(dev)
            except Exception as e:
                extras["c4_walk_node_formats"] = {"error": f"{type(e).__name__}: {e}"[:300]}
            log("c4_walk_node_formats", "%.1f s" % (time.perf_counter() - t0))
            torch.cuda.empty_cache()

    if rank == 0:
        peaks, peak_src = measured_peaks()
        tfl = C.c_double(0)
        _lib.check(lib.b2rt_fp32_peak(200000, C.byref(tfl), None), "b2rt_fp32_peak")
        classes = ["raygen", "extend", "bounce", "shadow", "accumulate"]
        total_ms = sum(ms[k] for k in range(5)) or 1.0
        dom = max(range(5), key=lambda k: ms[k])
        p0, close0, shad0, cull0 = float(c0[0]), float(c0[1]), float(c0[2]), float(c0[5])
        bounds0, hits0, lit0 = float(c0[6]), float(c0[7]), float(c0[3])
        queued0 = shad0 - cull0                              # shadow rays that reached the shadow kernel
        # algorithmic HBM bytes of the fused bounce kernel: every queued ray record is written once and read
        # once (bounce 0 generates its rays in registers), queued shadow records are written once, and every
        # path's radiance slot is initialised once
        bounce_bytes = (close0 - p0) * 2 * QUEUE_RECORD_BYTES + queued0 * QUEUE_RECORD_BYTES + p0 * 16.0
        shadow_bytes = queued0 * 32.0 + lit0 * (16.0 + 32.0)
        accum_bytes = p0 * 16.0 + (p0 / max(1, wave)) * 32.0
        alg_bytes = {2: bounce_bytes, 3: shadow_bytes}.get(dom, bounce_bytes)
        dom_gbs = alg_bytes / (ms[dom] * 1e-3) / 1e9
        kname = {2: "shade_kernel<float,PcgRng,MODE> (fused closest-hit + shade, one launch per bounce)",
                 3: "shadow_kernel<float>", 1: "extend_kernel<float>"}.get(dom, classes[dom])
        # ---- executed FP32 work of THIS rank: counted tests x canonical costs
        scan_cost = n_box * COST["box_record"] + n_loose * COST["rect"] + n_sph * COST["sphere"]
        masked_flops0 = float(c0[10])                        # camera rays: counted per-tile candidate-mask record tests
        ex = {
            "closest_hit_scans": ((close0 - p0) * scan_cost + masked_flops0) if masked_flops0 > 0
                                 else (close0 - bounds0) * scan_cost,
            "camera_rays_and_bounds_test": p0 * (COST["camera"] + COST["slab"]),
            "occluder_hint_tests": shad0 * COST["rect"],
            "shadow_scans_upper_bound": queued0 * scan_cost,
            "shading": hits0 * COST["shade"],
        }
        ex_total = sum(ex.values())
        step_s = elapsed_s                                   # all ranks run concurrently for elapsed_s
        useful_tf = (closest + shadow) * FLOPS_PER_RAY / elapsed_s / 1e12 / world
        measured_bytes_per_path = (NCU_BOUNCE_BYTES_PER_WAVE + NCU_SHADOW_BYTES_PER_WAVE + NCU_ACCUM_BYTES_PER_WAVE) / NCU_WAVE_PATHS
        alg_bytes_per_path = (bounce_bytes + shadow_bytes + accum_bytes) / max(1.0, p0)
        out = {
            "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_s / args.steps * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "width": W, "height": H, "spp": spp, "max_depth": DEPTH,
                       "parallelism": f"spp-split x{world} + " + ("fused reduce+resolve over NVLink peer memory" if fused_reduce else "1 NCCL reduce"),
                       "spp_per_wave": wave,
                       "l2": "inputs larger than L2 (wave state %.1f GB)" % ws_gb,
                       "scene": "34 primitives, 16 light points, 7 synthetic textures (52 MB RGB)"},
            "mrays_per_s": (closest + shadow - by_hint) / elapsed_s / 1e6,
            "mrays_answered_by_hint_per_s": by_hint / elapsed_s / 1e6,
            "rays_per_path": {"traced": (closest + shadow - by_hint) / paths, "answered_by_hint": by_hint / paths,
                              "closest_hit": closest / paths, "shadow_queued": (shadow - by_hint) / paths},
            "clocks": clocks,
            # what the bounce kernels themselves saw: cycle counter of CTA 0 against the global ns timer.  Some GPUs of the
            # pool run sustained FP32 load slower at an unchanged nvidia-smi reading; this makes such a box visible
            "sm_mhz_seen_by_kernels": (float(c0[11]) / float(c0[12]) * 1e3) if float(c0[12]) > 0 else None,
            "e2e": e2e,
            "gpu_launches": int(cnt[4]) // max(1, world) + args.steps,
            "roofline": {"bound": "hbm", "kernel": kname, "achieved": dom_gbs,
                         "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": dom_gbs / peaks["hbm_gbs"],
                         "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH, "traffic_source": NCU_SOURCE, "peak_source": peak_src,
                         "launches": int(nl[dom]), "avg_launch_ms": ms[dom] / max(1, nl[dom]),
                         "algorithmic_bytes_per_launch": alg_bytes / max(1, nl[dom]),
                         "share_of_step": ms[dom] / total_ms,
                         "note": "instruction-issue bound, not HBM bound: see fp32.executed"},
            "hbm": {"unit": "GB/s per GPU", "peak": peaks["hbm_gbs"],
                    "algorithmic_bytes_per_path": alg_bytes_per_path,
                    "achieved_algorithmic": (bounce_bytes + shadow_bytes + accum_bytes) / step_s / 1e9,
                    "measured_dram_bytes_per_path": measured_bytes_per_path,
                    "achieved_measured": p0 * measured_bytes_per_path / step_s / 1e9,
                    "frac_measured": p0 * measured_bytes_per_path / step_s / 1e9 / peaks["hbm_gbs"],
                    "measured_source": NCU_SOURCE + " (dram__bytes_read.sum + dram__bytes_write.sum, all kernels of one wave)"},
            "fp32": {"unit": "TFLOP/s per GPU", "peak": tfl.value,
                     "peak_source": "b2rt_fp32_peak FMA micro-benchmark, this run",
                     "executed": {"achieved": ex_total / step_s / 1e12, "frac": ex_total / step_s / 1e12 / tfl.value,
                                  "flops_per_path": ex_total / max(1.0, p0),
                                  "breakdown_flops_per_path": {k: v / max(1.0, p0) for k, v in ex.items()},
                                  "costs": COST, "records_per_scan": {"box": n_box, "planar": n_loose, "sphere": n_sph},
                                  "camera_ray_record_test_flops_per_path": masked_flops0 / max(1.0, p0),
                                  "camera_rays_in_empty_tiles_per_path": bounds0 / max(1.0, p0),
                                  "note": "counted tests x canonical costs; the kernels issue ~1 450 thread-instructions "
                                          "per path (ncu), most of them compare/select/logic, not FMA"},
                     "useful_reference_flops": {"achieved": useful_tf, "frac": useful_tf / tfl.value,
                                                "flops_per_ray": FLOPS_PER_RAY,
                                                "note": "what the REFERENCE's brute-force cuda_scene_hit spends on the same "
                                                        "rays (a work-equivalent like 2N^3 for a GEMM), not a utilisation"}},
            "kernel_ms_per_step": {c: ms[k] / args.steps for k, c in enumerate(classes)},
            "shadow_rays_culled_by_hint": by_hint,
            "cpu_baseline": base,
            "extra_configs": extras if extras else None,
        }
        if parity is not None:
            out["multi_gpu_parity"] = parity
            if not parity["ok"]:
                out["invalid"] = "the NCCL-reduced image differs from the single-GPU image"
        if spp != SPP:
            out["invalid"] = f"debug run at {spp} spp (the headline config is {SPP})"
        out["bench_wall_s"] = time.perf_counter() - t_process
        emit(out)
    if world > 1:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
