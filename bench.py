#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 path-tracing core.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload at every N: BASELINE.json configs[1] — Cornell box path trace, 1920x1080, 1024 spp,
max depth 8 (glass-sphere caustics), synthetic textures of the reference's dimensions.
A "step" renders that whole frame once.  With N > 1 (torchrun, one rank per GPU) the 1024 samples
per pixel are split across ranks (strong scaling: total work fixed) and summed with one NCCL reduce.

One JSON line on rank 0:
  value        Mpaths/s, inputs resident in HBM, CUDA-event time of K steps, max over ranks
  e2e          the same through the public renderer API: render(scene, camera, settings) -> PIL image
               (scene + texture H2D and image D2H inside the timed region)
  roofline     dominant kernel (fused closest-hit + shade bounce kernel), algorithmic queue bytes / measured launch time
  fp32         useful FP32 work (1 070 flop/ray) against the FMA peak measured in-run
  cpu_baseline the oracle port (C, float64, reference algorithm) on the host cores, bounded sample
--impl reference times that CPU implementation instead (all host threads, bounded sample per step).
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
STEP_BYTES_PER_PATH = 550.0     # whole-wavefront queue traffic per path
# dram__bytes_read.sum + dram__bytes_write.sum of the fused bounce kernel, averaged over the 8 bounce launches of one
# 32-spp wave at 1080p (ncu --set full, profiles/r1e_ncu_full_one_wave_32spp.csv: 13.08 GB per wave)
NCU_TRAFFIC_BYTES_PER_LAUNCH = 13.075816e9 / 8


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


def build_scene():
    from b200rt.cornell import CustomSceneBuilder
    random.seed(0)
    b = CustomSceneBuilder(texture_dir=False)
    return b.build_scene(), b.create_camera(W / H)


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
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(scene, camera, budget_s: float = 15.0, threads: int | None = None):
    """Oracle port (reference algorithm, float64, OpenMP) on a bounded sample of the same workload."""
    from oracle import cpu_oracle as O
    if threads:
        O.set_num_threads(threads)
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
    scene, camera = build_scene()
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--spp", type=int, default=SPP, help="debug only: a reduced spp makes the line invalid")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import numpy as np
    import torch
    import torch.distributed as td

    from b200rt import _lib, renderer
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
    scene, camera = build_scene()
    settings = RenderSettings(W, H, spp, DEPTH)
    lib = _lib.load()

    def barrier():
        if world > 1:
            td.barrier(device_ids=[local])
        torch.cuda.synchronize(dev)

    base = cpu_baseline(scene, camera) if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None

    r = renderer.B200PathTracer(precision="f32", rng="pcg", seed=0, device=dev)
    st = r.prepare(scene, camera, settings)                 # scene, textures and LBVH now resident in HBM

    def step():
        st["accum"].zero_()
        r.accumulate(st)
        from b200rt import dist
        dist.reduce_to_root(st["accum"])
        if rank == 0:
            r.resolve(st)
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
    paths, closest, shadow = int(cnt[0]), int(cnt[1]), int(cnt[2])
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
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h if rank == 0 else 0)}
        if rank == 0:
            assert img is not None and img.size == (W, H)

    if rank == 0:
        peaks, peak_src = measured_peaks()
        tfl = C.c_double(0)
        _lib.check(lib.b2rt_fp32_peak(200000, C.byref(tfl), None), "b2rt_fp32_peak")
        classes = ["raygen", "extend", "bounce", "shadow", "accumulate"]
        total_ms = sum(ms[k] for k in range(5)) or 1.0
        dom = max(range(5), key=lambda k: ms[k])
        c0 = st["counters"].cpu().numpy()                   # this rank's counters
        p0, close0, shad0, cull0 = int(c0[0]), int(c0[1]), int(c0[2]), int(c0[5])
        # algorithmic HBM bytes of the fused bounce kernel: every queued ray record is written once and read
        # once (bounce 0 generates its rays in registers), queued shadow records are written once, and every
        # path's radiance slot is initialised once
        bounce_bytes = (close0 - p0) * 2 * QUEUE_RECORD_BYTES + (shad0 - cull0) * QUEUE_RECORD_BYTES + p0 * 16.0
        shadow_bytes = (shad0 - cull0) * QUEUE_RECORD_BYTES
        alg_bytes = {2: bounce_bytes, 3: shadow_bytes}.get(dom, bounce_bytes)
        dom_gbs = alg_bytes / (ms[dom] * 1e-3) / 1e9
        kname = {2: "shade_kernel<float,PcgRng,MODE> (fused closest-hit + shade, one launch per bounce)",
                 3: "shadow_kernel<float>", 1: "extend_kernel<float>"}.get(dom, classes[dom])
        out = {
            "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_s / args.steps * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "width": W, "height": H, "spp": spp, "max_depth": DEPTH,
                       "parallelism": f"spp-split x{world} + 1 NCCL reduce", "spp_per_wave": st["wave"],
                       "l2": "inputs larger than L2 (wave state %.1f GB)" % (r._ws.numel() / 1e9),
                       "scene": "34 primitives, 16 light points, 7 synthetic textures (52 MB RGB)"},
            "mrays_per_s": (closest + shadow) / elapsed_s / 1e6,
            "rays_per_path": (closest + shadow) / paths,
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": int(cnt[4]) // max(1, world) + args.steps,
            "roofline": {"bound": "hbm", "kernel": kname, "achieved": dom_gbs,
                         "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": dom_gbs / peaks["hbm_gbs"],
                         "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH, "peak_source": peak_src, "launches": int(nl[dom]),
                         "avg_launch_ms": ms[dom] / max(1, nl[dom]),
                         "algorithmic_bytes_per_launch": alg_bytes / max(1, nl[dom]),
                         "share_of_step": ms[dom] / total_ms,
                         "note": "FP32-issue bound, not HBM bound: see fp32"},
            "step_hbm": {"achieved": paths / max(1, world) * STEP_BYTES_PER_PATH / elapsed_s / 1e9 * world,
                         "unit": "GB/s", "bytes_per_path": STEP_BYTES_PER_PATH},
            "fp32": {"achieved": (closest + shadow) * FLOPS_PER_RAY / elapsed_s / 1e12 / world, "unit": "TFLOP/s per GPU",
                     "peak": tfl.value, "frac": (closest + shadow) * FLOPS_PER_RAY / elapsed_s / 1e12 / world / tfl.value,
                     "peak_source": "b2rt_fp32_peak FMA micro-benchmark, this run", "flops_per_ray": FLOPS_PER_RAY},
            "kernel_ms_per_step": {c: ms[k] / args.steps for k, c in enumerate(classes)},
            "shadow_rays_culled_by_hint": int(cnt[5]),
            "cpu_baseline": base,
        }
        if spp != SPP:
            out["invalid"] = f"debug run at {spp} spp (the headline config is {SPP})"
        emit(out)
    if world > 1:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
