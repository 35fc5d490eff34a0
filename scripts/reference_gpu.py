"""Side measurement (not part of bench.py, the tests or the product): the UNMODIFIED reference GPU renderer
(`cuda_path_raytracer`, numba.cuda JIT) on the same B200, next to b200rt on the very same reference scene objects.

Needs the reference checkout under baseline/_ref (git-ignored; `cp -r /root/reference baseline/_ref` in the build
container — it travels to the GPU box with the snapshot).  Prints one JSON line."""
import json, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
if not os.path.isfile(os.path.join(REF, "main.py")):
    print(json.dumps({"unavailable": "baseline/_ref holds no reference checkout"})); sys.exit(0)
W, H, D = 1920, 1080, 8
spps = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "16,64").split(",")]
PATH_ONLY = "path-only" in sys.argv[2:]         # bench.py: only the reference path tracer, nothing of ours
os.chdir(REF); sys.path.insert(0, REF)
out = {"config": f"reference cuda_path_raytracer vs b200rt, {W}x{H}, depth {D}, reference scene objects + JPEG textures"}
try:
    from scene_builders.custom_scene_builder import CustomSceneBuilder
    from core.scene import RenderSettings
    from renderers.base_renderer import RendererFactory
    import renderers.cuda_path_tracer  # noqa: F401  (registers cuda_path_raytracer)
    random.seed(0)
    b = CustomSceneBuilder(); scene = b.build_scene(); cam = b.create_camera(W / H)
    ref = RendererFactory.create("cuda_path_raytracer")
    t0 = time.perf_counter(); ref.render(scene, cam, RenderSettings(W, H, 1, D)); out["reference_first_call_incl_jit_s"] = time.perf_counter() - t0
    out["reference"] = {}
    for spp in spps:
        dts = []
        for _ in range(2):                                        # the faster of two calls (host packing time varies)
            t0 = time.perf_counter(); img = ref.render(scene, cam, RenderSettings(W, H, spp, D)); dts.append(time.perf_counter() - t0)
        dt = min(dts)
        out["reference"][str(spp)] = {"render_s": dt, "Mpaths_per_s": W * H * spp / dt / 1e6, "all_s": dts}
    img.save(os.path.join(ROOT, "gpurun_out", "reference_gpu.png")) if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else None
except Exception as e:                                            # numba may not support this GPU / toolkit
    out["reference_error"] = f"{type(e).__name__}: {e}"[:400]
    scene = None
sys.path.insert(0, os.path.join(ROOT, "path-tracing__ray-tracer_b200"))
try:
    if scene is not None and not PATH_ONLY:
        import b200rt.renderer  # noqa: F401  registers into the REFERENCE's RendererFactory (plugin.py)
        ours = RendererFactory.create("b200_path_tracer")
        ours.render(scene, cam, RenderSettings(W, H, 8, D))
        out["b200rt"] = {}
        for spp in spps + [1024]:
            t0 = time.perf_counter(); img = ours.render(scene, cam, RenderSettings(W, H, spp, D)); dt = time.perf_counter() - t0
            out["b200rt"][str(spp)] = {"render_s": dt, "Mpaths_per_s": W * H * spp / dt / 1e6}
        img.save(os.path.join(ROOT, "gpurun_out", "b200rt_same_scene.png")) if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else None
except Exception as e:
    out["b200rt_error"] = f"{type(e).__name__}: {e}"[:400]
# ---- the reference's default renderer (cuda_texture_raytracer) at its golden setting: 2000x1500, 25 spp, depth 16
try:
    if scene is not None and not PATH_ONLY:
        import numpy as np
        import renderers.cuda_texture_renderer  # noqa: F401
        W2, H2, S2, D2 = 2000, 1500, 25, 16
        cam2 = b.create_camera(W2 / H2)
        tref = RendererFactory.create("cuda_texture_raytracer")
        t0 = time.perf_counter(); tref.render(scene, cam2, RenderSettings(W2, H2, 1, D2)); jit = time.perf_counter() - t0
        t0 = time.perf_counter(); img_ref = tref.render(scene, cam2, RenderSettings(W2, H2, S2, D2)); t_ref = time.perf_counter() - t0
        res = {"reference_render_s": t_ref, "reference_first_call_incl_jit_s": jit}
        for prec in ("f64", "f32"):
            tours = RendererFactory.create("b200_texture_raytracer", precision=prec)
            tours.render(scene, cam2, RenderSettings(W2, H2, S2, D2))
            t0 = time.perf_counter(); img = tours.render(scene, cam2, RenderSettings(W2, H2, S2, D2)); dt = time.perf_counter() - t0
            d = np.abs(np.asarray(img).astype(int) - np.asarray(img_ref).astype(int)).max(axis=2)
            res[prec] = {"render_s": dt, "pixels_differing": int((d > 0).sum()), "pixels_differing_by_more_than_1": int((d > 1).sum()),
                         "of": int(d.size)}
        out["texture_raytracer_2000x1500_25spp_d16"] = res
except Exception as e:
    out["texture_error"] = f"{type(e).__name__}: {e}"[:400]
print(json.dumps(out))
