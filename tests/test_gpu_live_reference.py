"""GPU parity against the LIVE reference GPU renderers (numba.cuda JIT on this device), when a reference checkout is
staged under the git-ignored ``baseline/_ref`` (``cp -r /root/reference baseline/_ref``; it travels to the GPU box
with the snapshot).  Skipped when it is absent or numba cannot JIT for the device — the committed golden fixtures
(tests/golden) and the oracle carry the parity claims; this is the same check against the real thing."""
import os
import random
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("no CUDA device", allow_module_level=True)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
if not os.path.isfile(os.path.join(REF, "renderers", "cuda_path_tracer.py")):
    pytest.skip("no reference checkout under baseline/_ref", allow_module_level=True)


@pytest.fixture(scope="module")
def live():
    """(reference module namespace, scene, builder) with cwd = the reference root (textures are opened by relative path)."""
    old_cwd, old_path = os.getcwd(), list(sys.path)
    os.chdir(REF)
    sys.path.insert(0, REF)
    for top in ("renderers", "core", "scene_builders"):       # a same-named foreign package must not shadow the reference
        for k in [k for k in sys.modules if k == top or k.startswith(top + ".")]:
            if not str(getattr(sys.modules[k], "__file__", "") or "").startswith(REF):
                del sys.modules[k]
    try:
        from scene_builders.custom_scene_builder import CustomSceneBuilder
        from core.scene import RenderSettings
        from renderers.base_renderer import RendererFactory
        import renderers.cuda_path_tracer  # noqa: F401
        import renderers.cuda_texture_renderer  # noqa: F401
        random.seed(0)
        b = CustomSceneBuilder()
        scene = b.build_scene()
        # JIT once here so that an unsupported device/toolkit skips instead of failing
        RendererFactory.create("cuda_texture_raytracer").render(scene, b.create_camera(4 / 3), RenderSettings(16, 12, 1, 2))
    except Exception as e:  # pragma: no cover
        os.chdir(old_cwd); sys.path[:] = old_path
        pytest.skip(f"the reference GPU renderers do not run here: {type(e).__name__}: {e}")
    yield RendererFactory, RenderSettings, scene, b
    os.chdir(old_cwd)
    sys.path[:] = old_path


def test_texture_raytracer_f64_equals_live_reference(live):
    """(b) deterministic renderer: the float64 instantiation against the reference's own GPU output, same scene objects."""
    from b200rt import renderer
    RendererFactory, RenderSettings, scene, b = live
    W, H, SPP, D = 400, 300, 9, 16
    cam = b.create_camera(W / H)
    ref = np.asarray(RendererFactory.create("cuda_texture_raytracer").render(scene, cam, RenderSettings(W, H, SPP, D)))
    got = np.asarray(renderer.B200TextureRaytracer(precision="f64").render(scene, cam, RenderSettings(W, H, SPP, D)))
    d = np.abs(got.astype(int) - ref.astype(int))
    assert np.count_nonzero(d) <= 6 and d.max() <= 1, f"{np.count_nonzero(d)} bytes differ, max {d.max()}"
    got32 = np.asarray(renderer.B200TextureRaytracer(precision="f32").render(scene, cam, RenderSettings(W, H, SPP, D)))
    d32 = np.abs(got32.astype(int) - ref.astype(int)).max(axis=2)
    assert (d32 <= 1).mean() > 0.99


def test_path_tracer_replays_live_reference(live):
    """(c) float64 + the reference's RNG replays the reference GPU path tracer sample for sample (frame 0)."""
    from b200rt import renderer
    RendererFactory, RenderSettings, scene, b = live
    W, H, SPP, D = 160, 90, 8, 8
    cam = b.create_camera(W / H)
    ref_r = RendererFactory.create("cuda_path_raytracer")
    ref = np.asarray(ref_r.render(scene, cam, RenderSettings(W, H, SPP, D)))
    ours = renderer.B200PathTracer(precision="f64", rng="reference")
    got = np.asarray(ours.render(scene, cam, RenderSettings(W, H, SPP, D)))
    same = (got == ref).all(axis=2).mean()
    assert same >= 0.99, f"only {same:.4f} of the pixels replay exactly"
    # and the float32 production kernels agree in energy with the live reference at a higher sample count
    n = 256
    ref2 = np.asarray(ref_r.render(scene, cam, RenderSettings(W, H, n, D))).astype(float)
    got2 = np.asarray(renderer.B200PathTracer(precision="f32", seed=3).render(scene, cam, RenderSettings(W, H, n, D))).astype(float)
    assert abs(ref2.mean() - got2.mean()) / ref2.mean() < 0.02
