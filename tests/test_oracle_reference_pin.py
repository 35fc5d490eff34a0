"""Container-only pins (skipped wherever /root/reference is absent, e.g. on the GPU box):
the oracle against the reference's ONLY golden vector, output_RayTracer.png, and against the
reference renderers run live with the REAL textures."""
import os
import random

import numpy as np
import pytest

from oracle import cpu_oracle as O
from oracle import ref_harness as RH

pytestmark = pytest.mark.reference

if not RH.available():  # pragma: no cover
    pytest.skip("reference not mounted", allow_module_level=True)


@pytest.fixture(scope="module")
def ref_scene():
    scene, cam = RH.build_reference_scene(0, 2000 / 1500)       # the reference's own classes + real JPEGs
    return scene, cam


def test_oracle_reproduces_output_RayTracer_png(ref_scene):
    """output_RayTracer.png == cuda_texture_raytracer at main.py defaults (2000x1500, 25 spp, depth 16,
    main.py:33-40).  All 3 000 000 pixels, bit-exact."""
    from PIL import Image
    scene, cam = ref_scene
    gold = np.asarray(Image.open(os.path.join(RH.REF_ROOT, "output_RayTracer.png")).convert("RGB"))
    assert gold.shape == (1500, 2000, 3)
    pk = O.nb_pack(scene, cam)
    u8, _, _ = O.nb_whitted_texture(pk, 2000, 1500, 25, 16, want_float=False)
    img = u8[::-1]                                               # device rows are bottom-up (:782)
    ndiff = np.count_nonzero((img != gold).any(axis=2))
    assert ndiff == 0, f"{ndiff} of 3000000 pixels differ from output_RayTracer.png"


def test_oracle_matches_live_reference_path_tracer_real_textures(ref_scene):
    scene, _ = ref_scene
    import importlib
    cam = importlib.import_module("scene_builders.custom_scene_builder").CustomSceneBuilder().create_camera(16 / 9)
    pk_ref = RH.reference_pack(scene, cam, "path")
    pk = O.nb_pack(scene, cam)
    assert np.array_equal(pk.scene, pk_ref["scene"]) and np.array_equal(pk.tex, pk_ref["tex"])
    W, H, SPP, D = 48, 27, 4, 8
    u8 = RH.run_path_kernel(pk_ref, W, H, SPP, D, 0)
    r = O.nb_path_trace(pk, W, H, SPP, D, 0)
    assert np.array_equal(r["u8"].reshape(-1), u8)


def test_mirror_scene_api_renders_like_reference_cpu_renderer(ref_scene):
    """b200rt.scene_api objects driven through the oracle == the reference CPURenderer._trace."""
    import sys
    sys.path.insert(0, RH.REF_ROOT)
    scene, _ = ref_scene
    import importlib
    cam = importlib.import_module("scene_builders.custom_scene_builder").CustomSceneBuilder().create_camera(4 / 3)
    R = RH.reference_cpu_renderer()
    W, H, D = 24, 18, 3
    exp = O.cpu_export(scene, cam)
    got = O.cpu_whitted(exp, W, H, D)["rgb"]
    for j in range(0, H, 3):
        for i in range(0, W, 3):
            c = R._trace(cam.get_ray((i + 0.5) / W, (j + 0.5) / H), scene, 0, D)
            assert np.abs(got[j, i] - np.array([c.x, c.y, c.z])).max() <= 1e-15


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_oracle_matches_reference_on_random_scenes(seed, tmp_path):
    """Beyond the Cornell box: random scenes (skewed rectangles, odd radii, glass/mirror/diffuse, textured and
    untextured triangles) through the REAL reference vs the oracle — all three renderers, bit-exact."""
    import sys
    from PIL import Image
    sys.path.insert(0, RH.REF_ROOT)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import random_scenes as RS
    api = RS.reference_api()
    scene, cam = RS.make_scene(api, seed)
    os.makedirs(tmp_path / "textures", exist_ok=True)
    seen = set()
    for o in scene.objects:
        t = o.material.texture
        if t is not None and t.path not in seen:
            seen.add(t.path)
            Image.fromarray(t.pixels, "RGB").save(str(tmp_path / t.path), format="PNG")
    pk_ref = RH.reference_pack(scene, cam, "path", texture_root=str(tmp_path))
    pk = O.nb_pack(scene, cam)
    for k in ("scene", "camera", "lights", "tex", "tex_info"):
        assert np.array_equal(getattr(pk, k), pk_ref[k]), k
    W, H = 40, 30
    r = O.nb_path_trace(pk, W, H, 6, 6, 0)
    assert np.array_equal(r["u8"].reshape(-1), RH.run_path_kernel(pk_ref, W, H, 6, 6, 0))
    s1, _ = RH.run_path_float(pk_ref, W, H, 6, 6, 0)
    assert np.allclose(r["sum"].reshape(-1), s1, rtol=1e-12, atol=1e-13)
    u8, _, _ = O.nb_whitted_texture(pk, W, H, 4, 8)
    assert np.array_equal(u8.reshape(-1), RH.run_texture_kernel(pk_ref, W, H, 4, 8))
    R = RH.reference_cpu_renderer()
    got = O.cpu_whitted(O.cpu_export(scene, cam), W, H, 3)["rgb"]
    for j in range(0, H, 2):
        for i in range(0, W, 2):
            c = R._trace(cam.get_ray((i + 0.5) / W, (j + 0.5) / H), scene, 0, 3)
            assert np.abs(got[j, i] - np.array([c.x, c.y, c.z])).max() <= 1e-14, (i, j)
