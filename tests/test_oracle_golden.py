"""CPU tests: the C oracle against the golden vectors produced by the REAL reference
(oracle/make_golden.py).  These pin the oracle; the GPU tests then compare CUDA against it."""
import hashlib

import numpy as np
import pytest

from oracle import cpu_oracle as O


@pytest.fixture(scope="module")
def scene(cornell):
    return cornell[0]


@pytest.fixture(scope="module")
def packed(cornell):
    scene, b = cornell
    return O.nb_pack(scene, b.create_camera(16 / 9))


def test_packer_restatement_equals_reference_packers(cornell, golden_dir, packed):
    """nb_pack == _prepare_scene_data/_camera_data/_light_data/_texture_data (reference output)."""
    g = np.load(f"{golden_dir}/packed_scene_seed0.npz")
    scene, b = cornell
    assert np.array_equal(packed.scene, g["scene"])
    assert np.array_equal(packed.camera, g["camera_16x9"])
    assert np.array_equal(O.nb_pack(scene, b.create_camera(4 / 3), with_textures=False).camera, g["camera_4x3"])
    assert np.array_equal(packed.lights, g["lights"])
    assert np.array_equal(packed.tex_info, g["tex_info"])
    assert hashlib.sha256(packed.tex.tobytes()).hexdigest() == str(g["tex_sha256"])
    assert [type(o).__name__ for o in scene.objects] == list(g["object_kinds"])
    assert packed.scene.size == 815 and packed.tex.size == 52070958


def test_rng_and_tonemap_known_answers(golden_dir):
    g = np.load(f"{golden_dir}/nb_rng_tonemap.npz")
    assert [O.xorshift(int(s)) for s in g["seeds"]] == list(g["xorshift"])
    assert [O.lib().orc_nb_random(int(s)) for s in g["seeds"]] == list(g["random"])
    s, chain = int(g["chain"][0]), [int(g["chain"][0])]
    for _ in range(len(g["chain"]) - 1):
        s = O.xorshift(s)
        chain.append(s)
    assert chain == list(g["chain"])
    assert np.array_equal(np.array([O.tonemap(float(x)) for x in g["tonemap_in"]]), g["tonemap"])


def test_scene_hit_bit_exact(packed, golden_dir):
    g = np.load(f"{golden_dir}/nb_scene_hit_rays.npz")
    ids, rec = O.nb_scene_hit_rays(packed, g["o"], g["d"])
    assert np.array_equal((ids >= 0).astype(np.int32), g["hit"])
    assert np.array_equal(rec, g["rec"])


@pytest.mark.parametrize("fc", [0, 1])
def test_path_tracer_bit_exact(packed, golden_dir, fc):
    g = np.load(f"{golden_dir}/nb_path_64x36_spp8_d8_f{fc}.npz")
    W, H, SPP, D, _ = (int(v) for v in g["params"])
    r = O.nb_path_trace(packed, W, H, SPP, D, fc)
    assert np.array_equal(r["u8"], g["u8"])
    assert np.allclose(r["sum"], g["sum"], rtol=1e-13, atol=1e-14)
    assert np.allclose(r["sumsq"], g["sumsq"], rtol=1e-13, atol=1e-14)


def test_texture_whitted_bit_exact(cornell, golden_dir):
    scene, b = cornell
    for name in ("nb_texture_96x54_spp4_d6", "nb_texture_64x48_spp9_d16"):
        g = np.load(f"{golden_dir}/{name}.npz")
        W, H, SPP, D = (int(v) for v in g["params"])
        u8, _, _ = O.nb_whitted_texture(O.nb_pack(scene, b.create_camera(W / H)), W, H, SPP, D)
        assert np.array_equal(u8, g["u8"]), name


def test_cpu_whitted_bit_exact(cornell, golden_dir):
    """CPURenderer._trace through the reference's own BVH (exported by cpu_export)."""
    scene, b = cornell
    g = np.load(f"{golden_dir}/cpu_whitted_64x48_d4.npz")
    W, H, D = (int(v) for v in g["params"])
    exp = O.cpu_export(scene, b.create_camera(4 / 3))
    r = O.cpu_whitted(exp, W, H, D)
    assert np.abs(r["rgb"] - g["rgb"]).max() <= 1e-15
    assert np.array_equal(r["t"], g["t"])
    ids, t = O.cpu_primary_ids_bruteforce(exp, W, H)
    assert np.array_equal(ids, g["ids"])


def test_path_counters_match_survey(packed):
    """Work characterisation quoted in SURVEY 3.1 / BASELINE.md (16:9, depth 8)."""
    r = O.nb_path_trace(packed, 240, 135, 16, 8, 0, want_stats=False)
    n = 240 * 135 * 16
    c = r["counters"]
    assert abs(c["closest_rays"] / n + c["shadow_rays"] / n - 3.85) < 0.1
    assert abs(c["segments"] / n - 1.54) < 0.1 or abs(c["closest_rays"] / n - 2.31) < 0.1
    assert abs(c["nee_unshadowed"] / n - 0.083) < 0.01
