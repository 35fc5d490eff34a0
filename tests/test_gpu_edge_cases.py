"""GPU parity beyond the Cornell box: random scenes, degenerate scenes, ragged sizes."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("no CUDA device", allow_module_level=True)

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import random_scenes as RS  # noqa: E402
from b200rt import packer, renderer  # noqa: E402
from b200rt.scene_api import Camera, Material, Plane, RenderSettings, Scene, Sphere, Triangle, Vec3  # noqa: E402
from oracle import cpu_oracle as O  # noqa: E402


def _rays(n, seed):
    rng = np.random.default_rng(seed)
    o = rng.uniform(-10, 10, (n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o, d


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_random_scene_intersection_bit_exact(seed):
    scene, cam = RS.make_scene(RS.local_api(), seed)
    pk = O.nb_pack(scene, cam)
    o, d = _rays(4000, seed)
    ref_ids, ref = O.nb_scene_hit_rays(pk, o, d)
    for use_bvh in (1, 0):
        ids, rec = renderer.trace_rays(scene, o, d, "numba", "f64", use_bvh=use_bvh)
        assert np.array_equal(ids, ref_ids)
        hit = ids >= 0
        assert np.array_equal(rec[hit], ref[hit][:, :9])
    # float32 kernels (LBVH walk and the small-scene scan records): same primitive, close distance
    for mode in (1, 2):
        ids32, rec32 = renderer.trace_rays(scene, o, d, "numba", "f32", use_bvh=mode)
        same = ids32 == ref_ids
        assert same.mean() > 0.995, (mode, np.count_nonzero(~same))
        both = same & (ref_ids >= 0)
        assert (np.abs(rec32[both, 0] - ref[both, 0]) / np.maximum(1, ref[both, 0])).max() < 5e-5


@pytest.mark.parametrize("seed", [11, 12])
def test_random_scene_renderers_match_oracle(seed):
    scene, cam = RS.make_scene(RS.local_api(), seed)
    pk = O.nb_pack(scene, cam)
    W, H = 48, 36
    # path tracer, float64 + reference RNG: exact replay
    ref = O.nb_path_trace(pk, W, H, 6, 6, 0)
    r = renderer.B200PathTracer(precision="f64", rng="reference", spp_per_wave=4)
    acc, cnt = r.render_accum(scene, cam, RenderSettings(W, H, 6, 6))
    ok = np.isclose(acc[..., :3], ref["sum"], rtol=1e-9, atol=1e-12).all(axis=2)
    assert ok.mean() >= 0.99, np.count_nonzero(~ok)
    # textured Whitted
    u8_ref, f_ref, _ = O.nb_whitted_texture(pk, W, H, 4, 8)
    f, u8 = renderer.B200TextureRaytracer(precision="f64").render_float(scene, cam, RenderSettings(W, H, 4, 8))
    assert np.abs(f - f_ref).max() <= 1e-4
    assert np.count_nonzero(u8 != u8_ref) <= 3
    # CPU-renderer semantics through the reference BVH vs the LBVH
    ref_cpu = O.cpu_whitted(O.cpu_export(scene, cam), W, H, 3)["rgb"]
    rgb = renderer.B200WhittedRenderer(precision="f64", jitter_seed=None).trace(scene, cam, W, H, 3)
    assert np.abs(rgb - ref_cpu).max() <= 1e-4


def test_random_scene_f32_production_statistics():
    """float32 + PCG + small-scene scan records + occluder hints on a scene with skewed rectangles and
    unpaired triangles: same Monte-Carlo tolerance as the Cornell test."""
    scene, cam = RS.make_scene(RS.local_api(), 12)
    W, H, D, n = 64, 48, 6, 1024
    ref = O.nb_path_trace(O.nb_pack(scene, cam), W, H, n, D)
    m_ref = ref["sum"] / n
    v_ref = np.maximum(ref["sumsq"] / n - m_ref ** 2, 0) * n / (n - 1)
    acc, cnt, sq = renderer.B200PathTracer(precision="f32", seed=5).render_accum(scene, cam, RenderSettings(W, H, n, D), want_sumsq=True)
    m = acc[..., :3].astype(np.float64) / n
    v = np.maximum(sq[..., :3].astype(np.float64) / n - m ** 2, 0) * n / (n - 1)
    lit = (v_ref > 1e-12) & (v > 1e-12)
    s2 = (v_ref + v) / n
    z2 = (m - m_ref) ** 2 / np.where(lit, s2, 1)
    assert 0.75 < z2[lit].mean() < 1.3, z2[lit].mean()
    assert (np.abs(m - m_ref)[lit] <= 3 * np.sqrt(s2[lit])).mean() >= 0.99
    assert abs(m.mean() - m_ref.mean()) / m_ref.mean() < 0.03
    ref_rpp = (ref["counters"]["closest_rays"] + ref["counters"]["shadow_rays"]) / (W * H * n)
    # shadow rays with an exactly-zero payload are never traced, so the GPU count is a little lower
    assert 0.8 * ref_rpp < (cnt[1] + cnt[2]) / cnt[0] <= 1.02 * ref_rpp


def _cam():
    return Camera(Vec3(0, 0, 20.0), Vec3(0, 0, 0), Vec3(0, 1, 0), 40.0, 33 / 17)


def test_empty_scene():
    scene = Scene()
    W, H = 33, 17
    img = np.asarray(renderer.B200PathTracer().render(scene, _cam(), RenderSettings(W, H, 3, 4)))
    assert img.shape == (H, W, 3) and (img == 32).all()          # ACES(0.1) * 255 -> 32, the reference's sky
    _, u8 = renderer.B200TextureRaytracer(precision="f64").render_float(scene, _cam(), RenderSettings(W, H, 4, 4))
    assert (u8 == 0).all()
    obj, t, pid = renderer.primary_hits(scene, _cam(), W, H)
    assert (pid == -1).all()


@pytest.mark.parametrize("kind", ["sphere", "plane", "triangle"])
def test_single_primitive_scene(kind):
    """n_prims == 1: the LBVH root is a leaf."""
    scene = Scene()
    m = Material(Vec3(0.8, 0.3, 0.2), diffuse=0.7, specular=0.2)
    if kind == "sphere":
        scene.add_object(Sphere(Vec3(0.5, 0.2, 0), 3.0, m))
    elif kind == "plane":
        scene.add_object(Plane(Vec3(-4, -3, 0), Vec3(0, 0, 1), Vec3(8, 0, 0), Vec3(0, 6, 0), 8, 6, m))
    else:
        scene.add_object(Triangle(Vec3(-4, -3, 0), Vec3(4, -3, 1), Vec3(0, 4, -1), None, None, None, m))
    scene.add_light_sample(Vec3(3, 6, 10))
    cam = _cam()
    W, H = 33, 17
    pk = O.nb_pack(scene, cam)
    ref_ids, ref_t = O.nb_primary_hits(pk, W, H)
    _, t, pid = renderer.primary_hits(scene, cam, W, H, "numba", "f64")
    assert np.array_equal(pid, ref_ids) and np.array_equal(t, ref_t) and (pid >= 0).any()
    ref = O.nb_path_trace(pk, W, H, 5, 4, 0)
    r = renderer.B200PathTracer(precision="f64", rng="reference", spp_per_wave=2)       # 5 spp in waves of 2, 2, 1
    acc, cnt = r.render_accum(scene, cam, RenderSettings(W, H, 5, 4))
    assert np.isclose(acc[..., :3], ref["sum"], rtol=1e-9, atol=1e-12).all(axis=2).mean() >= 0.99
    assert cnt[0] == W * H * 5
    img = renderer.B200PathTracer(precision="f32").render(scene, cam, RenderSettings(W, H, 4, 1))      # depth 1
    assert img.size == (W, H)


def test_no_lights_and_depth_one():
    scene, cam = RS.make_scene(RS.local_api(), 13, n_lights=0)
    pk = O.nb_pack(scene, cam)
    W, H = 40, 30
    for depth in (1, 5):
        ref = O.nb_path_trace(pk, W, H, 4, depth, 2)
        r = renderer.B200PathTracer(precision="f64", rng="reference")
        r.frame_count = 2
        acc, cnt = r.render_accum(scene, cam, RenderSettings(W, H, 4, depth))
        assert np.isclose(acc[..., :3], ref["sum"], rtol=1e-9, atol=1e-12).all(axis=2).mean() >= 0.99
        assert cnt[2] == 0                                           # no shadow rays without lights


def test_more_than_64_primitives_walks_the_lbvh():
    """Above the small-scene threshold the float32 path uses the LBVH for every ray; still the same image
    statistics as the generic scan of the float64 oracle."""
    scene, cam = RS.make_scene(RS.local_api(), 21, n_rect=6, n_sphere=10, n_tri=70, with_textures=True)
    assert len(scene.objects) > 64
    pk = O.nb_pack(scene, cam)
    o, d = _rays(3000, 5)
    ref_ids, ref = O.nb_scene_hit_rays(pk, o, d)
    ids, rec = renderer.trace_rays(scene, o, d, "numba", "f64")
    assert np.array_equal(ids, ref_ids) and np.array_equal(rec[ids >= 0], ref[ids >= 0][:, :9])
    W, H, n = 48, 36, 256
    oref = O.nb_path_trace(pk, W, H, n, 5)
    acc, cnt = renderer.B200PathTracer(precision="f32", seed=2).render_accum(scene, cam, RenderSettings(W, H, n, 5))
    m, m_ref = acc[..., :3] / n, oref["sum"] / n
    assert abs(m.mean() - m_ref.mean()) / m_ref.mean() < 0.05
    r64 = renderer.B200PathTracer(precision="f64", rng="reference")
    acc64, _ = r64.render_accum(scene, cam, RenderSettings(W, H, 4, 5))
    o4 = O.nb_path_trace(pk, W, H, 4, 5)
    assert np.isclose(acc64[..., :3], o4["sum"], rtol=1e-9, atol=1e-12).all(axis=2).mean() >= 0.99


def test_persistent_walk_kernel_equals_fused_walk():
    """Large-scene bounces: the persistent walk kernel (dynamic ray fetch, majority scheduling) + wavefront shade
    stage must reproduce the fused per-ray walk bit for bit — same closest hits, same RNG streams, same sums."""
    from b200rt import scenes
    from b200rt.scene_api import RenderSettings
    scene, b = scenes.heightfield_scene(nx=201, nz=101)            # 40 000 triangles: walks the LBVH
    cam = b.create_camera(16 / 9)
    st = RenderSettings(320, 180, 8, 4)
    out = {}
    for fused_walk, wide in ((False, True), (False, False), (True, True)):    # 4-wide nodes, binary nodes, fused per-ray walk
        for sort in (True, False):
            r = renderer.B200PathTracer(precision="f32", rng="pcg", seed=5, fused_walk=fused_walk, sort_rays=sort,
                                        wide_walk=wide)
            acc, cnt = r.render_accum(scene, cam, st)
            out[(fused_walk, sort, wide)] = (acc, cnt)
    ref_acc, ref_cnt = out[(True, False, True)]
    for key, (acc, cnt) in out.items():
        assert np.array_equal(cnt[:4], ref_cnt[:4]), (key, cnt, ref_cnt)
        # per-pixel sums are added in path order inside a wave, which the ray sort does not change (L[slot])
        assert np.array_equal(acc, ref_acc), key


def test_wide_nodes_hold_the_grandchildren():
    """b2rt_lbvh_widen: entry `ref` of the 4-wide array holds exactly the boxes and references of the grandchildren of
    binary node `ref` (a leaf child stays one slot), and the walk over them counts fewer box steps than the binary walk."""
    from b200rt import _lib, scenes
    from b200rt.device import DeviceScene
    from b200rt.packer import pack_scene
    scene, b = scenes.heightfield_scene(nx=101, nz=51)              # 10 000 triangles
    ds = DeviceScene(pack_scene(scene, "numba"), _lib.P_F32, wide_nodes=True)
    assert ds.wide is not None and ds.struct.d_bvh_wide
    torch.cuda.synchronize()
    n_top, n_int = ds.n_top, ds.n_internal
    nodes = ds.nodes.cpu().numpy().reshape(-1, 16)[:n_int]
    top = ds.top.cpu().numpy().reshape(-1, 16)[:n_top]
    wide = ds.wide.cpu().numpy().view(np.float32).reshape(-1, 32)
    assert wide.shape[0] == n_top + n_int

    def record(ref):
        rec = top[ref] if ref < n_top else nodes[ref - n_top]
        refs = rec[12:14].view(np.int32)
        return ((rec[0:3], rec[3:6], int(refs[0])), (rec[6:9], rec[9:12], int(refs[1])))

    empty = -(1 << 31)
    seen, frontier = 0, [ds.root]
    while frontier:                                                  # every wide node the walk can reach from the root
        ref = frontier.pop()
        want = []
        for lo, hi, c in record(ref):
            want += [(lo, hi, c)] if c < 0 else list(record(c))
        w = wide[ref]
        refs = w[24:28].view(np.int32)
        assert [int(x) for x in refs[:len(want)]] == [c for _, _, c in want]
        assert all(int(x) == empty for x in refs[len(want):])
        for k, (lo, hi, c) in enumerate(want):
            assert np.array_equal(w[[k, 4 + k, 8 + k]], lo) and np.array_equal(w[[12 + k, 16 + k, 20 + k]], hi)
            if c >= 0:
                frontier.append(c)
        seen += 1
    assert seen > n_int // 4                                         # about half the tree's levels
    # the counted walk: fewer box steps with the wide nodes, the same leaf tests or fewer, identical results
    cam = b.create_camera(16 / 9)
    st = RenderSettings(160, 90, 4, 3)
    steps = {}
    for wide_walk in (True, False):
        r = renderer.B200PathTracer(precision="f32", rng="pcg", seed=2, count_tests=True, wide_walk=wide_walk)
        acc, cnt = r.render_accum(scene, cam, st)
        steps[wide_walk] = (acc, cnt)
    assert np.array_equal(steps[True][0], steps[False][0])
    assert np.array_equal(steps[True][1][:4], steps[False][1][:4])
    assert 0 < steps[True][1][8] < 0.75 * steps[False][1][8], (steps[True][1][8:10], steps[False][1][8:10])


def test_progressive_accumulation_continues_the_sample_sequence(tmp_path):
    """progressive=True: two render() calls of 8 spp resolve the SAME 16 global samples as one 16-spp render
    (the accumulation the reference's frame_count reseed only hints at, cuda_path_tracer.py:28,739,809)."""
    import random
    from b200rt.cornell import CustomSceneBuilder
    random.seed(0)
    b = CustomSceneBuilder(texture_dir=False)
    scene, cam = b.build_scene(), b.create_camera(16 / 9)
    W, H, D = 160, 90, 6
    one = renderer.B200PathTracer(precision="f32", rng="pcg", seed=9)
    full, _ = one.render_accum(scene, cam, RenderSettings(W, H, 16, D))
    prog = renderer.B200PathTracer(precision="f32", rng="pcg", seed=9, progressive=True)
    a, _ = prog.render_accum(scene, cam, RenderSettings(W, H, 8, D))
    a = a.copy()
    b2, _ = prog.render_accum(scene, cam, RenderSettings(W, H, 8, D))
    assert not np.array_equal(a, b2)                       # the second call added its samples
    assert np.allclose(b2, full, rtol=1e-5, atol=1e-5)
    img16 = np.asarray(one.render(scene, cam, RenderSettings(W, H, 16, D)))
    prog.reset()
    prog.render(scene, cam, RenderSettings(W, H, 8, D))
    img_prog = np.asarray(prog.render(scene, cam, RenderSettings(W, H, 8, D)))
    assert img_prog.shape == img16.shape
    # `one` advanced its frame_count between calls (new seed), so only the statistics agree there; the progressive
    # image must equal a fresh 16-spp render with the same seed
    fresh = np.asarray(renderer.B200PathTracer(precision="f32", rng="pcg", seed=9).render(scene, cam, RenderSettings(W, H, 16, D)))
    assert np.abs(img_prog.astype(int) - fresh.astype(int)).max() <= 1
    prog.reset()
    c, _ = prog.render_accum(scene, cam, RenderSettings(W, H, 8, D))
    assert np.array_equal(c, a)                            # reset() starts the sequence again


def test_cli_renders_with_the_reference_flags(tmp_path):
    """b200rt.cli: the reference's main.py flags (-r/-w/--height/--path-samples/-d/-o) produce a PNG of that size."""
    from PIL import Image
    import b200rt.cli as cli
    out = tmp_path / "cli.png"
    rc = cli.main(["-r", "b200_path_tracer", "-w", "96", "--height", "54", "--path-samples", "4", "-d", "3",
                   "-o", str(out), "--stats"])
    assert rc == 0
    img = Image.open(out)
    assert img.size == (96, 54) and img.mode == "RGB"
    out2 = tmp_path / "cli2.png"
    assert cli.main(["-r", "b200_texture_raytracer", "-w", "64", "--height", "48", "-s", "4", "-d", "4", "-o", str(out2)]) == 0
    assert Image.open(out2).size == (64, 48)


def _sheared_open_box_scene(seed=3):
    """A sheared parallelepiped made of triangle pairs with ONE face missing, a rotated cube made of rectangles,
    two spheres and a few loose triangles."""
    rng = np.random.default_rng(seed)
    sc = Scene()
    m = Material(Vec3(0.7, 0.6, 0.5), diffuse=0.8)
    c = np.array([1.0, -2.0, 0.5])
    h = [np.array([3.0, 0.4, 0.2]), np.array([0.5, 2.5, -0.3]), np.array([-0.2, 0.6, 2.0])]      # sheared half axes
    V = lambda p: Vec3(*[float(x) for x in p])
    for k in range(3):
        i, j = (k + 1) % 3, (k + 2) % 3
        for sgn in (-1, 1):
            if k == 2 and sgn == 1:
                continue                                     # the open face
            p0 = c + sgn * h[k] - h[i] - h[j]
            p1, p3 = p0 + 2 * h[i], p0 + 2 * h[j]
            p2 = p1 + p3 - p0
            sc.add_object(Triangle(V(p0), V(p1), V(p2), None, None, None, m))
            sc.add_object(Triangle(V(p0), V(p2), V(p3), None, None, None, m))
    # rotated cube of rectangles
    a = 0.6
    R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    cc, s = np.array([-5.0, 1.0, -1.0]), 1.5
    ax = [R[:, 0], R[:, 1], R[:, 2]]
    for k in range(3):
        i, j = (k + 1) % 3, (k + 2) % 3
        for sgn in (-1, 1):
            anchor = cc + sgn * s * ax[k] - s * ax[i] - s * ax[j]
            sc.add_object(Plane(V(anchor), V(sgn * ax[k]), V(ax[i]), V(ax[j]), 2 * s, 2 * s, m))
    sc.add_object(Sphere(Vec3(0.5, -2.0, 0.3), 0.8, m))       # inside the sheared box
    sc.add_object(Sphere(Vec3(4.0, 4.0, 4.0), 1.1, m))
    for _ in range(3):
        q = rng.uniform(-6, 6, 3)
        p = [V(q + rng.normal(scale=1.5, size=3)) for _ in range(3)]
        sc.add_object(Triangle(p[0], p[1], p[2], None, None, None, m))
    sc.lights = [Vec3(0.0, 8.0, 0.0)]
    return sc


def test_box_records_on_sheared_open_box_and_rotated_cube():
    """group_scan_boxes on a sheared parallelepiped of triangle pairs with a missing face and a rotated cube of
    rectangles: two box records; the box scan equals the one-by-one planar scan and the float64 generic scan for
    rays from outside, from inside either box and through the open face."""
    sc = _sheared_open_box_scene()
    pk = packer.pack_scene(sc, "numba")
    quads = []
    rec = packer.build_scan_prims(pk, quads_out=quads)
    rec2, n_loose, boxes = packer.group_scan_boxes(rec, quads)
    boxes = boxes.reshape(-1, 4, 4)
    assert boxes.shape[0] == 2
    codes = [sorted(int(x) for x in b[3, :2].view(np.uint8)[:6]) for b in boxes]
    assert sorted(sum(1 for x in cds if x != 255) for cds in codes) == [5, 6]
    assert n_loose == 3                                        # the loose triangles
    rng = np.random.default_rng(2)
    n = 120000
    o = rng.uniform(-9, 9, size=(n, 3))
    o[: n // 4] = np.array([1.0, -2.0, 0.5]) + rng.uniform(-0.8, 0.8, size=(n // 4, 3))      # inside the sheared box
    o[n // 4: n // 2] = np.array([-5.0, 1.0, -1.0]) + rng.uniform(-1.0, 1.0, size=(n // 4, 3))   # inside the cube
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    a, ra = renderer.trace_rays(sc, o, d, "numba", "f32", use_bvh=2, scan_boxes=True)
    b, rb = renderer.trace_rays(sc, o, d, "numba", "f32", use_bvh=2, scan_boxes=False)
    g, rg = renderer.trace_rays(sc, o, d, "numba", "f64", use_bvh=0)
    assert (a == b).mean() > 0.9995 and (a == g).mean() > 0.999
    both = (a == g) & (g >= 0)
    cos = np.abs((rg[both, 4:7] * d[both]).sum(1))
    tol = 2e-5 * np.maximum(1.0, rg[both, 0]) + 8e-6 / np.maximum(cos, 1e-6)
    assert (np.abs(ra[both, 0] - rg[both, 0]) <= tol).all()
    assert (g >= 0).mean() > 0.5
    occ, _ = renderer.trace_rays(sc, o, d, "numba", "f32", use_bvh=2, any_hit=True, scan_boxes=True)
    assert np.mean((occ >= 0) == (g >= 0)) > 0.9995


def test_rectangles_outside_the_hierarchy_give_identical_hits():
    """Large mesh + a few room-sized rectangles: the walls stay outside the LBVH (tested directly before every walk).
    Closest and any-hit results equal the all-in-one hierarchy and the brute-force scan."""
    from b200rt import scenes
    scene, b = scenes.heightfield_scene(nx=121, nz=81)             # 19 200 triangles + 5 walls
    pk = packer.pack_scene(scene, "numba")
    rng = np.random.default_rng(8)
    n = 60000
    o = rng.uniform(-14, 14, (n, 3)); o[:, 1] = rng.uniform(-14.5, 14, n)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    a, ra = renderer.trace_rays(scene, o, d, "numba", "f32", use_bvh=True, packed=pk, rects_outside=True)
    c, rc = renderer.trace_rays(scene, o, d, "numba", "f32", use_bvh=True, packed=pk, rects_outside=False)
    g, rg = renderer.trace_rays(scene, o, d, "numba", "f32", use_bvh=False, packed=pk)
    assert np.array_equal(a, g) and np.array_equal(c, g)
    assert np.array_equal(ra[:, 0], rg[:, 0]) and np.array_equal(rc[:, 0], rg[:, 0])
    assert (g >= 0).mean() > 0.8 and ((g >= 0) & (g < 5)).mean() > 0.2    # walls (ids 0..4) are hit too
    occ, _ = renderer.trace_rays(scene, o, d, "numba", "f32", use_bvh=True, any_hit=True, packed=pk, rects_outside=True)
    assert np.array_equal(occ >= 0, g >= 0)
    st = RenderSettings(160, 90, 4, 4)
    cam = b.create_camera(16 / 9)
    x, cx = renderer.B200PathTracer(precision="f32", seed=2, rects_outside=True).render_accum(scene, cam, st)
    y, cy = renderer.B200PathTracer(precision="f32", seed=2, rects_outside=False).render_accum(scene, cam, st)
    assert np.array_equal(x, y) and np.array_equal(cx[:4], cy[:4])


@pytest.mark.parametrize("kind", ["coincident", "flat", "coplanar_overlap", "line", "walls_and_spheres"])
def test_lbvh_builder_degenerate_inputs(kind):
    """Morton bit allocation, rotations and outside-rectangles on degenerate large inputs: LBVH walk == brute force.
    coincident: 6 000 copies of one triangle (identical centroids: ties resolved by index bits);
    flat: a mesh in one plane (zero spread on an axis); coplanar_overlap: overlapping triangles in one plane; line: centroids on a line (zero spread on two axes);
    walls_and_spheres: 6 000 triangles + 8 room-sized rectangles kept outside the hierarchy + spheres inside it."""
    rng = np.random.default_rng(4)
    m = Material(Vec3(0.7, 0.7, 0.7), diffuse=0.8)
    sc = Scene()
    n = 6000
    if kind == "coincident":
        v = np.tile(np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.2]]), (n, 1))
    elif kind == "flat":
        # a regular grid of separated triangles (overlapping coplanar triangles would be exact-t ties, which the
        # float32 kernels only resolve to the lowest id in scan order — the documented tie cases)
        gx, gz = np.meshgrid(np.arange(100) * 0.2 - 10, np.arange(60) * 0.33 - 10)
        c = np.stack([gx.ravel(), np.zeros(n), gz.ravel()], 1)
        v = (c[:, None, :] + np.array([[0, 0, 0], [0.15, 0, 0], [0, 0, 0.25]])[None]).reshape(-1, 3)
    elif kind == "coplanar_overlap":
        # random OVERLAPPING triangles in one plane: rays hit several of them at float-equal distances, and the
        # walk (tree order) must keep the same lowest id as the scan (id order)
        c = np.stack([rng.uniform(-10, 10, n), np.zeros(n), rng.uniform(-10, 10, n)], 1)
        v = (c[:, None, :] + np.array([[0, 0, 0], [0.3, 0, 0], [0, 0, 0.3]])[None]).reshape(-1, 3)
    elif kind == "line":
        c = np.stack([np.linspace(-10, 10, n), np.zeros(n), np.zeros(n)], 1)
        v = (c[:, None, :] + np.array([[0, -0.5, -0.5], [0, 0.5, -0.5], [0, 0.0, 0.5]])[None]).reshape(-1, 3)
    else:
        c = rng.uniform(-8, 8, (n, 3))
        v = (c[:, None, :] + rng.normal(scale=0.2, size=(n, 3, 3))).reshape(-1, 3)
        for k in range(8):
            a = rng.uniform(-12, -10, 3)
            sc.add_object(Plane(Vec3(*a), Vec3(0, 1, 0), Vec3(1, 0, 0), Vec3(0, 0, 1), 22.0 + k, 21.0, m))
        for k in range(5):
            sc.add_object(Sphere(Vec3(*rng.uniform(-6, 6, 3)), 0.7, m))
    sc.objects.append(packer.TriangleMesh(v, np.arange(3 * n).reshape(-1, 3), m))
    pk = packer.pack_scene(sc, "numba")
    q = 20000
    o = rng.uniform(-12, 12, (q, 3))
    d = rng.normal(size=(q, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    if kind in ("coincident", "line"):
        d = (rng.uniform(-0.5, 0.5, (q, 3)) + [0.3, 0.3, 0.0]) - o * [0 if kind == "line" else 1, 1, 1]      # aim at the geometry
        d /= np.linalg.norm(d, axis=1, keepdims=True)
    a_ids, a_rec = renderer.trace_rays(sc, o, d, "numba", "f32", use_bvh=True, packed=pk)
    b_ids, b_rec = renderer.trace_rays(sc, o, d, "numba", "f32", use_bvh=False, packed=pk)
    assert np.array_equal(a_ids, b_ids)
    assert np.array_equal(a_rec[:, 0], b_rec[:, 0])
    assert (b_ids >= 0).mean() > 0.02
    occ, _ = renderer.trace_rays(sc, o, d, "numba", "f32", use_bvh=True, any_hit=True, packed=pk)
    assert np.array_equal(occ >= 0, b_ids >= 0)


# ------------------------------------------------------------------------------------ every bounce-kernel mode
# the scan records were staged into shared memory by launches that had not paid for them
# (MODE 0 / 1 / 2 / 4).  Each combination of documented kwargs that reaches one of those modes on a small float32
# scene with occluder hints is rendered here and compared with the default kernels (same estimator, same RNG
# streams: only rounding differs) — an out-of-range shared write faults or corrupts the image.
@pytest.mark.parametrize("kwargs", [
    dict(rng="reference"),                       # raygen + MODE 1 at bounce 0, MODE 3 afterwards
    dict(fused=False),                           # extend_kernel + MODE 0 (0 B of dynamic shared memory)
    dict(surface_records=False),                 # MODE 4 at bounce 0, MODE 2 (generic scan + hint records) afterwards
    dict(primary_walk=True),                     # MODE 6: top levels + scan + surface records
    dict(occluder_hints=False),
    dict(scan_boxes=False),
    dict(fused=False, surface_records=False),
    dict(top_nodes=0),
    dict(primary_masks=False),                   # camera rays scan every record behind the scene-bounds test
])
def test_cornell_f32_every_kernel_mode_agrees(cornell, kwargs):
    scene, b = cornell
    W, H, n, D = 96, 54, 64, 8
    cam = b.create_camera(W / H)
    base, cnt0 = renderer.B200PathTracer(precision="f32", seed=3).render_accum(scene, cam, RenderSettings(W, H, n, D))
    acc, cnt = renderer.B200PathTracer(precision="f32", seed=3, **kwargs).render_accum(scene, cam, RenderSettings(W, H, n, D))
    assert np.isfinite(acc).all()
    assert cnt[0] == W * H * n
    sky = base[..., :3].max(axis=2) == np.float32(0.1) * n          # pixels whose every sample missed: exact
    m0, m1 = base[..., :3].mean() / n, acc[..., :3].mean() / n
    # an independent RNG (the reference's xorshift) only agrees statistically; everything else replays the same paths
    assert abs(m1 - m0) / m0 < (0.12 if kwargs.get("rng") == "reference" else 0.02), (m0, m1)
    if "primary_masks" in kwargs:
        # the per-tile candidate masks only skip records that no ray of the tile can hit: bit-identical image
        assert np.array_equal(acc, base) and np.array_equal(cnt[:6], cnt0[:6])
    if kwargs.get("rng") != "reference":
        # same counter-based streams: paths differ only where float32 rounding flips a decision
        assert np.array_equal(acc[..., :3][sky], base[..., :3][sky])
        close = np.isclose(acc[..., :3], base[..., :3], rtol=2e-2, atol=1e-3).all(axis=2)
        assert close.mean() > 0.9, close.mean()
        assert abs(int(cnt[1]) - int(cnt0[1])) / cnt0[1] < 0.01


def test_large_top_level_copy_needs_smem_opt_in():
    """top_nodes = 1024 -> 64 KB of dynamic shared memory per CTA, above the 48 KB default: every walking kernel
    opts in with cudaFuncSetAttribute and finds the same hits as top_nodes = 0."""
    rng = np.random.default_rng(5)
    n = 6000
    v = rng.uniform(-8, 8, (n, 1, 3)) + rng.normal(scale=0.4, size=(n, 3, 3))
    scene = Scene()
    m = Material(Vec3(0.7, 0.7, 0.7), diffuse=0.8)
    scene.add_object(packer.TriangleMesh(v.reshape(-1, 3), np.arange(3 * n).reshape(-1, 3), m))
    scene.add_light_sample(Vec3(0, 12, 6))
    o, d = _rays(20000, 2)
    ids0, rec0 = renderer.trace_rays(scene, o, d, "numba", "f32", use_bvh=1, top_nodes=0)
    ids1, rec1 = renderer.trace_rays(scene, o, d, "numba", "f32", use_bvh=1, top_nodes=1024)
    assert np.array_equal(ids0, ids1) and np.array_equal(rec0, rec1)
    cam = Camera(Vec3(0, 0, 30.0), Vec3(0, 0, 0), Vec3(0, 1, 0), 40.0, 64 / 48)
    a0, c0 = renderer.B200PathTracer(precision="f32", seed=1, top_nodes=0).render_accum(scene, cam, RenderSettings(64, 48, 8, 4))
    a1, c1 = renderer.B200PathTracer(precision="f32", seed=1, top_nodes=1024).render_accum(scene, cam, RenderSettings(64, 48, 8, 4))
    assert np.array_equal(a0, a1) and np.array_equal(c0[:4], c1[:4])


@pytest.mark.parametrize("size", [(1920, 1080), (96, 54), (64, 37)])
def test_primary_candidate_masks_are_conservative(cornell, size):
    """Camera-ray candidate masks (rt_path.cuh:primary_mask_kernel) at sizes with W % 32 == 0 (masks on) and not
    (masks off): the sums equal the unmasked scan bit for bit, at the headline size too."""
    scene, b = cornell
    W, H = size
    cam = b.create_camera(W / H)
    spp = 2 if W > 1000 else 16
    a, ca = renderer.B200PathTracer(precision="f32", seed=11).render_accum(scene, cam, RenderSettings(W, H, spp, 8))
    c, cc = renderer.B200PathTracer(precision="f32", seed=11, primary_masks=False).render_accum(scene, cam, RenderSettings(W, H, spp, 8))
    assert np.array_equal(a, c) and np.array_equal(ca[:6], cc[:6])
    if W % 32 == 0:
        assert ca[10] > 0 and ca[10] < cc[0] * (3 * 42 + 33 + 3 * 28)      # fewer record tests than the full scan
