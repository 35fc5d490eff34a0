"""CPU tests of the host side: API mirror, packer, C-ABI surface, sample split, gloo reduce."""
import ctypes
import os
import random
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plugin_registry_contract():
    from b200rt import plugin
    class R(plugin.BaseRenderer):
        def __init__(self, tag="x"):
            super().__init__("dummy"); self.tag = tag
        def render(self, scene, camera, settings): return None
        def get_capabilities(self): return ["a", "b"]
    plugin.RendererFactory.register("dummy", R)
    r = plugin.RendererFactory.create("dummy", tag="y")
    assert r.get_name() == "dummy" and r.tag == "y" and r.supports("a") and not r.supports("z")
    assert "dummy" in plugin.RendererFactory.list_available()
    with pytest.raises(ValueError):
        plugin.RendererFactory.create("nope")


def test_scene_api_semantics():
    from b200rt.scene_api import AABB, HitRecord, Material, Plane, Ray, Sphere, Triangle, Vec3
    r = Ray(Vec3(0, 0, 5), Vec3(0, 0, -2))
    assert r.direction.z == -1.0
    rec = HitRecord()
    s = Sphere(Vec3(0, 0, 0), 1.0, Material())
    assert s.hit(r, 1e-3, float("inf"), rec) and rec.t == 4.0 and rec.normal.z == 1.0
    p = Plane(Vec3(-1, -1, 0), Vec3(0, 0, 1), Vec3(2, 0, 0), Vec3(0, 2, 0), 2, 2, Material())
    assert p.hit(r, 1e-3, 5.0, rec) and rec.t == 5.0          # closed range on t_max
    assert (rec.u, rec.v) == (0.5, 0.5)
    t = Triangle(Vec3(-1, -1, 0), Vec3(1, -1, 0), Vec3(0, 1, 0), np.array([0, 0]), np.array([1, 0]), np.array([0, 1]), Material())
    assert t.hit(r, 1e-3, float("inf"), rec) and rec.normal.z == 1.0
    assert not t.hit(r, 1e-3, 5.0, rec)                       # strict upper bound
    with pytest.raises(ZeroDivisionError):
        AABB(Vec3(-1, -1, -1), Vec3(1, 1, 1)).hit(r, 0, 10)
    ok, d = Vec3(0, -1, 0).refract(Vec3(0, 1, 0), 1 / 1.5)
    assert ok and abs(d.y + 1) < 1e-12


def test_cornell_builder_matches_reference_when_present(cornell):
    """Bit-for-bit agreement with the reference's CustomSceneBuilder (container only)."""
    from oracle import ref_harness as RH
    if not RH.available():
        pytest.skip("reference not mounted")
    ref_scene, ref_cam = RH.build_reference_scene(0, 16 / 9)
    scene, b = cornell
    def sig(o):
        n = type(o).__name__
        if n == "Plane":
            return (n, *o.anchor.__dict__.values()) if hasattr(o.anchor, "__dict__") else \
                (n, o.anchor.x, o.anchor.y, o.anchor.z, o.normal.x, o.normal.y, o.normal.z, o.u_unit.x, o.v_unit.z, o.u_len)
        if n == "Sphere":
            return (n, o.center.x, o.center.y, o.center.z, o.radius, o.material.refractive, o.material.ior)
        return (n, o.v0.x, o.v0.y, o.v0.z, o.v1.x, o.v1.y, o.v1.z, o.v2.x, o.v2.y, o.v2.z, o.normal.x, o.normal.y,
                o.normal.z, tuple(o.uv0), tuple(o.uv1), tuple(o.uv2), o.material.texture.path)
    def sig2(o):
        n = type(o).__name__
        if n == "Plane":
            return (n, o.anchor.x, o.anchor.y, o.anchor.z, o.normal.x, o.normal.y, o.normal.z, o.u_unit.x, o.v_unit.z, o.u_len)
        return sig(o)
    assert [sig2(o) for o in scene.objects] == [sig2(o) for o in ref_scene.objects]
    assert [(l.x, l.y, l.z) for l in scene.lights] == [(l.x, l.y, l.z) for l in ref_scene.lights]
    cam = b.create_camera(16 / 9)
    for f in ("origin", "lower_left_corner", "horizontal", "vertical"):
        assert getattr(cam, f).x == getattr(ref_cam, f).x and getattr(cam, f).z == getattr(ref_cam, f).z


def test_packer_layout(cornell, golden_dir):
    """SoA streams vs the reference's AoS block (golden): same values, derived fields as Numba rounds them."""
    from b200rt import packer
    scene, b = cornell
    p = packer.pack_scene(scene, "numba")
    sd = np.load(f"{golden_dir}/packed_scene_seed0.npz")["scene"]
    nP = int(sd[0]); pl = sd[1:1 + 20 * nP].reshape(nP, 20); off = 1 + 20 * nP
    nS = int(sd[off]); sp = sd[off + 1: off + 1 + 12 * nS].reshape(nS, 12); off += 1 + 12 * nS
    nT = int(sd[off]); tr = sd[off + 1:].reshape(nT, 26)
    assert (p.n_rect, p.n_sphere, p.n_tri) == (nP, nS, nT) == (5, 3, 26)
    R = p.rect.reshape(-1, 4, 4)
    assert np.array_equal(R[:, 0, :3], pl[:, 0:3]) and np.array_equal(R[:, 1, :3], pl[:, 3:6])
    assert np.array_equal(R[:, 0, 3], pl[:, 12]) and np.array_equal(R[:, 1, 3], pl[:, 13])
    un = np.sqrt((pl[:, 6:9] ** 2).sum(1, dtype=np.float32), dtype=np.float32)
    assert np.array_equal(R[:, 2, :3], (pl[:, 6:9] / un[:, None]).astype(np.float64))
    S = p.sphere.reshape(-1, 2, 4)
    assert np.array_equal(S[:, 0], sp[:, 0:4]) and np.array_equal(S[:, 1, 0], (sp[:, 3] * sp[:, 3]))
    T = p.tri.reshape(-1, 3, 4)
    assert np.array_equal(T[:, 0, :3], tr[:, 0:3])
    assert np.array_equal(T[:, 1, :3], (tr[:, 3:6] - tr[:, 0:3]).astype(np.float64))
    assert np.array_equal(T[:, 2, :3], (tr[:, 6:9] - tr[:, 0:3]).astype(np.float64))
    SH = p.shade.reshape(-1, 3, 4)[nP + nS:]
    assert np.array_equal(SH[:, 0, :3], tr[:, 9:12]) and np.array_equal(SH[:, 1], tr[:, 20:24])
    # materials: planes/triangles never refractive, only triangles textured (numba semantics)
    M = p.mat.reshape(-1, 2, 4)
    tri_m = p.prim_mat[nP + nS:]
    assert np.array_equal(M[tri_m, 0, :3], tr[:, 12:15]) and np.array_equal(p.mat_tex[tri_m], tr[:, 19].astype(np.int32))
    assert (M[p.prim_mat[:nP], 1, 2] == 0).all() and (p.mat_tex[p.prim_mat[:nP + nS]] == -1).all()
    assert np.array_equal(M[p.prim_mat[nP:nP + nS], 1, 2:4], sp[:, 10:12])
    # textures: RGBX8, ids by sorted path
    texels, info, ids = packer.pack_textures(scene)
    assert info[:, 1:3].tolist() == [[1318, 1319], [1317, 1316], [2978, 2393], [1269, 1268], [1315, 1314], [1296, 1296], [1320, 1320]]
    assert texels.size == 52070958 // 3 and (texels >> 24 == 255).all()
    t0 = scene.objects[[type(o).__name__ for o in scene.objects].index("Triangle")].material.texture
    k = ids[t0.path]
    px = texels[info[k, 0] + 5 * info[k, 1] + 7]
    assert (px & 255, (px >> 8) & 255, (px >> 16) & 255) == tuple(int(v) for v in t0.pixels[5, 7])
    # cpu semantics keep un-rounded doubles and full materials
    pc = packer.pack_scene(scene, "cpu")
    assert pc.semantics == 1 and (pc.sphere.reshape(-1, 2, 4)[:, 1, 0] == 9.0).all()
    assert pc.order[:, 0].tolist() == p.order[:, 0].tolist()


def test_triangle_mesh_side_door():
    from b200rt import packer
    from b200rt.scene_api import Material, Scene, Triangle, Vec3
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.0]])
    f = np.array([[0, 1, 2], [0, 2, 3]])
    a, b_ = Scene(), Scene()
    m = Material(Vec3(0.5, 0.6, 0.7), diffuse=0.8)
    a.objects.append(packer.TriangleMesh(v, f, m))
    for tri in f:
        b_.objects.append(Triangle(*(Vec3(*v[i]) for i in tri), None, None, None, m))
    pa, pb = packer.pack_scene(a, "numba"), packer.pack_scene(b_, "numba")
    assert np.array_equal(pa.tri, pb.tri) and np.array_equal(pa.shade, pb.shade)
    assert pa.order.tolist() == [[0, 0], [0, 1]]


def test_c_abi_exports_every_declared_symbol():
    """libb200rt.so loads (no GPU needed) and exports exactly what include/b200rt.h declares."""
    from b200rt import _lib, build
    path = build.build()
    lib = ctypes.CDLL(path)
    hdr = open(os.path.join(ROOT, "include", "b200rt.h")).read()
    declared = set(re.findall(r"\b(b2rt_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert ctypes.sizeof(_lib.SceneStruct) == 8 * 4 + 12 * 8 + 2 * 4 + 0 or ctypes.sizeof(_lib.SceneStruct) % 8 == 0
    lib.b2rt_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.b2rt_last_error(), bytes)
    nbytes = ctypes.c_size_t(0)
    assert lib.b2rt_path_workspace_bytes(0, 1920, 1080, 8, 8, ctypes.byref(nbytes)) == 0
    assert nbytes.value >= 1920 * 1080 * 8 * 176
    assert lib.b2rt_path_workspace_bytes(0, 0, 0, 0, 0, ctypes.byref(nbytes)) != 0       # error path, no GPU call


def test_renderer_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from b200rt import renderer
    from b200rt.plugin import RendererFactory
    assert {"b200_path_tracer", "b200_texture_raytracer", "b200_raytracer"} <= set(RendererFactory.list_available())
    with pytest.raises(RuntimeError):
        RendererFactory.create("b200_path_tracer")


def test_split_samples_is_exhaustive_and_contiguous():
    from b200rt.dist import split_samples
    for spp in (1, 7, 8, 1024, 4096):
        for world in (1, 2, 3, 4, 8):
            parts = [split_samples(spp, r, world) for r in range(world)]
            assert sum(p[0] for p in parts) == spp
            off = 0
            for n, o in parts:
                assert o == off
                off += n
            assert max(p[0] for p in parts) - min(p[0] for p in parts) <= 1


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as td
from b200rt import dist
td.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=int(sys.argv[3]), world_size=2)
rank, world = dist.rank_world()
n, off = dist.split_samples(9, rank, world)
buf = torch.full((16,), float(n), dtype=torch.float32) + off
dist.reduce_to_root(buf)
if rank == 0:
    assert torch.allclose(buf, torch.full((16,), 9.0 + 5.0)), buf
    print("root ok")
td.destroy_process_group()
"""


def test_gloo_world2_reduce(tmp_path):
    """N > 1 host path on CPU: two processes, spp split + one SUM reduce onto rank 0."""
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    port = str(29500 + random.randint(0, 400))
    pkg = os.path.join(ROOT, "path-tracing__ray-tracer_b200")
    procs = [subprocess.Popen([sys.executable, str(script), pkg, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "root ok" in outs[0]


def test_obj_loader(tmp_path):
    from b200rt import packer
    from b200rt.scene_api import Material, Scene, Vec3
    p = tmp_path / "quad.obj"
    p.write_text("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nf 1/1 2/2 3/3 4/4\n")
    mesh = packer.load_obj(str(p), Material(Vec3(1, 1, 1)), scale=2.0, translate=(1, 0, 0))
    assert mesh.faces.shape == (2, 3) and mesh.uvs.shape == (6, 2)
    sc = Scene(); sc.objects.append(mesh)
    pk = packer.pack_scene(sc, "numba")
    assert pk.n_tri == 2
    assert np.allclose(pk.tri.reshape(-1, 3, 4)[0, 0, :3], [1, 0, 0]) and np.allclose(pk.tri.reshape(-1, 3, 4)[0, 1, :3], [2, 0, 0])
    rec = packer.build_scan_prims(pk).reshape(-1, 4, 4)
    assert rec.shape[0] == 1 and (rec[0, 3, 2].view(np.int32) >> 28) == 2          # the two triangles pair into one parallelogram


def test_cli_parser_matches_reference_flags():
    import b200rt.cli as cli
    src = open(cli.__file__).read()
    for flag in ("--renderer", "--scene", "--width", "--height", "--samples", "--depth", "--output", "--path-samples"):
        assert flag in src


def _np_planar(rec, o, d, t_min=1e-3, t_max=1e6):
    """numpy restatement of the planar scan-record test -> (t, record index)"""
    best_t = np.full(o.shape[0], t_max); best_k = np.full(o.shape[0], -1)
    kinds = rec[:, 3, 2].astype(np.float32).view(np.int32) >> 28
    for k in range(rec.shape[0]):
        q0, q1, q2, q3 = rec[k].astype(np.float64)
        dn = d @ q0[:3]
        with np.errstate(divide="ignore", invalid="ignore"):
            t = (q0[3] - o @ q0[:3]) / dn
            X = o + t[:, None] * d
            u, v = X @ q1[:3] + q1[3], X @ q2[:3] + q2[3]
            inside = (u >= 0) & (v >= 0) & ((u + v <= 1) if kinds[k] == 1 else ((u <= q3[0]) & (v <= q3[1])))
            ok = inside & (np.abs(dn) > 1e-6) & (t > t_min) & (t < best_t)
        best_t[ok], best_k[ok] = t[ok], k
    return best_t, best_k


def _np_boxes(boxes, o, d, t_min=1e-3, t_max=1e6):
    """numpy restatement of csrc/rt_scene.cuh:scan_box -> (t, planar record index of the face)"""
    best_t = np.full(o.shape[0], t_max); best_k = np.full(o.shape[0], -1)
    for b in range(boxes.shape[0]):
        M, off = boxes[b, :3, :3].astype(np.float64), boxes[b, :3, 3].astype(np.float64)
        w = boxes[b, 3, :2].view(np.uint32)
        code = np.array([(int(w[0]) >> (8 * i)) & 255 for i in range(4)] + [int(w[1]) & 255, (int(w[1]) >> 8) & 255])
        lo, ld = o @ M.T + off, d @ M.T
        with np.errstate(divide="ignore", invalid="ignore"):
            ta, tb = (-1 - lo) / ld, (1 - lo) / ld
        tn, tf = np.fmin(ta, tb), np.fmax(ta, tb)
        ke, kx = np.argmax(tn, axis=1), np.argmin(tf, axis=1)
        rows = np.arange(o.shape[0])
        te, tx = tn[rows, ke], tf[rows, kx]
        fe = 2 * ke + (ld[rows, ke] < 0)
        fx = 2 * kx + (ld[rows, kx] > 0)
        ce, cx = code[fe], code[fx]
        use_e = (te > t_min) & (ce != 255)
        t = np.where(use_e, te, tx); c = np.where(use_e, ce, cx)
        ok = (te <= tx) & (t > t_min) & (t < best_t) & (c != 255)
        best_t[ok], best_k[ok] = t[ok], c[ok]
    return best_t, best_k


def test_box_records_group_cornell_and_match_planar_scan(cornell):
    """packer.group_scan_boxes: the Cornell walls (5 faces) and the two cubes (6 faces each) become three box
    records, the canvas stays loose; the three-slab box test finds the same face and distance as the planar
    records it replaces (numpy restatement of both device tests, float64)."""
    from b200rt import packer
    pk = packer.pack_scene(cornell[0], "numba")
    quads = []
    rec0 = packer.build_scan_prims(pk, quads_out=quads)
    rec, n_loose, boxes = packer.group_scan_boxes(rec0, quads)
    rec, boxes = rec.reshape(-1, 4, 4), boxes.reshape(-1, 4, 4)
    assert rec.shape[0] == 18 and n_loose == 1 and boxes.shape[0] == 3
    faces = sorted(int(c) for b in boxes for c in b[3, :2].view(np.uint8)[:6] if c != 255)
    assert faces == list(range(1, 18))                     # every non-loose record is a face of exactly one box
    assert sorted(map(bytes, rec)) == sorted(map(bytes, rec0.reshape(-1, 4, 4)))   # a permutation, nothing altered
    rng = np.random.default_rng(5)
    n = 100000
    o = rng.uniform(-14.9, 14.9, size=(n, 3))
    o[: n // 4] = np.array([0.0, 0.0, 40.0]) + rng.normal(size=(n // 4, 3))        # from outside, like the camera
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[: n // 4, 2] = -np.abs(d[: n // 4, 2])
    tp, kp = _np_planar(rec[n_loose:], o, d)
    tb, kb = _np_boxes(boxes, o, d)
    kp = np.where(kp >= 0, kp + n_loose, -1)
    assert ((kp >= 0) == (kb >= 0)).mean() > 0.9999
    both = (kp >= 0) & (kb >= 0)
    assert np.abs(tp[both] - tb[both]).max() < 1e-4
    # same distance => same point; the label differs only on shared edges and on the coplanar overlapping faces
    # (cube 1 top / cube 2 bottom at y = -9.4, cube 1 bottom / floor), reachable only from inside a cube
    assert (kp[both] == kb[both]).mean() > 0.995
    outside = np.arange(n) < n // 4
    assert (kp[both & outside] == kb[both & outside]).mean() > 0.9995
    assert (kp >= 0).mean() > 0.6


def test_box_grouping_leaves_unrelated_quads_alone():
    from b200rt import packer
    from b200rt.scene_api import Material, Plane, Scene, Vec3
    sc = Scene()
    m = Material(Vec3(1, 1, 1))
    sc.objects.append(Plane(Vec3(0, 0, 0), Vec3(0, 1, 0), Vec3(1, 0, 0), Vec3(0, 0, 1), 2.0, 2.0, m))
    sc.objects.append(Plane(Vec3(0, 3, 0), Vec3(0, 1, 0), Vec3(1, 0, 0), Vec3(0, 0, 1), 2.0, 1.0, m))   # other size
    pk = packer.pack_scene(sc, "numba")
    quads = []
    rec0 = packer.build_scan_prims(pk, quads_out=quads)
    rec, n_loose, boxes = packer.group_scan_boxes(rec0, quads)
    assert n_loose == 2 and boxes.shape[0] == 0 and np.array_equal(rec, rec0)


def test_surface_records_and_bounds(cornell):
    """packer.build_surface_records: one 5-float4 record per primitive restating what cuda_scene_hit returns beside t
    (normal / sphere centre, material, affine uv map); PackedScene.bounds contains every primitive."""
    from b200rt import packer
    pk = packer.pack_scene(cornell[0], "numba")
    rec = packer.build_surface_records(pk).reshape(-1, 5, 4)
    assert rec.shape[0] == pk.n_prims == 34
    R, S, SH = pk.rect.reshape(-1, 4, 4), pk.sphere.reshape(-1, 2, 4), pk.shade.reshape(-1, 3, 4)
    M = pk.mat.reshape(-1, 2, 4)
    flags = rec[:, 4, 3].view(np.int32)
    tex = rec[:, 3, 3].view(np.int32)
    for k in range(pk.n_prims):
        m = pk.prim_mat[k]
        assert np.array_equal(rec[k, 1], M[m, 0].astype(np.float32)) and np.array_equal(rec[k, 2], M[m, 1].astype(np.float32))
        assert tex[k] == pk.mat_tex[m]
    for i in range(pk.n_rect):                       # u = a / u_len, v = b / v_len, never flipped
        assert np.array_equal(rec[i, 0, :3], R[i, 1, :3].astype(np.float32)) and rec[i, 0, 3] == 0 and flags[i] == 0
        a, b = 0.3 * R[i, 0, 3], 0.8 * R[i, 1, 3]
        u = rec[i, 3, 0] + a * rec[i, 3, 1] + b * rec[i, 3, 2]
        v = rec[i, 4, 0] + a * rec[i, 4, 1] + b * rec[i, 4, 2]
        assert abs(u - 0.3) < 1e-6 and abs(v - 0.8) < 1e-6
    for i in range(pk.n_sphere):                     # centre + 1/r: n = (p - c) / r
        k = pk.n_rect + i
        assert np.array_equal(rec[k, 0, :3], S[i, 0, :3].astype(np.float32)) and abs(rec[k, 0, 3] * S[i, 0, 3] - 1) < 1e-6
    base = pk.n_rect + pk.n_sphere
    for i in range(pk.n_tri):                        # w uv0 + a uv1 + b uv2 (cuda_path_tracer.py:722-724), flipped to face the ray
        k = base + i
        assert flags[k] == 1
        uv0, uv1, uv2 = SH[k, 1, 0:2], SH[k, 1, 2:4], SH[k, 2, 0:2]
        a, b = 0.25, 0.6
        want = (1 - a - b) * uv0 + a * uv1 + b * uv2
        got = np.array([rec[k, 3, 0] + a * rec[k, 3, 1] + b * rec[k, 3, 2], rec[k, 4, 0] + a * rec[k, 4, 1] + b * rec[k, 4, 2]])
        assert np.abs(got - want).max() < 1e-6
    lo, hi = pk.bounds()
    assert (lo < -15).all() and (lo > -15.1).all() and (hi[:2] > 15).all() and (hi < 15.1).all()
    from b200rt.scene_api import Scene
    elo, ehi = packer.pack_scene(Scene(), "numba").bounds()
    assert (elo > ehi).all()                         # empty scene: "unknown"


def test_scene_add_mesh_and_add_obj(tmp_path):
    """Bulk mesh ingestion through the Scene mirror: add_mesh / add_obj put one TriangleMesh into scene.objects and the
    packer expands it; build_bvh() leaves such scenes to the device builder."""
    from b200rt import packer
    from b200rt.scene_api import Material, Plane, Scene, Vec3
    sc = Scene()
    m = Material(Vec3(0.5, 0.6, 0.7), diffuse=0.8)
    sc.add_object(Plane(Vec3(-1, 0, -1), Vec3(0, 1, 0), Vec3(1, 0, 0), Vec3(0, 0, 1), 2.0, 2.0, m))
    v = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], dtype=float)
    mesh = sc.add_mesh(v, [[0, 1, 2], [0, 2, 3]], m, uvs=[[0, 0], [1, 0], [1, 1], [0, 1]])
    obj = tmp_path / "tri.obj"
    obj.write_text("v 0 0 1\nv 1 0 1\nv 0 1 1\nf 1 2 3\n")
    sc.add_obj(str(obj), m, scale=2.0)
    sc.build_bvh()
    assert sc.bvh_root is None and len(sc.objects) == 3 and sc.objects[1] is mesh
    pk = packer.pack_scene(sc, "numba")
    assert (pk.n_rect, pk.n_sphere, pk.n_tri) == (1, 0, 3)
    T = pk.tri.reshape(-1, 3, 4)
    assert np.allclose(T[2, 0, :3], [0, 0, 2]) and np.allclose(T[2, 1, :3], [2, 0, 0])       # scaled OBJ triangle
    assert pk.order[1].tolist() == [1, 0] and pk.order[3].tolist() == [2, 0]                  # (object index, face index)


# ------------------------------------------------------------------------------------ library-side scene preparation
def _prepare_both(packed):
    """(numpy records, library records) for one packed scene: (planar, n_loose, boxes, surface, hints, bounds)."""
    import ctypes as C
    from b200rt import _lib, device
    from b200rt.packer import build_surface_records
    rec, hints, n_loose, boxes = device._small_scene_records(packed, True, True)
    ref = (rec, n_loose, boxes, build_surface_records(packed), hints, packed.bounds())
    lay, host = device._library_records(_lib.load(), packed, True, True, True)
    nrec = lay.n_scan_prims + lay.n_scan_boxes
    allrec = np.frombuffer(host, dtype=np.float32, count=16 * nrec, offset=lay.scan_offset).reshape(-1, 4)
    got = (allrec[:4 * lay.n_scan_prims], int(lay.n_scan_loose), allrec[4 * lay.n_scan_prims:],
           np.frombuffer(host, dtype=np.float32, count=20 * packed.n_prims, offset=lay.surface_offset).reshape(-1, 4),
           np.frombuffer(host, dtype=np.int32, count=packed.lights.shape[0], offset=lay.hint_offset),
           (np.array(lay.bounds_lo[:]), np.array(lay.bounds_hi[:])))
    return ref, got


def _assert_records_equal(a, b, bit_cols):
    """float columns to float32 rounding, bit-packed columns exactly."""
    a, b = np.asarray(a, np.float32).reshape(-1, 4, 4), np.asarray(b, np.float32).reshape(-1, 4, 4)
    assert a.shape == b.shape
    ai, bi = a.view(np.int32), b.view(np.int32)
    mask = np.zeros((4, 4), bool)
    for r, c in bit_cols:
        mask[r, c] = True
    assert np.array_equal(ai[:, mask], bi[:, mask])
    assert np.allclose(a[:, ~mask], b[:, ~mask], rtol=2e-6, atol=1e-6)


def test_library_scene_prepare_matches_the_numpy_packer(cornell):
    """b2rt_scene_prepare_host (host C++ inside libb200rt.so: what a C-ABI binder gets) derives the same planar / box /
    surface records, bounds and occluder hints as the independent numpy implementation in b200rt.packer."""
    from b200rt import packer
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import random_scenes as RS
    scenes = [cornell[0]] + [RS.make_scene(RS.local_api(), s)[0] for s in (11, 12, 13)]
    for i, scene in enumerate(scenes):
        pk = packer.pack_scene(scene, "numba")
        ref, got = _prepare_both(pk)
        _assert_records_equal(ref[0], got[0], [(3, 2), (3, 3)])                      # planar: kind | idA, idB
        assert ref[1] == got[1]
        _assert_records_equal(ref[2], got[2], [(3, 0), (3, 1), (3, 2), (3, 3)])      # boxes: face codes, flags
        s_ref, s_got = np.asarray(ref[3], np.float32).reshape(-1, 5, 4), got[3].reshape(-1, 5, 4)
        assert np.array_equal(s_ref.view(np.int32)[:, 3:, 3], s_got.view(np.int32)[:, 3:, 3])      # texture id, flags
        assert np.allclose(s_ref[:, :3], s_got[:, :3], rtol=1e-6, atol=1e-7) and np.allclose(s_ref[:, 3:, :3], s_got[:, 3:, :3], rtol=1e-6, atol=1e-7)
        assert np.allclose(ref[5][0], got[5][0], rtol=1e-6) and np.allclose(ref[5][1], got[5][1], rtol=1e-6)
        if i == 0:
            # Cornell: three boxes (walls, two cubes), the canvas loose, every light sample shadowed by the ceiling record
            assert got[2].shape[0] // 4 == 3 and got[1] == 1
            assert np.array_equal(ref[4], got[4]) and len(set(got[4].tolist())) == 1


def test_scene_struct_carries_size_and_abi_version():
    """b2rt_scene starts with struct_size / abi_version; the ctypes mirror has the size the header's struct has."""
    import ctypes as C
    import re
    from b200rt import _lib
    s = _lib.new_scene_struct()
    assert s.struct_size == C.sizeof(_lib.SceneStruct) and s.abi_version == _lib.ABI_VERSION
    hdr = open(os.path.join(ROOT, "include", "b200rt.h")).read()
    assert int(re.search(r"#define B2RT_ABI_VERSION (\d+)", hdr).group(1)) == _lib.ABI_VERSION
    # compile the header with the host compiler and compare sizeof / the offset of the last field
    import subprocess, tempfile
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write('#include <stdio.h>\n#include <stddef.h>\n#include "b200rt.h"\n'
                             'int main(void){printf("%zu %zu %zu %zu %zu", sizeof(b2rt_scene), offsetof(b2rt_scene, bounds_hi), sizeof(b2rt_prepare_layout), offsetof(b2rt_scene, d_bvh_wide), offsetof(b2rt_scene, d_bvh_quant));return 0;}')
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", os.path.join(d, "t")], check=True)
        size, off, lay, off_wide, off_quant = (int(x) for x in subprocess.run([os.path.join(d, "t")], capture_output=True, text=True).stdout.split())
    assert size == C.sizeof(_lib.SceneStruct) and off == _lib.SceneStruct.bounds_hi.offset
    assert off_wide == _lib.SceneStruct.d_bvh_wide.offset and off_quant == _lib.SceneStruct.d_bvh_quant.offset
    assert lay == C.sizeof(_lib.PrepareLayout)


def test_morton_face_order_is_a_locality_improving_permutation_with_the_same_hits():
    """TriangleMesh.spatially_sorted(): a permutation of the faces (adaptive-bit Morton curve) that keeps every closest
    hit — same distance, the same face through the permutation — and makes consecutive faces spatial neighbours."""
    from b200rt import packer, scenes
    from b200rt.scene_api import Scene
    from oracle import cpu_oracle as O
    mesh = scenes.heightfield_mesh(41, 21, seed=5)                   # 1 600 triangles, listed as two half-meshes
    order = packer.morton_face_order(mesh.vertices, mesh.faces)
    assert sorted(order.tolist()) == list(range(mesh.faces.shape[0]))
    sorted_mesh = mesh.spatially_sorted()
    assert np.array_equal(sorted_mesh.faces, mesh.faces[order]) and np.array_equal(sorted_mesh.vertices, mesh.vertices)

    def block_extent(m, k=16):                                       # mean xz bounding-box diagonal of k consecutive faces
        c = m.vertices[m.faces].mean(1)[:, [0, 2]]
        c = c[: len(c) // k * k].reshape(-1, k, 2)
        return float(np.linalg.norm(c.max(1) - c.min(1), axis=1).mean())
    assert block_extent(sorted_mesh) < 0.5 * block_extent(mesh)
    # closest hits through the oracle's brute-force scan, float64
    rng = np.random.default_rng(1)
    n = 2000
    o = np.stack([rng.uniform(-14, 14, n), np.full(n, 5.0), rng.uniform(-14, 14, n)], 1)
    d = np.stack([rng.normal(0, 0.3, n), -np.ones(n), rng.normal(0, 0.3, n)], 1)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    hits = []
    from b200rt.cornell import CustomSceneBuilder
    cam = CustomSceneBuilder(texture_dir=False).create_camera(1.0)
    for m in (mesh, sorted_mesh):
        sc = Scene()
        sc.objects.append(m)
        ids, rec = O.nb_scene_hit_rays(O.nb_pack(sc, cam), o, d)
        hits.append((ids, rec[:, 0]))
    (ia, ta), (ib, tb) = hits
    assert np.array_equal(ta, tb) and (ia >= 0).mean() > 0.5
    hit = ia >= 0
    assert np.array_equal(hit, ib >= 0)
    same_face = order[ib[hit]] == ia[hit]
    assert same_face.mean() > 0.999                                  # anything else is an exact tie on a shared edge
    sc = Scene()
    assert np.array_equal(sc.add_mesh(mesh.vertices, mesh.faces, mesh.material, spatial_order=True).faces, sorted_mesh.faces)
    # degenerate input: all centroids equal -> identity; empty mesh -> empty permutation
    assert np.array_equal(packer.morton_face_order(np.zeros((3, 3)), np.array([[0, 1, 2], [0, 1, 2]])), [0, 1])
    assert packer.morton_face_order(np.zeros((0, 3)), np.zeros((0, 3), np.int64)).shape == (0,)
