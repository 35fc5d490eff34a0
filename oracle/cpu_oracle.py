"""ctypes front-end of the CPU ORACLE (``oracle/rt_oracle.c``) — test infrastructure only.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this.  The product (``path-tracing__ray-tracer_b200/``) never does.

Scene export here restates the reference's host packers so the oracle consumes the *same bytes*
the reference kernels would:
  * ``nb_pack``   <- ``CUDAPathTracer._prepare_scene_data/_camera_data/_light_data/_texture_data``
                     (``renderers/cuda_path_tracer.py:819-946``; identical in cuda_texture_renderer.py)
  * ``cpu_export``<- the object graph the CPU renderer walks (``core/geometry.py``, ``core/acceleration.py``)
Works on the reference's own objects and on the ``b200rt.scene_api`` mirror alike (duck typing).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "librt_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "rt_oracle.c")
    if force or not os.path.isfile(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["sh", os.path.join(_HERE, "build.sh")], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_nb_xorshift.restype = C.c_int64
        _lib.orc_nb_xorshift.argtypes = [C.c_int64]
        _lib.orc_nb_random.restype = C.c_double
        _lib.orc_nb_random.argtypes = [C.c_int64]
        _lib.orc_nb_tonemap.restype = C.c_double
        _lib.orc_nb_tonemap.argtypes = [C.c_double]
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a, ty):
    return a.ctypes.data_as(C.POINTER(ty)) if a is not None else None


# ------------------------------------------------------------------------------ numba family
def _kind(obj) -> str:
    if hasattr(obj, "vertices") and hasattr(obj, "faces"):
        return "mesh"
    if hasattr(obj, "anchor"):
        return "plane"
    if hasattr(obj, "radius"):
        return "sphere"
    if hasattr(obj, "v0"):
        return "triangle"
    raise TypeError(f"unsupported object {type(obj).__name__}")


def texture_order(scene):
    """Distinct texture path strings, sorted — the reference's texture-id rule (:824-832)."""
    paths = []
    for o in scene.objects:
        tex = getattr(getattr(o, "material", None), "texture", None)
        if tex is not None:
            p = getattr(tex, "path", None)
            if p and p not in paths:
                paths.append(p)
    return sorted(paths)


@dataclass
class NbPacked:
    scene: np.ndarray      # float32 [1+20P+1+12S+1+26T]
    camera: np.ndarray     # float32 [12]
    lights: np.ndarray     # float32 [1+3L]
    tex: np.ndarray        # uint8 flat RGB
    tex_info: np.ndarray   # int32 [3*ntex]: offset, w, h
    order: np.ndarray      # int32 [n_prims]: packed index -> index in scene.objects


def nb_pack(scene, camera, with_textures: bool = True) -> NbPacked:
    tex_paths = texture_order(scene)
    tex_id = {p: i for i, p in enumerate(tex_paths)}
    planes, spheres, tris = [], [], []
    o_pl, o_sp, o_tr = [], [], []
    mesh_blocks, mesh_order, n_mesh_tris = [], [], 0       # meshes go behind the individual triangles
    for idx, o in enumerate(scene.objects):
        k, m = _kind(o), o.material
        if k == "plane":
            planes += [o.anchor.x, o.anchor.y, o.anchor.z, o.normal.x, o.normal.y, o.normal.z,
                       o.u_dir.x, o.u_dir.y, o.u_dir.z, o.v_dir.x, o.v_dir.y, o.v_dir.z, o.u_len, o.v_len,
                       m.color.x, m.color.y, m.color.z, m.diffuse, m.specular, m.reflective]
            o_pl.append(idx)
        elif k == "sphere":
            spheres += [o.center.x, o.center.y, o.center.z, o.radius, m.color.x, m.color.y, m.color.z,
                        m.diffuse, m.specular, m.reflective, getattr(m, "refractive", 0.0), getattr(m, "ior", 1.0)]
            o_sp.append(idx)
        elif k == "mesh":
            # bulk triangles (b200rt.packer.TriangleMesh): the same 26 floats per triangle as _prepare_scene_data
            # (:861-885), geometric normal as Triangle.__init__ computes it (core/geometry.py:130), default uvs (:869-874)
            V = np.asarray(o.vertices, dtype=np.float64); F = np.asarray(o.faces, dtype=np.int64)
            v0, v1, v2 = V[F[:, 0]], V[F[:, 1]], V[F[:, 2]]
            nn = np.cross(v1 - v0, v2 - v0)
            ln = np.linalg.norm(nn, axis=1, keepdims=True)
            nn = np.divide(nn, ln, out=np.zeros_like(nn), where=ln > 0)
            blk = np.zeros((F.shape[0], 26))
            blk[:, 0:3], blk[:, 3:6], blk[:, 6:9], blk[:, 9:12] = v0, v1, v2, nn
            blk[:, 12:18] = (m.color.x, m.color.y, m.color.z, m.diffuse, m.specular, m.reflective)
            blk[:, 18], blk[:, 19] = 0.0, -1.0
            if getattr(o, "uvs", None) is not None:
                U = np.asarray(o.uvs, dtype=np.float64)
                blk[:, 20:22], blk[:, 22:24], blk[:, 24:26] = U[F[:, 0]], U[F[:, 1]], U[F[:, 2]]
            else:
                blk[:, 20:26] = (0.0, 0.0, 1.0, 0.0, 1.0, 1.0)
            mesh_blocks.append((len(tris) // 26, blk.astype(np.float32)))
            n_mesh_tris += F.shape[0]
            mesh_order.append(np.full(F.shape[0], idx, dtype=np.int32))
        else:
            has = 1.0 if m.texture is not None else 0.0
            tid = -1.0
            if m.texture is not None and getattr(m.texture, "path", None) in tex_id:
                tid = float(tex_id[m.texture.path])
            uv = [(o.uv0, (0.0, 0.0)), (o.uv1, (1.0, 0.0)), (o.uv2, (1.0, 1.0))]
            uvs = [c for a, dflt in uv for c in ((a[0], a[1]) if a is not None else dflt)]
            tris += [o.v0.x, o.v0.y, o.v0.z, o.v1.x, o.v1.y, o.v1.z, o.v2.x, o.v2.y, o.v2.z,
                     o.normal.x, o.normal.y, o.normal.z, m.color.x, m.color.y, m.color.z,
                     m.diffuse, m.specular, m.reflective, has, tid] + uvs
            o_tr.append(idx)
    data = [len(planes) // 20] + planes + [len(spheres) // 12] + spheres + [len(tris) // 26 + n_mesh_tris] + tris
    cam = [camera.origin.x, camera.origin.y, camera.origin.z,
           camera.lower_left_corner.x, camera.lower_left_corner.y, camera.lower_left_corner.z,
           camera.horizontal.x, camera.horizontal.y, camera.horizontal.z,
           camera.vertical.x, camera.vertical.y, camera.vertical.z]
    lights = [len(scene.lights)] + [c for l in scene.lights for c in (l.x, l.y, l.z)]

    by_path = {}
    for o in scene.objects:
        t = getattr(getattr(o, "material", None), "texture", None)
        if t is not None:
            by_path.setdefault(t.path, t)
    chunks, info, off = [], [], 0
    for p in tex_paths:
        px = np.ascontiguousarray(by_path[p].pixels, dtype=np.uint8).reshape(-1)
        h, w = by_path[p].pixels.shape[:2]
        info += [off, w, h]
        off += px.size
        if with_textures:
            chunks.append(px)
    tex = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.uint8)
    scene_arr = np.array(data, dtype=np.float32)
    order = np.array(o_pl + o_sp + o_tr, dtype=np.int32)
    if mesh_blocks:
        scene_arr = np.concatenate([scene_arr] + [b.reshape(-1) for _, b in mesh_blocks])
        order = np.concatenate([order] + mesh_order)
    return NbPacked(scene_arr, np.array(cam, dtype=np.float32),
                    np.array(lights, dtype=np.float32), tex, np.array(info, dtype=np.int32), order)


def _nb_args(pk: NbPacked):
    return (_p(pk.scene, C.c_float), _p(pk.camera, C.c_float), _p(pk.lights, C.c_float), C.c_int(pk.lights.size),
            _p(pk.tex, C.c_uint8), C.c_long(pk.tex.size), _p(pk.tex_info, C.c_int32), C.c_int(pk.tex_info.size))


def nb_whitted_texture(pk: NbPacked, width, height, spp, max_depth, want_float=True):
    """-> (uint8 [H,W,3] device row order (row 0 = bottom), float64 [H,W,3] or None, scene_hit calls)."""
    out = np.zeros((height, width, 3), dtype=np.uint8)
    outf = np.zeros((height, width, 3)) if want_float else None
    calls = C.c_uint64(0)
    lib().orc_nb_whitted_texture(*_nb_args(pk), C.c_int(width), C.c_int(height), C.c_int(spp), C.c_int(max_depth),
                                 _p(out, C.c_uint8), _p(outf, C.c_double), C.byref(calls))
    return out, outf, calls.value


def nb_path_trace(pk: NbPacked, width, height, spp, max_depth, frame_count=0, want_stats=True):
    """-> dict(u8 [H,W,3] device row order, sum, sumsq [H,W,3] float64, counters[4])."""
    out = np.zeros((height, width, 3), dtype=np.uint8)
    s1 = np.zeros((height, width, 3)) if want_stats else None
    s2 = np.zeros((height, width, 3)) if want_stats else None
    cnt = np.zeros(4, dtype=np.uint64)
    lib().orc_nb_path_trace(*_nb_args(pk), C.c_int(width), C.c_int(height), C.c_int(spp), C.c_int(max_depth),
                            C.c_long(frame_count), _p(out, C.c_uint8), _p(s1, C.c_double), _p(s2, C.c_double),
                            _p(cnt, C.c_uint64))
    return dict(u8=out, sum=s1, sumsq=s2,
                counters=dict(closest_rays=int(cnt[0]), shadow_rays=int(cnt[1]), segments=int(cnt[2]),
                              nee_unshadowed=int(cnt[3])))


def nb_trace_path_one(pk: NbPacked, o, d, max_depth, rng):
    o = np.ascontiguousarray(o, dtype=np.float64); d = np.ascontiguousarray(d, dtype=np.float64)
    rgb = np.zeros(3)
    lib().orc_nb_trace_path_one(*_nb_args(pk), _p(o, C.c_double), _p(d, C.c_double), C.c_int(max_depth),
                                C.c_int64(rng), _p(rgb, C.c_double))
    return rgb


def nb_scene_hit_rays(pk: NbPacked, origins, dirs, t_min=0.001, t_max=1000000.0):
    o = np.ascontiguousarray(origins, dtype=np.float64); d = np.ascontiguousarray(dirs, dtype=np.float64)
    n = o.shape[0]
    ids = np.zeros(n, dtype=np.int32); rec = np.zeros((n, 19))
    lib().orc_nb_scene_hit_rays(_p(pk.scene, C.c_float), C.c_int(n), _p(o, C.c_double), _p(d, C.c_double),
                                C.c_double(t_min), C.c_double(t_max), _p(ids, C.c_int32), _p(rec, C.c_double))
    return ids, rec


def nb_primary_hits(pk: NbPacked, width, height, du=0.5, dv=0.5):
    ids = np.zeros((height, width), dtype=np.int32); t = np.zeros((height, width))
    lib().orc_nb_primary_hits(_p(pk.scene, C.c_float), _p(pk.camera, C.c_float), C.c_int(width), C.c_int(height),
                              C.c_double(du), C.c_double(dv), _p(ids, C.c_int32), _p(t, C.c_double))
    return ids, t


def xorshift(s: int) -> int:
    return lib().orc_nb_xorshift(C.c_int64(s))


def tonemap(x: float) -> float:
    return lib().orc_nb_tonemap(C.c_double(x))


# ------------------------------------------------------------------------------ CPU-renderer family
class _CpuDesc(C.Structure):
    _fields_ = [("n_obj", C.c_int), ("type", C.POINTER(C.c_int32)), ("mat_id", C.POINTER(C.c_int32)),
                ("obj", C.POINTER(C.c_double)), ("mat", C.POINTER(C.c_double)),
                ("n_node", C.c_int), ("box", C.POINTER(C.c_double)), ("child", C.POINTER(C.c_int32)),
                ("lights", C.POINTER(C.c_double)), ("n_lights", C.c_int),
                ("light_color", C.POINTER(C.c_double)), ("ambient", C.POINTER(C.c_double)),
                ("tex", C.POINTER(C.c_uint8)), ("tex_info", C.POINTER(C.c_int32)), ("n_tex", C.c_int),
                ("cam", C.POINTER(C.c_double))]


@dataclass
class CpuExport:
    arrays: dict
    desc: _CpuDesc


def cpu_export(scene, camera) -> CpuExport:
    """Flatten Scene/Camera objects (un-rounded float64) + the BVH the builder made."""
    objs = list(scene.objects)
    index_of = {id(o): i for i, o in enumerate(objs)}
    n = len(objs)
    typ = np.zeros(n, dtype=np.int32); mat_id = np.zeros(n, dtype=np.int32)
    obj = np.zeros((n, 24)); mats, mat_index = [], {}
    tex_objs, tex_index = [], {}

    def tex_of(t):
        if t is None:
            return -1.0
        if id(t) not in tex_index:
            tex_index[id(t)] = len(tex_objs); tex_objs.append(t)
        return float(tex_index[id(t)])

    for i, o in enumerate(objs):
        k, m = _kind(o), o.material
        if id(m) not in mat_index:
            mat_index[id(m)] = len(mats)
            mats.append([m.color.x, m.color.y, m.color.z, m.diffuse, m.specular, m.reflective,
                         m.refractive, m.ior, tex_of(m.texture)])
        mat_id[i] = mat_index[id(m)]
        if k == "plane":
            typ[i] = 0
            obj[i, :14] = [o.anchor.x, o.anchor.y, o.anchor.z, o.normal.x, o.normal.y, o.normal.z,
                           o.u_unit.x, o.u_unit.y, o.u_unit.z, o.v_unit.x, o.v_unit.y, o.v_unit.z,
                           o.u_extent, o.v_extent]
        elif k == "sphere":
            typ[i] = 1
            obj[i, :4] = [o.center.x, o.center.y, o.center.z, o.radius]
        else:
            typ[i] = 2
            obj[i, :12] = [o.v0.x, o.v0.y, o.v0.z, o.v1.x, o.v1.y, o.v1.z, o.v2.x, o.v2.y, o.v2.z,
                           o.normal.x, o.normal.y, o.normal.z]
            if o.uv0 is not None:
                obj[i, 12:19] = [o.uv0[0], o.uv0[1], o.uv1[0], o.uv1[1], o.uv2[0], o.uv2[1], 1.0]

    boxes, children = [], []

    def walk(node):
        if id(node) in index_of and not hasattr(node, "left"):
            return ~index_of[id(node)]
        me = len(boxes)
        b = node.box
        boxes.append([b.min.x, b.min.y, b.min.z, b.max.x, b.max.y, b.max.z]); children.append([0, 0])
        children[me][0] = walk(node.left)
        children[me][1] = walk(node.right)
        return me

    if getattr(scene, "bvh_root", None) is not None:
        walk(scene.bvh_root)
    box = np.array(boxes, dtype=np.float64).reshape(-1, 6)
    child = np.array(children, dtype=np.int32).reshape(-1, 2)

    chunks, info, off = [], [], 0
    for t in tex_objs:
        px = np.ascontiguousarray(t.pixels, dtype=np.uint8).reshape(-1)
        h, w = t.pixels.shape[:2]
        info += [off, w, h]; off += px.size; chunks.append(px)
    tex = np.concatenate(chunks) if chunks else np.zeros(1, dtype=np.uint8)
    tex_info = np.array(info if info else [0, 1, 1], dtype=np.int32)

    arr = dict(
        type=typ, mat_id=mat_id, obj=obj, mat=np.array(mats, dtype=np.float64), box=box, child=child,
        lights=np.array([[l.x, l.y, l.z] for l in scene.lights], dtype=np.float64).reshape(-1, 3),
        light_color=np.array([scene.light_color.x, scene.light_color.y, scene.light_color.z]),
        ambient=np.array([scene.ambient.x, scene.ambient.y, scene.ambient.z]),
        tex=tex, tex_info=tex_info,
        cam=np.array([camera.origin.x, camera.origin.y, camera.origin.z,
                      camera.lower_left_corner.x, camera.lower_left_corner.y, camera.lower_left_corner.z,
                      camera.horizontal.x, camera.horizontal.y, camera.horizontal.z,
                      camera.vertical.x, camera.vertical.y, camera.vertical.z]))
    d = _CpuDesc(n, _p(typ, C.c_int32), _p(mat_id, C.c_int32), _p(obj, C.c_double), _p(arr["mat"], C.c_double),
                 box.shape[0], _p(box, C.c_double), _p(child, C.c_int32),
                 _p(arr["lights"], C.c_double), arr["lights"].shape[0],
                 _p(arr["light_color"], C.c_double), _p(arr["ambient"], C.c_double),
                 _p(tex, C.c_uint8), _p(tex_info, C.c_int32), len(tex_objs), _p(arr["cam"], C.c_double))
    return CpuExport(arr, d)


def cpu_whitted(exp: CpuExport, width, height, max_depth, jitter=None, want_rgb=True, want_ids=True):
    """-> dict(rgb [H,W,3] float64 (row 0 = bottom), ids [H,W] scene.objects index, t, hit_calls)."""
    rgb = np.zeros((height, width, 3)) if want_rgb else None
    ids = np.zeros((height, width), dtype=np.int32) if want_ids else None
    tt = np.zeros((height, width)) if want_ids else None
    if jitter is not None:
        jitter = np.ascontiguousarray(jitter, dtype=np.float64).reshape(height, width, 2)
    calls = C.c_uint64(0)
    lib().orc_cpu_whitted(C.byref(exp.desc), C.c_int(width), C.c_int(height), _p(jitter, C.c_double),
                          C.c_int(max_depth), _p(rgb, C.c_double), _p(ids, C.c_int32), _p(tt, C.c_double),
                          C.byref(calls))
    return dict(rgb=rgb, ids=ids, t=tt, hit_calls=calls.value)


def cpu_primary_ids_bruteforce(exp: CpuExport, width, height, du=0.5, dv=0.5):
    ids = np.zeros((height, width), dtype=np.int32); tt = np.zeros((height, width))
    lib().orc_cpu_primary_ids_bruteforce(C.byref(exp.desc), C.c_int(width), C.c_int(height),
                                         C.c_double(du), C.c_double(dv), _p(ids, C.c_int32), _p(tt, C.c_double))
    return ids, tt


def num_threads() -> int:
    return lib().orc_num_threads()


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(C.c_int(n))
