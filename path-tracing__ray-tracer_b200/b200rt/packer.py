"""Scene -> SoA record streams (the layouts of ``include/b200rt.h``).

Replaces the reference's host packers ``_prepare_scene_data`` / ``_prepare_texture_data`` /
``_prepare_camera_data`` / ``_prepare_light_data`` (``renderers/cuda_path_tracer.py:819-946``,
identical copies in ``cuda_texture_renderer.py:790-973``).  Differences by design:

* hot (intersection) and cold (shading) attributes go to separate float4 record streams instead of
  one AoS float block; derived quantities the reference recomputes per ray (rectangle unit axes,
  triangle edges, radius^2) are computed ONCE here, with the reference's own rounding;
* textures become RGBX8 (one 32-bit load per texel) and are taken from the already decoded
  ``Texture.pixels`` — the reference re-opens every JPEG and builds a 52 M-element Python list on
  every ``render()`` (4.7 s);
* ``semantics`` selects which of the reference's two arithmetic regimes is reproduced:
    ``"numba"`` : values rounded to float32 first, float32 arithmetic where Numba types it so
                  (f32 (+-*/) f32 and ``math.sqrt(f32)`` stay f32: unit axes :552-562, edges :669-675,
                  ``radius*radius`` :604); planes/triangles are never refractive, only triangles are
                  textured (:573-574, :630-631, :727-728);
    ``"cpu"``   : un-rounded float64 objects as ``core/geometry.py`` holds them.

Primitive ids ("packed order") = rectangles, then spheres, then triangles, each in ``scene.objects``
order — the scan order of ``cuda_scene_hit`` (:511,:582,:639).  ``PackedScene.order`` maps a packed id
back to ``(index in scene.objects, face index or -1)``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

SEM_NUMBA, SEM_CPU = 0, 1
_SEM = {"numba": SEM_NUMBA, "cpu": SEM_CPU}


class TriangleMesh:
    """Bulk triangle container (side door for scenes the object-per-triangle API cannot express).

    ``vertices`` float [nv, 3], ``faces`` int [nf, 3], optional per-vertex ``uvs`` float [nv, 2].
    Drop it into ``scene.objects``; the packer expands it (vectorised) into ``nf`` triangles whose
    geometric normal is ``normalize((v1-v0) x (v2-v0))`` like ``Triangle.__init__``
    (``core/geometry.py:130``).
    """

    def __init__(self, vertices, faces, material, uvs=None):
        self.vertices = np.ascontiguousarray(vertices, dtype=np.float64).reshape(-1, 3)
        self.faces = np.ascontiguousarray(faces, dtype=np.int64).reshape(-1, 3)
        self.uvs = None if uvs is None else np.ascontiguousarray(uvs, dtype=np.float64).reshape(-1, 2)
        self.material = material

    def spatially_sorted(self) -> "TriangleMesh":
        """The same mesh with its faces listed along a Morton curve through their centroids (vertices untouched).

        Packed triangle records follow the face order, so neighbours in space become neighbours in memory: the leaf
        tests of rays that travel together then read the same cache lines.  Closest hits do not depend on the order
        except where two triangles are hit at exactly the same distance (the tie goes to the lower packed id)."""
        return TriangleMesh(self.vertices, self.faces[morton_face_order(self.vertices, self.faces)], self.material, self.uvs)


def morton_face_order(vertices: np.ndarray, faces: np.ndarray, max_bits: int = 30) -> np.ndarray:
    """Permutation that sorts the faces along a Morton curve through their centroids (stable: equal codes keep their order).

    Bits per axis follow the rule of the LBVH builder (``lbvh.cu:morton_bits``): an axis gets
    ``log2(centroid spread / mean face extent) + 1`` bits — finer cells than the faces cannot separate them — so the
    flat axis of a terrain or a city does not scatter neighbours; bits are interleaved from the top, always taking the
    next bit of the axis that has the most left."""
    v = np.asarray(vertices, dtype=np.float64).reshape(-1, 3)
    f = np.asarray(faces, dtype=np.int64).reshape(-1, 3)
    if f.shape[0] == 0:
        return np.zeros(0, dtype=np.int64)
    p0, p1, p2 = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    c = (p0 + p1 + p2) / 3.0
    ext = (np.maximum(np.maximum(p0, p1), p2) - np.minimum(np.minimum(p0, p1), p2)).mean(0)
    lo, hi = c.min(0), c.max(0)
    spread = hi - lo
    bits = [0, 0, 0]
    for k in range(3):
        if spread[k] > 0:
            bits[k] = int(np.clip(np.floor(np.log2(spread[k] / max(ext[k], 1e-300))) + 1, 0, 16))
    while sum(bits) > max_bits:
        bits[int(np.argmax(bits))] -= 1
    if sum(bits) == 0:
        return np.arange(f.shape[0], dtype=np.int64)
    q = [np.minimum(((c[:, k] - lo[k]) / max(spread[k], 1e-300) * (1 << bits[k])).astype(np.int64), (1 << bits[k]) - 1)
         if bits[k] else None for k in range(3)]
    code = np.zeros(f.shape[0], dtype=np.int64)
    rem = list(bits)
    for _ in range(sum(bits)):
        k = max(range(3), key=lambda a: (rem[a], -a))
        rem[k] -= 1
        code = (code << 1) | ((q[k] >> rem[k]) & 1)
    return np.argsort(code, kind="stable")


def load_obj(path: str, material, scale: float = 1.0, translate=(0.0, 0.0, 0.0)) -> TriangleMesh:
    """Minimal Wavefront OBJ reader (v / vt / f with fan triangulation) -> ``TriangleMesh``."""
    verts, uvs, faces, face_uv = [], [], [], []
    with open(path) as f:
        for line in f:
            p = line.split()
            if not p:
                continue
            if p[0] == "v":
                verts.append([float(x) for x in p[1:4]])
            elif p[0] == "vt":
                uvs.append([float(x) for x in p[1:3]])
            elif p[0] == "f":
                idx = [tok.split("/") for tok in p[1:]]
                vi = [int(t[0]) - 1 if int(t[0]) > 0 else len(verts) + int(t[0]) for t in idx]
                ti = [int(t[1]) - 1 if len(t) > 1 and t[1] else -1 for t in idx]
                for k in range(1, len(vi) - 1):
                    faces.append([vi[0], vi[k], vi[k + 1]])
                    face_uv.append([ti[0], ti[k], ti[k + 1]])
    V = np.asarray(verts, dtype=np.float64) * scale + np.asarray(translate, dtype=np.float64)
    F = np.asarray(faces, dtype=np.int64)
    mesh_uv = None
    if uvs and all(t >= 0 for tri in face_uv for t in tri):
        # per-corner uv: duplicate vertices so that uvs are per vertex
        FU = np.asarray(face_uv, dtype=np.int64)
        V = V[F.reshape(-1)]
        mesh_uv = np.asarray(uvs, dtype=np.float64)[FU.reshape(-1)]
        F = np.arange(F.size).reshape(-1, 3)
    return TriangleMesh(V, F, material, mesh_uv)


def kind_of(obj) -> str:
    if isinstance(obj, TriangleMesh):
        return "mesh"
    if hasattr(obj, "anchor") and hasattr(obj, "u_dir"):
        return "plane"
    if hasattr(obj, "center") and hasattr(obj, "radius"):
        return "sphere"
    if hasattr(obj, "v0") and hasattr(obj, "v2"):
        return "triangle"
    raise TypeError(f"b200rt cannot pack object of type {type(obj).__name__}")


def texture_paths_sorted(scene) -> List[str]:
    """Distinct ``Texture.path`` strings, sorted: the reference's texture-id rule (:824-832)."""
    seen: List[str] = []
    for o in scene.objects:
        tex = getattr(getattr(o, "material", None), "texture", None)
        if tex is not None:
            p = getattr(tex, "path", None)
            if p and p not in seen:
                seen.append(p)
    return sorted(seen)


def scene_signature(scene) -> tuple:
    """Every value ``pack_scene`` reads from ``scene`` as one hashable tuple (the cache key of the renderers'
    packed-scene cache): object kinds and all their float fields, materials, texture paths, lights, light colour and
    ambient.  Meshes contribute the identity, shape and a strided fingerprint of their arrays."""
    sig = []
    for o in scene.objects:
        k, m = kind_of(o), o.material
        t = getattr(m, "texture", None)
        ms = (m.color.x, m.color.y, m.color.z, m.diffuse, m.specular, m.reflective, getattr(m, "refractive", 0.0),
              getattr(m, "ior", 1.0), getattr(t, "path", None) if t is not None else None)
        if k == "plane":
            sig.append((0, *_v(o.anchor), *_v(o.normal), *_v(o.u_dir), *_v(o.v_dir), o.u_len, o.v_len,
                        *(_v(o.u_unit) if hasattr(o, "u_unit") else ()), getattr(o, "u_extent", None), getattr(o, "v_extent", None), ms))
        elif k == "sphere":
            sig.append((1, *_v(o.center), o.radius, ms))
        elif k == "triangle":
            sig.append((2, *_v(o.v0), *_v(o.v1), *_v(o.v2), *_v(o.normal),
                        tuple(o.uv0) if o.uv0 is not None else None, tuple(o.uv1) if o.uv1 is not None else None,
                        tuple(o.uv2) if o.uv2 is not None else None, ms))
        else:
            V, F = o.vertices, o.faces
            sig.append((3, id(V), V.shape, float(V[:: max(1, V.shape[0] // 1024)].sum()), id(F), F.shape,
                        int(F[:: max(1, F.shape[0] // 1024)].sum()), None if o.uvs is None else (id(o.uvs), o.uvs.shape), ms))
    lc, am = getattr(scene, "light_color", None), getattr(scene, "ambient", None)
    return (tuple(sig), tuple(_v(l) for l in scene.lights), _v(lc) if lc is not None else None, _v(am) if am is not None else None)


@dataclass
class PackedScene:
    semantics: int
    n_rect: int
    n_sphere: int
    n_tri: int
    rect: np.ndarray          # float64 [4*n_rect, 4]
    sphere: np.ndarray        # float64 [2*n_sphere, 4]
    tri: np.ndarray           # float64 [3*n_tri, 4]
    shade: np.ndarray         # float64 [3*n_prims, 4]
    prim_mat: np.ndarray      # int32 [n_prims]
    mat: np.ndarray           # float64 [2*n_mat, 4]
    mat_tex: np.ndarray       # int32 [n_mat]
    texels: np.ndarray        # uint32 [total texels]  (R | G<<8 | B<<16 | 0xFF<<24)
    tex_info: np.ndarray      # int32 [n_tex, 4]: offset, w, h, 0
    lights: np.ndarray        # float64 [n_lights, 4]
    order: np.ndarray         # int32 [n_prims, 2]
    light_color: np.ndarray = field(default_factory=lambda: np.ones(3))
    ambient: np.ndarray = field(default_factory=lambda: np.full(3, 0.5))

    @property
    def n_prims(self) -> int:
        return self.n_rect + self.n_sphere + self.n_tri

    @property
    def n_mat(self) -> int:
        return int(self.mat_tex.shape[0])

    @property
    def n_tex(self) -> int:
        return int(self.tex_info.shape[0])

    def bounds(self, pad_rel: float = 1e-4):
        """(lo[3], hi[3]) of all primitives, padded outward (float32).  An empty scene gives lo > hi."""
        lo, hi = np.full(3, np.inf), np.full(3, -np.inf)
        if self.n_rect:
            r = self.rect.reshape(-1, 4, 4)
            a = r[:, 0, :3]
            u, v = r[:, 2, :3] * r[:, 0, 3:4], r[:, 3, :3] * r[:, 1, 3:4]
            for c in (a, a + u, a + v, a + u + v):
                lo, hi = np.minimum(lo, c.min(0)), np.maximum(hi, c.max(0))
        if self.n_sphere:
            sp = self.sphere.reshape(-1, 2, 4)
            lo = np.minimum(lo, (sp[:, 0, :3] - sp[:, 0, 3:4]).min(0)); hi = np.maximum(hi, (sp[:, 0, :3] + sp[:, 0, 3:4]).max(0))
        if self.n_tri:
            t = self.tri.reshape(-1, 3, 4)
            for c in (t[:, 0, :3], t[:, 0, :3] + t[:, 1, :3], t[:, 0, :3] + t[:, 2, :3]):
                lo, hi = np.minimum(lo, c.min(0)), np.maximum(hi, c.max(0))
        if not np.all(np.isfinite(lo)):
            return np.ones(3, np.float32), -np.ones(3, np.float32)
        pad = pad_rel * max(float(np.abs(lo).max()), float(np.abs(hi).max()), 1e-3)
        return (lo - pad).astype(np.float32), (hi + pad).astype(np.float32)

    def max_abs_coordinate(self) -> float:
        m = 0.0
        if self.n_rect:
            r = self.rect.reshape(-1, 4, 4)
            m = max(m, float(np.abs(r[:, 0, :3]).max() + max(r[:, 0, 3].max(), r[:, 1, 3].max())))
        if self.n_sphere:
            s = self.sphere.reshape(-1, 2, 4)
            m = max(m, float((np.abs(s[:, 0, :3]).max(axis=1) + s[:, 0, 3]).max()))
        if self.n_tri:
            t = self.tri.reshape(-1, 3, 4)
            m = max(m, float(np.abs(t[:, 0, :3]).max() + np.abs(t[:, 1:, :3]).max()))
        return m


def _f32(x):
    return np.asarray(x, dtype=np.float32)


def _v(p):
    return (p.x, p.y, p.z)


class _Materials:
    def __init__(self):
        self.index: Dict[tuple, int] = {}
        self.rows: List[tuple] = []

    def get(self, color, diffuse, specular, reflective, refractive, ior, tex) -> int:
        key = (float(color[0]), float(color[1]), float(color[2]), float(diffuse), float(specular),
               float(reflective), float(refractive), float(ior), int(tex))
        if key not in self.index:
            self.index[key] = len(self.rows)
            self.rows.append(key)
        return self.index[key]


def pack_camera(camera, semantics: str = "numba") -> np.ndarray:
    """12 doubles origin, lower_left, horizontal, vertical (``_prepare_camera_data`` :934-940).
    numba semantics round every component to float32 like the reference's ``np.float32`` array."""
    c = np.array([*_v(camera.origin), *_v(camera.lower_left_corner), *_v(camera.horizontal), *_v(camera.vertical)],
                 dtype=np.float64)
    if _SEM[semantics] == SEM_NUMBA:
        c = c.astype(np.float32).astype(np.float64)
    return c


def pack_textures(scene):
    """-> (texels uint32, tex_info int32 [n,4], {path: id}).  Uses the decoded ``Texture.pixels``."""
    paths = texture_paths_sorted(scene)
    by_path = {}
    for o in scene.objects:
        t = getattr(getattr(o, "material", None), "texture", None)
        if t is not None and getattr(t, "path", None):
            by_path.setdefault(t.path, t)
    chunks, info, off = [], [], 0
    for p in paths:
        px = np.ascontiguousarray(by_path[p].pixels, dtype=np.uint8)
        h, w = px.shape[:2]
        rgbx = np.empty((h, w, 4), dtype=np.uint8)
        rgbx[..., :3] = px[..., :3]
        rgbx[..., 3] = 255
        chunks.append(rgbx.reshape(-1).view("<u4"))
        info.append((off, w, h, 0))
        off += h * w
    texels = np.concatenate(chunks) if chunks else np.zeros(1, dtype=np.uint32)
    tex_info = np.array(info, dtype=np.int32).reshape(-1, 4)
    return texels, tex_info, {p: i for i, p in enumerate(paths)}


def pack_scene(scene, semantics: str = "numba", textures=None) -> PackedScene:
    sem = _SEM[semantics]
    nb = sem == SEM_NUMBA
    if textures is None:
        textures = pack_textures(scene)
    texels, tex_info, tex_id = textures
    if texels is None:                       # textures live on the device already (renderer._TextureCache)
        texels = np.zeros(1, dtype=np.uint32)
    mats = _Materials()

    def rnd(a):
        a = np.asarray(a, dtype=np.float64)
        return a.astype(np.float32).astype(np.float64) if nb else a

    def tex_of(m) -> int:
        t = getattr(m, "texture", None)
        if t is None:
            return -1
        return tex_id.get(getattr(t, "path", None), -1)

    rects, spheres, tri_hot, tri_cold = [], [], [], []
    rect_mat, sph_mat, tri_mat = [], [], []
    rect_nrm = []
    o_rect, o_sph, o_tri = [], [], []

    for idx, o in enumerate(scene.objects):
        k, m = kind_of(o), o.material
        col = rnd(_v(m.color))
        dif, spe, refl = (float(rnd(x)) for x in (m.diffuse, m.specular, m.reflective))
        refr = float(rnd(getattr(m, "refractive", 0.0)))
        ior = float(rnd(getattr(m, "ior", 1.0)))
        if k == "plane":
            anchor, normal = rnd(_v(o.anchor)), rnd(_v(o.normal))
            ul, vl = float(rnd(o.u_len)), float(rnd(o.v_len))
            if nb:      # in-kernel float32 normalisation of u_dir / v_dir (:552-562)
                ud, vd = _f32(_v(o.u_dir)), _f32(_v(o.v_dir))
                un = np.sqrt(ud[0] * ud[0] + ud[1] * ud[1] + ud[2] * ud[2], dtype=np.float32)
                vn = np.sqrt(vd[0] * vd[0] + vd[1] * vd[1] + vd[2] * vd[2], dtype=np.float32)
                if not (un > 0 and vn > 0):
                    continue                     # the reference can never hit such a plane (:555)
                uu, vv = (ud / un).astype(np.float64), (vd / vn).astype(np.float64)
                mid = mats.get(col, dif, spe, refl, 0.0, 1.0, -1)
            else:       # CPU path: v_unit = normal x u_unit (core/geometry.py:35-36)
                uu, vv = np.array(_v(o.u_unit)), np.array(_v(o.v_unit))
                ul, vl = float(o.u_extent), float(o.v_extent)
                mid = mats.get(col, dif, spe, refl, refr, ior, tex_of(m))
            rects.append([[*anchor, ul], [*normal, vl], [*uu, 0.0], [*vv, 0.0]])
            rect_nrm.append(normal)
            rect_mat.append(mid)
            o_rect.append((idx, -1))
        elif k == "sphere":
            c, r = rnd(_v(o.center)), float(rnd(o.radius))
            r2 = float(np.float32(r) * np.float32(r)) if nb else r * r
            spheres.append([[*c, r], [r2, 0.0, 0.0, 0.0]])
            sph_mat.append(mats.get(col, dif, spe, refl, refr, ior, -1 if nb else tex_of(m)))
            o_sph.append((idx, -1))
        elif k == "triangle":
            v0, v1, v2 = (rnd(_v(p)) for p in (o.v0, o.v1, o.v2))
            if nb:
                e1 = (_f32(v1) - _f32(v0)).astype(np.float64)
                e2 = (_f32(v2) - _f32(v0)).astype(np.float64)
            else:
                e1, e2 = v1 - v0, v2 - v0
            n = rnd(_v(o.normal))
            if o.uv0 is not None:
                uv = [float(o.uv0[0]), float(o.uv0[1]), float(o.uv1[0]), float(o.uv1[1]),
                      float(o.uv2[0]), float(o.uv2[1])]
                has_uv = 1.0
            else:           # defaults of the reference packer (:869-874); the CPU path has no uv at all
                uv, has_uv = [0.0, 0.0, 1.0, 0.0, 1.0, 1.0], (1.0 if nb else 0.0)
            uv = list(rnd(uv))
            tri_hot.append([[*v0, 0.0], [*e1, 0.0], [*e2, 0.0]])
            tri_cold.append([[*n, 0.0], uv[:4], [uv[4], uv[5], has_uv, 0.0]])
            tri_mat.append(mats.get(col, dif, spe, refl, 0.0 if nb else refr, 1.0 if nb else ior, tex_of(m)))
            o_tri.append((idx, -1))
        else:   # mesh
            V = rnd(o.vertices)
            F = o.faces
            v0, v1, v2 = V[F[:, 0]], V[F[:, 1]], V[F[:, 2]]
            if nb:
                e1 = (v1.astype(np.float32) - v0.astype(np.float32)).astype(np.float64)
                e2 = (v2.astype(np.float32) - v0.astype(np.float32)).astype(np.float64)
            else:
                e1, e2 = v1 - v0, v2 - v0
            Vd = np.asarray(o.vertices, dtype=np.float64)
            nn = np.cross(Vd[F[:, 1]] - Vd[F[:, 0]], Vd[F[:, 2]] - Vd[F[:, 0]])
            ln = np.linalg.norm(nn, axis=1, keepdims=True)
            nn = rnd(np.divide(nn, ln, out=np.zeros_like(nn), where=ln > 0))
            nf = F.shape[0]
            hot = np.zeros((nf, 3, 4)); hot[:, 0, :3] = v0; hot[:, 1, :3] = e1; hot[:, 2, :3] = e2
            cold = np.zeros((nf, 3, 4)); cold[:, 0, :3] = nn
            if o.uvs is not None:
                U = rnd(o.uvs)
                cold[:, 1, 0:2] = U[F[:, 0]]; cold[:, 1, 2:4] = U[F[:, 1]]; cold[:, 2, 0:2] = U[F[:, 2]]
                cold[:, 2, 2] = 1.0
            else:
                cold[:, 1, :] = (0.0, 0.0, 1.0, 0.0); cold[:, 2, 0:2] = (1.0, 1.0)
                cold[:, 2, 2] = 1.0 if nb else 0.0
            mid = mats.get(col, dif, spe, refl, 0.0 if nb else refr, 1.0 if nb else ior, tex_of(m))
            tri_hot.append(hot); tri_cold.append(cold)
            tri_mat.append(np.full(nf, mid, dtype=np.int32))
            o_tri.append(np.stack([np.full(nf, idx), np.arange(nf)], axis=1))

    def cat(parts, width):
        arrs = [np.asarray(p, dtype=np.float64).reshape(-1, width, 4) for p in parts]
        return np.concatenate(arrs).reshape(-1, 4) if arrs else np.zeros((0, 4))

    rect = cat(rects, 4); sphere = cat(spheres, 2); tri = cat(tri_hot, 3)
    n_rect, n_sphere, n_tri = rect.shape[0] // 4, sphere.shape[0] // 2, tri.shape[0] // 3
    shade = np.zeros((3 * (n_rect + n_sphere + n_tri), 4))
    if n_rect:
        shade.reshape(-1, 3, 4)[:n_rect, 0, :3] = np.asarray(rect_nrm)
    if n_tri:
        shade.reshape(-1, 3, 4)[n_rect + n_sphere:] = cat(tri_cold, 3).reshape(-1, 3, 4)

    def cat_i(parts):
        flat = [np.atleast_1d(np.asarray(p, dtype=np.int32)) for p in parts]
        return np.concatenate(flat) if flat else np.zeros(0, dtype=np.int32)

    prim_mat = np.concatenate([cat_i(rect_mat), cat_i(sph_mat), cat_i(tri_mat)]).astype(np.int32)
    order_parts = [np.asarray(p, dtype=np.int32).reshape(-1, 2) for p in (o_rect, o_sph)] + \
                  [np.asarray(p, dtype=np.int32).reshape(-1, 2) for p in o_tri]
    order = np.concatenate(order_parts) if order_parts else np.zeros((0, 2), dtype=np.int32)

    mat = np.zeros((2 * max(1, len(mats.rows)), 4))
    mat_tex = np.full(max(1, len(mats.rows)), -1, dtype=np.int32)
    for i, row in enumerate(mats.rows):
        mat[2 * i] = row[0:4]
        mat[2 * i + 1] = row[4:8]
        mat_tex[i] = row[8]

    lights = np.zeros((len(scene.lights), 4))
    for i, l in enumerate(scene.lights):
        lights[i, :3] = rnd(_v(l))

    lc = getattr(scene, "light_color", None)
    am = getattr(scene, "ambient", None)
    return PackedScene(sem, n_rect, n_sphere, n_tri, rect, sphere, tri, shade, prim_mat, mat, mat_tex,
                       texels, tex_info, lights, order,
                       np.array(_v(lc)) if lc is not None else np.ones(3),
                       np.array(_v(am)) if am is not None else np.full(3, 0.5))


def build_scan_prims(p: PackedScene, pair_tol: float = 1e-5, quads_out: list = None) -> np.ndarray:
    """Small-scene scan records (``b2rt_scene.d_scan_prims``): float32 [4*n, 4].

    Every rectangle, triangle, and coplanar triangle PAIR ``(p0,p1,p2), (p0,p2,p3)`` with
    ``p2 = p1 + p3 - p0`` (a parallelogram split along its diagonal, as ``_create_single_cube`` /
    ``_create_canvas`` emit them, ``custom_scene_builder.py:356-366,470-476``) becomes one record
    "plane + two edge planes":  t = (cN - N.o)/(N.d),  P = o + t d,  u = n1.P + d1,  v = n2.P + d2.
    N is the unit normal for rectangles and e1 x e2 (un-normalised) for triangles, so the |N.d| > 1e-6
    guard equals the reference's ``abs(denom)`` / ``abs(a)`` guards (``cuda_path_tracer.py:536,683``).
    """
    recs = []
    quads = quads_out if quads_out is not None else []      # per record: (corner, edge a, edge b) or None
    R = p.rect.reshape(-1, 4, 4)
    for i in range(p.n_rect):
        anchor, ul = R[i, 0, :3], R[i, 0, 3]
        n, vl = R[i, 1, :3], R[i, 1, 3]
        uu, vv = R[i, 2, :3], R[i, 3, :3]
        recs.append(([*n, n @ anchor], [*uu, -(uu @ anchor)], [*vv, -(vv @ anchor)], (ul, vl), 0, i, 0))
        quads.append((anchor.astype(np.float64), uu.astype(np.float64) * ul, vv.astype(np.float64) * vl))
    T = p.tri.reshape(-1, 3, 4)
    base = p.n_rect + p.n_sphere
    used = np.zeros(p.n_tri, dtype=bool)

    def edge_planes(v0, e1, e2):
        N = np.cross(e1, e2)
        a1 = np.cross(e2, N); a1 = a1 / (e1 @ a1)
        a2 = np.cross(N, e1); a2 = a2 / (e2 @ a2)
        return N, a1, a2

    for i in range(p.n_tri):
        if used[i]:
            continue
        v0, e1, e2 = T[i, 0, :3], T[i, 1, :3], T[i, 2, :3]
        scale = max(np.abs(v0).max(), np.abs(e1).max(), np.abs(e2).max(), 1e-30)
        mate = -1
        if p.n_tri <= 4096:                       # pairing is for small scenes; large ones walk the LBVH
            for j in range(p.n_tri):
                if j == i or used[j]:
                    continue
                w0, f1, f2 = T[j, 0, :3], T[j, 1, :3], T[j, 2, :3]
                # j = (p0, p2, p3) with p2 = i.v2 and p3 = p0 + (e2_i - e1_i)
                if np.abs(w0 - v0).max() <= pair_tol * scale and np.abs(f1 - e2).max() <= pair_tol * scale \
                        and np.abs(f2 - (e2 - e1)).max() <= pair_tol * scale:
                    mate = j
                    break
        used[i] = True
        if mate >= 0:
            used[mate] = True
            eq1, eq2 = e1, T[mate, 2, :3]          # parallelogram edges p1-p0 and p3-p0
            N, a1, a2 = edge_planes(v0, eq1, eq2)
            kind = 2 if i < mate else 3            # diagonal ties go to the lower packed id
            recs.append(([*N, N @ v0], [*a1, -(a1 @ v0)], [*a2, -(a2 @ v0)], (1.0, 1.0), kind, base + i, base + mate))
            quads.append((v0.astype(np.float64), eq1.astype(np.float64), eq2.astype(np.float64)))
        else:
            N, a1, a2 = edge_planes(v0, e1, e2)
            recs.append(([*N, N @ v0], [*a1, -(a1 @ v0)], [*a2, -(a2 @ v0)], (1.0, 1.0), 1, base + i, 0))
            quads.append(None)
    out = np.zeros((len(recs), 4, 4), dtype=np.float32)
    for k, (q0, q1, q2, lim, kind, ida, idb) in enumerate(recs):
        out[k, 0], out[k, 1], out[k, 2] = q0, q1, q2
        out[k, 3, 0], out[k, 3, 1] = lim
        out[k, 3, 2:4] = np.array([(kind << 28) | ida, idb], dtype=np.int32).view(np.float32)
    return out.reshape(-1, 4)


def group_scan_boxes(rec: np.ndarray, quads: list, tol: float = 1e-5):
    """Groups parallelogram scan records that are faces of a common parallelepiped (the Cornell walls, every
    cube) into BOX records -> (records reordered loose-first, n_loose, boxes float32 [4*n_box, 4]).

    A ray crosses the boundary of a convex box at most twice, at t_enter and t_exit of a three-slab test in the
    box's own coordinates l = M (P - C) in [-1, 1]^3.  Whatever subset of the six faces exists, the closest hit
    on the group is the first of those two crossings that lies beyond t_min AND whose face exists — one ~55
    instruction test instead of up to six ~40 instruction plane tests.  Box record (4 float4):
        (m0.xyz, d0) (m1.xyz, d1) (m2.xyz, d2)   l_k = m_k . P + d_k
        (bits(f0 | f1<<8 | f2<<16 | f3<<24), bits(f4 | f5<<8), 0, 0)   f[2k + (l_k = +1)] = index of the face's
        planar record (kept in the array after the loose ones: the hit's (u, v) / triangle id come from it), 255 = no face
    Needs an OPPOSITE pair of faces to define the box; groups with < 3 faces are not worth a box test.
    """
    rec = np.asarray(rec, dtype=np.float32).reshape(-1, 4, 4)
    n = rec.shape[0]
    faces = {}
    for k, q in enumerate(quads):
        if q is not None:
            p0, ea, eb = q
            faces[k] = (p0 + 0.5 * (ea + eb), 0.5 * ea, 0.5 * eb)

    def same_dir(x, y, scale):
        return min(np.abs(x - y).max(), np.abs(x + y).max()) <= tol * scale

    def same_edges(a, b, x, y, scale):
        return (same_dir(a, x, scale) and same_dir(b, y, scale)) or (same_dir(a, y, scale) and same_dir(b, x, scale))

    cands = []
    keys = sorted(faces)
    for ii, i in enumerate(keys):
        ci, ai, bi = faces[i]
        for j in keys[ii + 1:]:
            cj, aj, bj = faces[j]
            scale = max(np.abs(ai).max(), np.abs(bi).max(), np.abs(cj - ci).max(), 1e-30)
            if not same_edges(ai, bi, aj, bj, scale):
                continue
            h = 0.5 * (cj - ci)
            Hm = np.stack([ai, bi, h], axis=1)               # columns = half axes
            if abs(np.linalg.det(Hm)) <= 1e-9 * scale ** 3:
                continue
            C = 0.5 * (ci + cj)
            slots = {4: i, 5: j}
            for f in keys:
                if f in (i, j):
                    continue
                cf, af, bf = faces[f]
                for k in (0, 1):
                    others = (Hm[:, 1 - k], h)
                    for sgn in (-1, 1):
                        if np.abs(cf - (C + sgn * Hm[:, k])).max() <= tol * scale and \
                                same_edges(af, bf, others[0], others[1], scale):
                            slots.setdefault(2 * k + (1 if sgn > 0 else 0), f)
            cands.append((len(slots), C, Hm, slots))
    cands.sort(key=lambda c: -c[0])
    taken, boxes = set(), []
    for cnt, C, Hm, slots in cands:
        if cnt < 3 or any(f in taken for f in slots.values()):
            continue
        taken.update(slots.values())
        boxes.append((C, Hm, slots))
    loose = [k for k in range(n) if k not in taken]
    order = loose + sorted(taken)
    new_index = {old: new for new, old in enumerate(order)}
    out = np.zeros((len(boxes), 4, 4), dtype=np.float32)
    for b, (C, Hm, slots) in enumerate(boxes):
        M = np.linalg.inv(Hm)                                # rows m_k: l = M (P - C)
        out[b, :3, :3] = M
        out[b, :3, 3] = -(M @ C)
        code = [255] * 8
        for slot, f in slots.items():
            code[slot] = new_index[f]
        assert max(new_index.values(), default=0) < 255
        w = np.array([code[0] | code[1] << 8 | code[2] << 16 | code[3] << 24, code[4] | code[5] << 8],
                     dtype=np.uint32)
        out[b, 3, 0:2] = w.view(np.float32)
        # third info word, bit 0: CLOSED box (all six faces exist) -> the kernels take the tag-free three-slab test
        out[b, 3, 2] = np.array([1 if all(c != 255 for c in code[:6]) else 0], dtype=np.uint32).view(np.float32)[0]
    return rec[order].reshape(-1, 4), len(loose), out.reshape(-1, 4)


def build_surface_records(p: PackedScene) -> np.ndarray:
    """Per-primitive shading records for small scenes (``b2rt_scene.d_surface_records``): float32 [5*n_prims, 4].

    Everything ``cuda_scene_hit`` returns beside t (normal, uv, material; ``cuda_path_tracer.py:568-574,616-631,
    706-728``) in ONE branch-free, one-hop record per primitive, staged in shared memory — instead of the per-type
    branches and the prim -> material -> texture chain of dependent global loads of the generic streams:
        (n.xyz | sphere centre.xyz, 1/radius or 0)  (color.rgb, diffuse)  (specular, reflective, refractive, ior)
        (u0, du_a, du_b, bits(texture id))  (v0, dv_a, dv_b, bits(flags))     uv = uv0 + a * d_a + b * d_b
    a, b = the hit's rectangle coordinates / triangle barycentrics; flags bit 0: flip the normal to face the ray
    (triangles, ``dot(n, d) > 0``).  Rectangles: u = a / u_len, v = b / v_len; triangles: w uv0 + a uv1 + b uv2.
    """
    n = p.n_prims
    out = np.zeros((n, 5, 4), dtype=np.float32)
    R, S = p.rect.reshape(-1, 4, 4), p.sphere.reshape(-1, 2, 4)
    SH = p.shade.reshape(-1, 3, 4)
    flags = np.zeros(n, dtype=np.int32)
    for i in range(p.n_rect):
        out[i, 0, :3] = R[i, 1, :3]
        out[i, 3, 1] = 1.0 / R[i, 0, 3]
        out[i, 4, 2] = 1.0 / R[i, 1, 3]
    for i in range(p.n_sphere):
        k = p.n_rect + i
        out[k, 0, :3] = S[i, 0, :3]
        out[k, 0, 3] = 1.0 / S[i, 0, 3]
    base = p.n_rect + p.n_sphere
    for i in range(p.n_tri):
        k = base + i
        out[k, 0, :3] = SH[k, 0, :3]
        flags[k] = 1
        if SH[k, 2, 2] != 0:
            uv0, uv1, uv2 = SH[k, 1, 0:2], SH[k, 1, 2:4], SH[k, 2, 0:2]
            out[k, 3, :3] = (uv0[0], uv1[0] - uv0[0], uv2[0] - uv0[0])
            out[k, 4, :3] = (uv0[1], uv1[1] - uv0[1], uv2[1] - uv0[1])
    M = p.mat.reshape(-1, 2, 4)
    for k in range(n):
        m = int(p.prim_mat[k])
        out[k, 1], out[k, 2] = M[m, 0], M[m, 1]
    out[:, 3, 3] = p.mat_tex[p.prim_mat[:n]].astype(np.int32).view(np.float32) if n else 0
    out[:, 4, 3] = flags.view(np.float32)
    return out.reshape(-1, 4)


def rect_scan_records(p: PackedScene) -> np.ndarray:
    """Planar records of the rectangles only (candidate occluders of scenes that walk the LBVH)."""
    R = p.rect.reshape(-1, 4, 4)
    out = np.zeros((p.n_rect, 4, 4), dtype=np.float32)
    for i in range(p.n_rect):
        anchor, n, uu, vv = R[i, 0, :3], R[i, 1, :3], R[i, 2, :3], R[i, 3, :3]
        out[i, 0], out[i, 1], out[i, 2] = [*n, n @ anchor], [*uu, -(uu @ anchor)], [*vv, -(vv @ anchor)]
        out[i, 3, 0], out[i, 3, 1] = R[i, 0, 3], R[i, 1, 3]
        out[i, 3, 2:4] = np.array([i, 0], dtype=np.int32).view(np.float32)
    return out.reshape(-1, 4)


def build_occluder_hints(p: PackedScene, scan: np.ndarray, n_points: int = 4096, seed: int = 0,
                         generic: bool = False) -> np.ndarray:
    """For every light sample, the scan record (code k) or sphere (code 64 + i) that blocks the most
    next-event shadow rays towards it -> int32 [n_lights] (-1: none).

    Performance hint only (``b2rt_scene.d_occluder_hint``): the shade stage tests this one primitive
    before queueing a shadow ray, and a hit answers the occlusion query exactly.  In the reference's
    Cornell box the light samples sit at y = 14 *below* the ceiling at y = 15 and NEE shadow rays run to
    t_max = 1e6 (``cuda_path_tracer.py:275-277``), so the ceiling blocks ~92 % of them.
    Estimated here from area-weighted random surface points, both sides of every surface.
    ``generic=True``: ``scan`` holds ``rect_scan_records`` and the codes are 128 + packed id (rectangle i ->
    128 + i, sphere i -> 128 + n_rect + i), for scenes that have no scan records.
    """
    rng = np.random.default_rng(seed)
    n_lights = p.lights.shape[0]
    hints = np.full(max(1, n_lights), -1, dtype=np.int32)
    rec = np.asarray(scan, dtype=np.float64).reshape(-1, 4, 4)
    if n_lights == 0 or rec.shape[0] == 0:
        return hints
    pts, nrm, area = [], [], []
    R = p.rect.reshape(-1, 4, 4)
    for i in range(p.n_rect):
        area.append(("r", i, R[i, 0, 3] * R[i, 1, 3]))
    S = p.sphere.reshape(-1, 2, 4)
    for i in range(p.n_sphere):
        area.append(("s", i, 4 * np.pi * S[i, 0, 3] ** 2))
    T = p.tri.reshape(-1, 3, 4)
    tri_ids = range(p.n_tri) if p.n_tri <= 2048 else rng.choice(p.n_tri, 2048, replace=False)
    for i in tri_ids:
        area.append(("t", int(i), 0.5 * np.linalg.norm(np.cross(T[i, 1, :3], T[i, 2, :3])) * (p.n_tri / len(tri_ids))))
    total = sum(a for _, _, a in area) or 1.0
    for kind, i, a in area:
        m = max(4, int(round(n_points * a / total)))
        u, v = rng.random(m), rng.random(m)
        if kind == "r":
            P = R[i, 0, :3] + np.outer(u * R[i, 0, 3], R[i, 2, :3]) + np.outer(v * R[i, 1, 3], R[i, 3, :3])
            N = np.tile(R[i, 1, :3], (m, 1))
        elif kind == "s":
            d = rng.normal(size=(m, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
            P, N = S[i, 0, :3] + S[i, 0, 3] * d, d
        else:
            flip = u + v > 1
            u, v = np.where(flip, 1 - u, u), np.where(flip, 1 - v, v)
            P = T[i, 0, :3] + np.outer(u, T[i, 1, :3]) + np.outer(v, T[i, 2, :3])
            n = np.cross(T[i, 1, :3], T[i, 2, :3]); n = n / (np.linalg.norm(n) or 1.0)
            N = np.tile(n, (m, 1)) * np.where(rng.random(m) < 0.5, 1.0, -1.0)[:, None]
        pts.append(P); nrm.append(N)
    P, N = np.concatenate(pts), np.concatenate(nrm)
    kinds = rec[:, 3, 2].astype(np.float32).view(np.int32) >> 28
    for j in range(n_lights):
        L = p.lights[j, :3]
        d = L - P
        dist = np.linalg.norm(d, axis=1)
        ok = dist > 1e-3
        d = d / np.where(ok, dist, 1.0)[:, None]
        ok &= (d * N).sum(1) > 0                       # zero-payload shadow rays are never queued
        o = P + 1e-3 * N
        counts = {}
        for k in range(rec.shape[0]):
            q0, q1, q2, q3 = rec[k]
            dn = d @ q0[:3]
            with np.errstate(divide="ignore", invalid="ignore"):
                t = (q0[3] - o @ q0[:3]) / dn
            X = o + t[:, None] * d
            u, v = X @ q1[:3] + q1[3], X @ q2[:3] + q2[3]
            inside = (u >= 0) & (v >= 0) & ((u + v <= 1) if kinds[k] == 1 else ((u <= q3[0]) & (v <= q3[1])))
            counts[k] = int((ok & inside & (np.abs(dn) > 1e-6) & (t > 1e-3) & (t < 1e6)).sum())
        for i in range(p.n_sphere):
            oc = o - S[i, 0, :3]
            b = (oc * d).sum(1)
            disc = b * b - ((oc * oc).sum(1) - S[i, 1, 0])
            sq = np.sqrt(np.maximum(disc, 0))
            hit = (disc > 0) & (((-b - sq) > 1e-3) | ((-b + sq) > 1e-3))
            counts[64 + i] = int((ok & hit).sum())
        best = max(counts, key=counts.get)
        code = best
        if generic:
            code = 128 + (best if best < 64 else p.n_rect + (best - 64))
        hints[j] = code if counts[best] > 0 else -1
    return hints
