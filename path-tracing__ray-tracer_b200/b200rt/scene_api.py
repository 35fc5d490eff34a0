"""Host-side object model, API-compatible with the reference's ``core`` package.

The B200 renderers consume the *reference's own* objects when they run inside a
reference checkout (they only duck-type on attribute names).  This module is the
stand-alone mirror of that object model so that scenes can be built on machines
where the reference is not present (the GPU box, the tests, ``bench.py``).

Interface mirrored (attribute names, constructor argument order and numeric
semantics; all arithmetic is float64 Python floats exactly like the reference):

* ``Vec3`` / ``Ray`` / ``AABB``            <- reference ``core/math.py:4-117``
* ``Texture`` / ``Material`` / ``HitRecord`` <- ``core/material.py:6-58``
* ``Plane`` / ``Sphere`` / ``Triangle``      <- ``core/geometry.py:18-174``
* ``BVHNode``                                <- ``core/acceleration.py:7-43``
* ``Camera``                                 <- ``core/camera.py:5-31``
* ``RenderSettings`` / ``Scene`` / ``create_area_light`` <- ``core/scene.py:19-80``

Nothing here runs on the hot path: the packer (``packer.py``) flattens these
objects into SoA arrays once per scene and the CUDA library does the rest.
"""
from __future__ import annotations

import math
import random
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

__all__ = [
    "Vec3", "Ray", "AABB", "Texture", "Material", "HitRecord", "Hittable",
    "Plane", "Sphere", "Triangle", "BVHNode", "Camera", "RenderSettings",
    "Scene", "create_area_light",
]


# --------------------------------------------------------------------------- math
class Vec3:
    """3-vector of Python floats (reference ``core/math.py:4-71``)."""

    __slots__ = ("x", "y", "z")

    def __init__(self, x=0.0, y=0.0, z=0.0):
        self.x, self.y, self.z = float(x), float(y), float(z)

    # element access helpers (not in the reference; used by the packer)
    def astuple(self):
        return (self.x, self.y, self.z)

    def __iter__(self):
        return iter((self.x, self.y, self.z))

    def __add__(self, o):
        return Vec3(self.x + o.x, self.y + o.y, self.z + o.z)

    def __sub__(self, o):
        return Vec3(self.x - o.x, self.y - o.y, self.z - o.z)

    def __mul__(self, k):
        if isinstance(k, Vec3) or (hasattr(k, "x") and hasattr(k, "z")):
            return Vec3(self.x * k.x, self.y * k.y, self.z * k.z)   # Hadamard
        return Vec3(self.x * k, self.y * k, self.z * k)

    __rmul__ = __mul__

    def __truediv__(self, k):
        return Vec3(self.x / k, self.y / k, self.z / k)

    def __neg__(self):
        return Vec3(-self.x, -self.y, -self.z)

    def dot(self, o):
        return self.x * o.x + self.y * o.y + self.z * o.z

    def cross(self, o):
        return Vec3(self.y * o.z - self.z * o.y,
                    self.z * o.x - self.x * o.z,
                    self.x * o.y - self.y * o.x)

    def length(self):
        return math.sqrt(self.x * self.x + self.y * self.y + self.z * self.z)

    def normalize(self):
        n = self.length()
        return Vec3(0, 0, 0) if n == 0 else self / n

    def reflect(self, normal):
        return self - normal * (2 * self.dot(normal))

    def refract(self, normal, ni_over_nt):
        unit = self.normalize()
        dt = unit.dot(normal)
        disc = 1.0 - ni_over_nt * ni_over_nt * (1 - dt * dt)
        if disc > 0:
            return True, (unit - normal * dt) * ni_over_nt - normal * math.sqrt(disc)
        return False, None

    def to_np(self):
        return np.array([self.x, self.y, self.z], dtype=np.float32)

    def __repr__(self):
        return f"Vec3({self.x:.3f}, {self.y:.3f}, {self.z:.3f})"


class Ray:
    """Origin + *normalised* direction (reference ``core/math.py:76-82``)."""

    __slots__ = ("origin", "direction")

    def __init__(self, origin: Vec3, direction: Vec3):
        self.origin = origin
        self.direction = direction.normalize()

    def point_at_parameter(self, t):
        return self.origin + self.direction * t


class AABB:
    """Axis-aligned box with the reference's inclusive slab test (``core/math.py:85-117``)."""

    __slots__ = ("min", "max")

    def __init__(self, min_pt: Vec3, max_pt: Vec3):
        self.min, self.max = min_pt, max_pt

    @staticmethod
    def surrounding_box(a: "AABB", b: "AABB") -> "AABB":
        lo = Vec3(min(a.min.x, b.min.x), min(a.min.y, b.min.y), min(a.min.z, b.min.z))
        hi = Vec3(max(a.max.x, b.max.x), max(a.max.y, b.max.y), max(a.max.z, b.max.z))
        return AABB(lo, hi)

    def hit(self, ray: Ray, t_min: float, t_max: float) -> bool:
        o, d = ray.origin.astuple(), ray.direction.astuple()
        lo, hi = self.min.astuple(), self.max.astuple()
        for axis in range(3):
            inv = 1.0 / d[axis]            # ZeroDivisionError on axis-parallel rays, like the reference
            near = (lo[axis] - o[axis]) * inv
            far = (hi[axis] - o[axis]) * inv
            if inv < 0.0:
                near, far = far, near
            if near > t_min:
                t_min = near
            if far < t_max:
                t_max = far
            if t_max < t_min:
                return False
        return True


# ----------------------------------------------------------------------- materials
class Texture:
    """RGB8 image sampled nearest-texel with a V flip (``core/material.py:6-21``).

    ``Texture(path)`` decodes with PIL like the reference; ``Texture.from_array``
    is the side door used for synthetic textures (no file needed).
    """

    def __init__(self, path: str):
        from PIL import Image
        self.path = path
        img = Image.open(path).convert("RGB")
        self.width, self.height = img.size
        self.pixels = np.array(img)

    @classmethod
    def from_array(cls, pixels: np.ndarray, path: str) -> "Texture":
        tex = cls.__new__(cls)
        tex.path = path
        tex.pixels = np.ascontiguousarray(pixels, dtype=np.uint8)
        tex.height, tex.width = tex.pixels.shape[:2]
        return tex

    def sample(self, u: float, v: float) -> Vec3:
        iu = int(max(0, min(self.width - 1, u * (self.width - 1))))
        iv = int(max(0, min(self.height - 1, (1.0 - v) * (self.height - 1))))
        r, g, b = self.pixels[iv, iu]
        return Vec3(r / 255.0, g / 255.0, b / 255.0)


class Material:
    """Phong-ish material record (``core/material.py:24-48``)."""

    def __init__(self, color: Vec3 = None, diffuse=1.0, specular=0.0, reflective=0.0,
                 refractive=0.0, ior=1.0, texture: Optional[Texture] = None):
        self.color = Vec3(1, 1, 1) if color is None else color
        self.diffuse = diffuse
        self.specular = specular
        self.reflective = reflective
        self.refractive = refractive
        self.ior = ior
        self.texture = texture


class HitRecord:
    """Mutable hit record (``core/material.py:51-58``)."""

    __slots__ = ("t", "point", "normal", "material", "u", "v")

    def __init__(self):
        self.t = float("inf")
        self.point = None
        self.normal = None
        self.material = None
        self.u = 0.0
        self.v = 0.0


# ------------------------------------------------------------------------ geometry
class Hittable:
    def hit(self, ray: Ray, t_min: float, t_max: float, rec: HitRecord) -> bool:  # pragma: no cover
        raise NotImplementedError

    def bounding_box(self) -> AABB:  # pragma: no cover
        raise NotImplementedError


def _box_of(points: Sequence[Vec3]) -> AABB:
    xs, ys, zs = zip(*(p.astuple() for p in points))
    return AABB(Vec3(min(xs), min(ys), min(zs)), Vec3(max(xs), max(ys), max(zs)))


class Plane(Hittable):
    """Finite rectangle ``anchor + a*u_unit + b*v_unit`` (``core/geometry.py:18-75``).

    ``v_unit`` is derived as ``normal x u_unit`` (NOT from ``v_dir``) exactly like the
    reference's CPU path; ``v_dir`` is kept because the reference's GPU packers ship it.
    """

    def __init__(self, anchor: Vec3, normal: Vec3, u_dir: Vec3, v_dir: Vec3,
                 u_len: float, v_len: float, material: Material):
        self.anchor = anchor
        self.normal = normal.normalize()
        self.u_dir, self.v_dir = u_dir, v_dir
        self.u_len, self.v_len = u_len, v_len
        self.material = material
        self.u_unit = u_dir.normalize()
        self.v_unit = self.normal.cross(self.u_unit).normalize()
        self.u_extent, self.v_extent = u_len, v_len
        du, dv = self.u_unit * u_len, self.v_unit * v_len
        self.box = _box_of([anchor, anchor + du, anchor + dv, anchor + du + dv])

    def hit(self, ray, t_min, t_max, rec):
        denom = self.normal.dot(ray.direction)
        if abs(denom) < 1e-6:
            return False
        t = (self.anchor - ray.origin).dot(self.normal) / denom
        if t < t_min or t > t_max:          # inclusive range: the reference accepts t == t_max here
            return False
        p = ray.point_at_parameter(t)
        rel = p - self.anchor
        a, b = rel.dot(self.u_unit), rel.dot(self.v_unit)
        if a < 0 or a > self.u_extent or b < 0 or b > self.v_extent:
            return False
        rec.t, rec.point, rec.normal, rec.material = t, p, self.normal, self.material
        rec.u, rec.v = a / self.u_extent, b / self.v_extent
        return True

    def bounding_box(self):
        return self.box


class Sphere(Hittable):
    """Sphere with the half-b quadratic, near root first (``core/geometry.py:78-114``)."""

    def __init__(self, center: Vec3, radius: float, material: Material):
        self.center, self.radius, self.material = center, radius, material
        r = Vec3(radius, radius, radius)
        self.box = AABB(center - r, center + r)

    def hit(self, ray, t_min, t_max, rec):
        oc = ray.origin - self.center
        a = ray.direction.dot(ray.direction)
        b = oc.dot(ray.direction)
        c = oc.dot(oc) - self.radius * self.radius
        disc = b * b - a * c
        if disc > 0:
            root = math.sqrt(disc)
            for t in ((-b - root) / a, (-b + root) / a):
                if t_min < t < t_max:
                    rec.t = t
                    rec.point = ray.point_at_parameter(t)
                    rec.normal = (rec.point - self.center) / self.radius
                    rec.material = self.material
                    rec.u = rec.v = 0.0
                    return True
        return False

    def bounding_box(self):
        return self.box


class Triangle(Hittable):
    """Moeller-Trumbore triangle with per-vertex UVs (``core/geometry.py:117-174``)."""

    def __init__(self, v0: Vec3, v1: Vec3, v2: Vec3, uv0=None, uv1=None, uv2=None,
                 material: Material = None):
        self.v0, self.v1, self.v2 = v0, v1, v2
        self.uv0, self.uv1, self.uv2 = uv0, uv1, uv2
        self.material = material
        self.normal = (v1 - v0).cross(v2 - v0).normalize()
        self.box = _box_of([v0, v1, v2])

    def hit(self, ray, t_min, t_max, rec):
        e1, e2 = self.v1 - self.v0, self.v2 - self.v0
        h = ray.direction.cross(e2)
        det = e1.dot(h)
        if abs(det) < 1e-6:
            return False
        f = 1.0 / det
        s = ray.origin - self.v0
        u = f * s.dot(h)
        if u < 0.0 or u > 1.0:
            return False
        q = s.cross(e1)
        v = f * ray.direction.dot(q)
        if v < 0.0 or u + v > 1.0:
            return False
        t = f * e2.dot(q)
        if not (t_min < t < t_max):
            return False
        rec.t = t
        rec.point = ray.point_at_parameter(t)
        rec.normal = self.normal if self.normal.dot(ray.direction) < 0 else -self.normal
        rec.material = self.material
        if self.uv0 is not None:
            w = 1 - u - v
            rec.u = u * self.uv1[0] + v * self.uv2[0] + w * self.uv0[0]
            rec.v = u * self.uv1[1] + v * self.uv2[1] + w * self.uv0[1]
        else:
            rec.u = rec.v = 0.0
        return True

    def bounding_box(self):
        return self.box


class BVHNode(Hittable):
    """Median-split BVH on a random axis (``core/acceleration.py:7-43``).

    Kept call-for-call compatible with the reference so that ``random.seed(k)``
    followed by ``Scene.build_bvh()`` yields the *same* permutation of
    ``scene.objects`` and the same tree: one ``random.randint(0, 2)`` per node in
    pre-order, a stable sort on the bbox minimum along that axis, and the split at
    ``start + span // 2``.  The B200 path never traverses this tree (it builds an
    LBVH on the device); it exists for the CPU mirror and for the oracle, which
    replays the reference's CPU renderer through exactly this structure.
    """

    def __init__(self, objects: list, start: int, end: int):
        axis = "xyz"[random.randint(0, 2)]
        objects[start:end] = sorted(objects[start:end],
                                    key=lambda o: getattr(o.bounding_box().min, axis))
        span = end - start
        if span == 1:
            self.left = self.right = objects[start]
        elif span == 2:
            self.left, self.right = objects[start], objects[start + 1]
        else:
            mid = start + span // 2
            self.left = BVHNode(objects, start, mid)
            self.right = BVHNode(objects, mid, end)
        self.box = AABB.surrounding_box(self.left.bounding_box(), self.right.bounding_box())

    def hit(self, ray, t_min, t_max, rec):
        if not self.box.hit(ray, t_min, t_max):
            return False
        got_left = self.left.hit(ray, t_min, t_max, rec)
        if got_left:
            t_max = rec.t
        got_right = self.right.hit(ray, t_min, t_max, rec)
        return got_left or got_right

    def bounding_box(self):
        return self.box


# -------------------------------------------------------------------------- camera
class Camera:
    """Pinhole camera (``core/camera.py:5-31``)."""

    def __init__(self, lookfrom: Vec3, lookat: Vec3, vup: Vec3, vfov: float, aspect: float):
        self.origin = lookfrom
        half_h = math.tan(math.radians(vfov) / 2)
        half_w = aspect * half_h
        w = (lookfrom - lookat).normalize()
        u = vup.cross(w).normalize()
        v = w.cross(u)
        self.lower_left_corner = self.origin - u * half_w - v * half_h - w
        self.horizontal = u * (2 * half_w)
        self.vertical = v * (2 * half_h)

    def get_ray(self, s: float, t: float) -> Ray:
        return Ray(self.origin,
                   self.lower_left_corner + self.horizontal * s + self.vertical * t - self.origin)


# --------------------------------------------------------------------------- scene
@dataclass
class RenderSettings:
    """``core/scene.py:19-24``."""
    width: int = 800
    height: int = 600
    samples_per_pixel: int = 9
    max_depth: int = 4


class Scene:
    """Object/light container (``core/scene.py:27-64``)."""

    def __init__(self):
        self.objects: List[Hittable] = []
        self.bvh_root = None
        self.lights: List[Vec3] = []
        self.light_color = Vec3(1.0, 1.0, 1.0)
        self.ambient = Vec3(0.5, 0.5, 0.5)

    def add_object(self, obj: Hittable):
        self.objects.append(obj)

    def add_mesh(self, vertices, faces, material, uvs=None, spatial_order: bool = False):
        """Bulk triangle ingestion: one ``packer.TriangleMesh`` entry in ``objects`` that the packer
        expands, vectorised, into ``len(faces)`` triangles — the object-per-triangle API (``add_object(Triangle(..))``,
        ``core/scene.py:35-36``) cannot express million-triangle scenes in reasonable time.  ``spatial_order`` lists the
        faces along a Morton curve first (``TriangleMesh.spatially_sorted``).  Returns the mesh."""
        from .packer import TriangleMesh
        mesh = TriangleMesh(vertices, faces, material, uvs)
        if spatial_order:
            mesh = mesh.spatially_sorted()
        self.objects.append(mesh)
        return mesh

    def add_obj(self, path: str, material, scale: float = 1.0, translate=(0.0, 0.0, 0.0), spatial_order: bool = False):
        """``add_mesh`` from a Wavefront OBJ file (``packer.load_obj``)."""
        from .packer import load_obj
        mesh = load_obj(path, material, scale, translate)
        if spatial_order:
            mesh = mesh.spatially_sorted()
        self.objects.append(mesh)
        return mesh

    def build_bvh(self):
        # the host-side BVH mirrors the reference's (it shuffles ``objects``); bulk meshes only exist on the device
        prims = [o for o in self.objects if hasattr(o, "bounding_box")]
        if prims and len(prims) == len(self.objects):
            self.bvh_root = BVHNode(self.objects, 0, len(self.objects))

    def add_light_sample(self, pos: Vec3):
        self.lights.append(pos)

    def hit(self, ray: Ray, t_min: float, t_max: float, rec: HitRecord) -> bool:
        if self.bvh_root:
            return self.bvh_root.hit(ray, t_min, t_max, rec)
        found, closest, tmp = False, t_max, HitRecord()
        for obj in self.objects:
            if obj.hit(ray, t_min, closest, tmp):
                found, closest = True, tmp.t
                rec.t, rec.point, rec.normal = tmp.t, tmp.point, tmp.normal
                rec.material, rec.u, rec.v = tmp.material, tmp.u, tmp.v
        return found


def create_area_light(scene: Scene, center: Vec3, u_vec: Vec3, v_vec: Vec3,
                      u_size: float, v_size: float, n_u: int, n_v: int):
    """n_u x n_v point samples on a rectangle (``core/scene.py:67-80``)."""
    half_u = u_vec.normalize() * (u_size / 2.0)
    half_v = v_vec.normalize() * (v_size / 2.0)
    for i in range(n_u):
        for j in range(n_v):
            ru = (i + 0.5) / n_u - 0.5
            rv = (j + 0.5) / n_v - 0.5
            scene.add_light_sample(center + half_u * (2 * ru) + half_v * (2 * rv))
