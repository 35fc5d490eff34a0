"""Seeded random scenes built through ANY implementation of the reference object model.

``make_scene(api, seed)`` takes a namespace with ``Vec3, Material, Texture, Plane, Sphere, Triangle, Scene,
Camera`` (``b200rt.scene_api`` or the reference's ``core.*`` modules) so the *same* scene can be driven
through the real reference (container only), the oracle and the CUDA kernels.  Scenes mix every
primitive type and material branch the renderers have: skewed rectangles (non-unit, non-orthogonal
``u_dir``/``v_dir`` so the float32 axis normalisation matters), spheres with non-integer radii, glass /
mirror / diffuse / semi-reflective materials, textured and untextured triangles with arbitrary UVs.
"""
from __future__ import annotations

import random
import types

import numpy as np


def local_api():
    from b200rt import scene_api as A
    return A


def reference_api():
    """The reference's own classes (needs /root/reference on sys.path with cwd-independent imports)."""
    import importlib
    ns = types.SimpleNamespace()
    m, g, mt, sc, cam = (importlib.import_module(n) for n in
                         ("core.math", "core.geometry", "core.material", "core.scene", "core.camera"))
    ns.Vec3, ns.Plane, ns.Sphere, ns.Triangle = m.Vec3, g.Plane, g.Sphere, g.Triangle
    ns.Material, ns.Texture, ns.Scene, ns.Camera = mt.Material, mt.Texture, sc.Scene, cam.Camera
    return ns


def _texture(api, rng, idx, size):
    h, w = size
    px = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    if hasattr(api.Texture, "from_array"):
        return api.Texture.from_array(px, f"textures/rand{idx}.png")
    t = api.Texture.__new__(api.Texture)          # the reference's Texture only loads from a file
    t.path, t.pixels, t.width, t.height = f"textures/rand{idx}.png", px, w, h
    return t


def make_scene(api, seed: int, n_rect=4, n_sphere=4, n_tri=10, n_lights=5, build_bvh=True, with_textures=True):
    rng = np.random.default_rng(seed)
    V = api.Vec3
    textures = [_texture(api, rng, i, s) for i, s in enumerate(((7, 5), (16, 16), (3, 11)))] if with_textures else []

    def material(allow_glass=True, tex=None):
        kind = rng.integers(0, 4 if allow_glass else 3)
        col = V(*rng.uniform(0.1, 1.0, 3))
        if kind == 0:
            return api.Material(col, diffuse=float(rng.uniform(0.3, 0.95)), specular=float(rng.uniform(0, 0.6)), texture=tex)
        if kind == 1:
            return api.Material(col, diffuse=0.05, specular=0.95, reflective=float(rng.uniform(0.55, 0.98)), texture=tex)
        if kind == 2:
            return api.Material(col, diffuse=float(rng.uniform(0.2, 0.8)), specular=0.3,
                                reflective=float(rng.uniform(0.02, 0.45)), texture=tex)
        return api.Material(col, diffuse=0.1, specular=0.9, reflective=0.1, refractive=float(rng.uniform(0.55, 0.9)),
                            ior=float(rng.uniform(1.2, 2.0)))

    scene = api.Scene()
    for _ in range(n_rect):
        anchor = V(*rng.uniform(-8, 8, 3))
        u = rng.normal(size=3); u /= np.linalg.norm(u)
        v = rng.normal(size=3); v -= u * (v @ u); v /= np.linalg.norm(v)
        n = np.cross(u, v)
        ul, vl = float(rng.uniform(2, 9)), float(rng.uniform(2, 9))
        su, sv = float(rng.uniform(0.5, 3.0)), float(rng.uniform(0.5, 3.0))      # non-unit direction vectors
        scene.add_object(api.Plane(anchor, V(*n), V(*(u * su)), V(*(v * sv)), ul, vl, material(allow_glass=False)))
    for _ in range(n_sphere):
        scene.add_object(api.Sphere(V(*rng.uniform(-7, 7, 3)), float(rng.uniform(0.6, 2.7)), material()))
    for k in range(n_tri):
        c = rng.uniform(-7, 7, 3)
        p = [V(*(c + rng.normal(scale=2.5, size=3))) for _ in range(3)]
        tex = textures[k % len(textures)] if (textures and k % 2 == 0) else None
        uv = [np.array(rng.uniform(-0.2, 1.2, 2)) for _ in range(3)] if k % 3 else [None, None, None]
        if uv[0] is None and hasattr(api, "__name__") and False:
            pass
        scene.add_object(api.Triangle(p[0], p[1], p[2], uv[0], uv[1], uv[2], material(allow_glass=False, tex=tex)))
    for _ in range(n_lights):
        scene.add_light_sample(V(*rng.uniform(-9, 9, 3)))
    scene.light_color = V(*rng.uniform(0.4, 1.0, 3))
    scene.ambient = V(*rng.uniform(0.1, 0.6, 3))
    if build_bvh:
        random.seed(seed)
        scene.build_bvh()
    camera = api.Camera(V(0.3, 0.7, 24.0), V(0.1, -0.2, 0.0), V(0, 1, 0), 48.0, 4 / 3)
    return scene, camera
