"""Cornell-box scene builder, value-for-value compatible with the reference's
``scene_builders/custom_scene_builder.py:10-490`` (``CustomSceneBuilder``).

``CustomSceneBuilder().build_scene()`` returns a ``Scene`` holding, in this order
before the BVH shuffle: 5 wall ``Plane``s, 2 textured cubes (24 ``Triangle``s), 3
``Sphere``s (glass, chrome, glass), a textured canvas (2 ``Triangle``s) and a 4x4
grid of light samples; ``create_camera(aspect)`` returns the (0,0,50) -> origin,
49.5 degree pinhole.  Every coordinate is produced by the same float64 expression
order as the reference so the two builders agree bit-for-bit
(``tests/test_scene_api.py`` checks that when the reference is present).

Textures: the reference opens ``textures/<name>.jpg`` relative to the working
directory (``custom_scene_builder.py:77-86``).  This builder does the same when
those files exist (``texture_dir``); otherwise it substitutes deterministic
synthetic images of the *same dimensions* (``synthetic_texture``), keeping the
52 070 958-byte gather footprint of the real scene.  The texture ``path`` strings
are identical either way because the reference derives texture ids from the
sorted path strings (``cuda_path_tracer.py:830-832``).
"""
from __future__ import annotations

import math
import os
from typing import Dict, Optional

import numpy as np

from .scene_api import (Camera, Material, Plane, Scene, Sphere, Texture, Triangle, Vec3,
                        create_area_light)

# (width, height) of the reference's JPEGs, measured from /root/reference/textures.
TEXTURE_SIZES = {
    "blue": (1318, 1319), "green": (1317, 1316), "meinsf": (2978, 2393), "orange": (1269, 1268),
    "red": (1315, 1314), "white": (1296, 1296), "yellow": (1320, 1320),
}
_STICKER_RGB = {
    "blue": (20, 60, 200), "green": (10, 150, 40), "orange": (250, 110, 10), "red": (200, 20, 25),
    "white": (235, 235, 230), "yellow": (250, 225, 20),
}


def synthetic_texture(name: str) -> np.ndarray:
    """Deterministic stand-in for ``textures/<name>.jpg``: uint8 [H, W, 3].

    Cube faces get a 3x3 sticker layout with dark gaps plus a fine integer-hash
    grain (so neighbouring texels differ and a wrong texel index is detectable);
    the canvas gets a smooth two-frequency colour field with the same grain.
    Pure integer/array arithmetic — no RNG, identical on every machine.
    """
    w, h = TEXTURE_SIZES[name]
    yy, xx = np.meshgrid(np.arange(h, dtype=np.int64), np.arange(w, dtype=np.int64), indexing="ij")
    grain = (((xx * 73856093) ^ (yy * 19349663) ^ (len(name) * 83492791)) >> 7) & 31   # 0..31
    img = np.empty((h, w, 3), dtype=np.int64)
    if name == "meinsf":
        fx, fy = xx / w, yy / h
        img[..., 0] = 128 + 100 * np.sin(6.0 * fx + 2.0 * fy)
        img[..., 1] = 128 + 100 * np.sin(4.0 * fy - 3.0 * fx + 1.0)
        img[..., 2] = 128 + 100 * np.cos(5.0 * fx * fy + 0.5)
        img += grain[..., None] - 16
    else:
        cell_x, cell_y = (xx * 3) // w, (yy * 3) // h
        # sticker interior = 8%..92% of every cell
        ix, iy = (xx * 3 - cell_x * w) / w, (yy * 3 - cell_y * h) / h
        interior = (ix > 0.08) & (ix < 0.92) & (iy > 0.08) & (iy < 0.92)
        base = np.array(_STICKER_RGB[name], dtype=np.int64)
        shade = 1.0 - 0.04 * ((cell_x + 2 * cell_y) % 3)
        for c in range(3):
            img[..., c] = np.where(interior, base[c] * shade, 18)
        img += grain[..., None] - 16
    return np.clip(img, 0, 255).astype(np.uint8)


def _load_texture(name: str, texture_dir: Optional[str]) -> Texture:
    rel = f"textures/{name}.jpg"
    if texture_dir is not None:
        real = os.path.join(texture_dir, f"{name}.jpg")
        if os.path.isfile(real):
            tex = Texture(real)
            tex.path = rel              # ids come from the sorted *relative* path strings
            return tex
    return Texture.from_array(synthetic_texture(name), rel)


class CustomSceneBuilder:
    """Same public surface as the reference builder: ``build_scene()``, ``create_camera(aspect)``.

    ``texture_dir=None`` looks for ``./textures`` (the reference's convention) and falls
    back to synthetic textures; pass ``texture_dir=False`` to force synthetic ones.
    """

    def __init__(self, texture_dir=None):
        self.box_size = 30.0
        self.foam_thickness = 0.5
        self.cube_size = 5.6
        self.canvas_width = 27.5
        self.canvas_height = 22.0
        self.canvas_depth = 1.5
        self.canvas_angle = 112.0
        self.light_size = 3.0
        if texture_dir is None:
            texture_dir = "textures" if os.path.isdir("textures") else False
        self.texture_dir = texture_dir or None

    # ---------------------------------------------------------------- public
    def build_scene(self) -> Scene:
        scene = Scene()
        mats = self._create_materials()
        self._create_walls(scene, mats)
        self._create_rubiks_cubes(scene, mats)
        self._create_spheres(scene, mats)
        self._create_canvas(scene, mats)
        self._create_lighting(scene)
        scene.build_bvh()                          # permutes scene.objects (random axis per node)
        scene.light_color = Vec3(0.7, 0.7, 0.7)
        scene.ambient = Vec3(0.5, 0.5, 0.5)
        return scene

    def create_camera(self, aspect_ratio: float = 4.0 / 3.0) -> Camera:
        return Camera(Vec3(0, 0, 50.0), Vec3(0, 0, 0), Vec3(0, 1, 0), 49.5, aspect_ratio)

    # --------------------------------------------------------------- pieces
    def _create_materials(self) -> Dict[str, Material]:
        tex = {n: _load_texture(n, self.texture_dir) for n in
               ("blue", "green", "orange", "red", "white", "yellow")}
        canvas_tex = _load_texture("meinsf", self.texture_dir)
        wall = dict(diffuse=0.8, specular=0.1)
        cube = dict(diffuse=0.7, specular=0.4, reflective=0.0)
        return {
            "floor": Material(color=Vec3(0.9, 0.9, 0.9), **wall),
            "back": Material(color=Vec3(0.9, 0.9, 0.9), **wall),
            "left": Material(color=Vec3(255 / 255, 105 / 255, 180 / 255), **wall),
            "right": Material(color=Vec3(52 / 255, 157 / 255, 204 / 255), **wall),
            "ceiling": Material(color=Vec3(0.9, 0.9, 0.9), **wall),
            "cube_blue": Material(color=Vec3(0.0, 0.2, 0.8), texture=tex["blue"], **cube),
            "cube_green": Material(color=Vec3(0.0, 0.6, 0.0), texture=tex["green"], **cube),
            "cube_orange": Material(color=Vec3(1.0, 0.4, 0.0), texture=tex["orange"], **cube),
            "cube_red": Material(color=Vec3(0.8, 0.0, 0.0), texture=tex["red"], **cube),
            "cube_white": Material(color=Vec3(0.9, 0.9, 0.9), texture=tex["white"], **cube),
            "cube_yellow": Material(color=Vec3(1.0, 0.9, 0.0), texture=tex["yellow"], **cube),
            "canvas": Material(color=Vec3(0.9, 0.8, 0.6), diffuse=0.9, specular=0.1, texture=canvas_tex),
            "sphere_metal": Material(color=Vec3(0.9, 0.9, 0.9), diffuse=0.05, specular=0.95,
                                     reflective=0.95),
            "glass": Material(color=Vec3(0.95, 0.95, 0.95), diffuse=0.1, specular=0.9,
                              reflective=0.1, refractive=0.85, ior=1.5),
        }

    def _create_wall_materials(self) -> Dict[str, Material]:
        wall = dict(diffuse=0.8, specular=0.1)
        return {
            "floor": Material(color=Vec3(0.9, 0.9, 0.9), **wall),
            "back": Material(color=Vec3(0.9, 0.9, 0.9), **wall),
            "left": Material(color=Vec3(255 / 255, 105 / 255, 180 / 255), **wall),
            "right": Material(color=Vec3(52 / 255, 157 / 255, 204 / 255), **wall),
            "ceiling": Material(color=Vec3(0.9, 0.9, 0.9), **wall),
        }

    def _create_walls(self, scene: Scene, mats) -> None:
        s, h = self.box_size, self.box_size / 2.0
        # (material, anchor, normal, u_dir, v_dir)
        walls = [
            ("floor",   Vec3(-h, -h, h),  Vec3(0, 1, 0),  Vec3(s, 0, 0),  Vec3(0, 0, -s)),
            ("back",    Vec3(-h, -h, -h), Vec3(0, 0, 1),  Vec3(s, 0, 0),  Vec3(0, s, 0)),
            ("left",    Vec3(-h, -h, h),  Vec3(1, 0, 0),  Vec3(0, 0, -s), Vec3(0, s, 0)),
            ("right",   Vec3(h, -h, -h),  Vec3(-1, 0, 0), Vec3(0, 0, s),  Vec3(0, s, 0)),
            ("ceiling", Vec3(-h, h, -h),  Vec3(0, -1, 0), Vec3(s, 0, 0),  Vec3(0, 0, s)),
        ]
        for name, anchor, normal, u_dir, v_dir in walls:
            scene.add_object(Plane(anchor=anchor, normal=normal, u_dir=u_dir, v_dir=v_dir,
                                   u_len=s, v_len=s, material=mats[name]))

    def _create_rubiks_cubes(self, scene: Scene, mats) -> None:
        half = self.cube_size / 2.0
        floor_y = -self.box_size / 2.0
        self._create_single_cube(scene, mats, Vec3(0, floor_y + half, 0), 225.0)
        self._create_single_cube(scene, mats, Vec3(0, floor_y + half + self.cube_size, 0), 0.0)

    def _create_single_cube(self, scene: Scene, mats, center: Vec3, rotation_y: float) -> None:
        k = self.cube_size / 2.0
        corners = [(-k, -k, k), (k, -k, k), (k, k, k), (-k, k, k),
                   (-k, -k, -k), (k, -k, -k), (k, k, -k), (-k, k, -k)]
        ang = math.radians(rotation_y)
        c, s = math.cos(ang), math.sin(ang)
        world = [center + Vec3(x * c - z * s, y, x * s + z * c) for x, y, z in corners]
        uv = [np.array(p) for p in ((0, 0), (1, 0), (1, 1), (0, 1))]
        faces = [((0, 1, 2, 3), "cube_red"), ((1, 5, 6, 2), "cube_blue"), ((3, 2, 6, 7), "cube_yellow"),
                 ((4, 5, 1, 0), "cube_white"), ((4, 0, 3, 7), "cube_orange"), ((5, 4, 7, 6), "cube_green")]
        for (a, b, cc, d), mat in faces:
            scene.add_object(Triangle(world[a], world[b], world[cc], uv[0], uv[1], uv[2], mats[mat]))
            scene.add_object(Triangle(world[a], world[cc], world[d], uv[0], uv[2], uv[3], mats[mat]))

    def _create_spheres(self, scene: Scene, mats) -> None:
        floor_y = -self.box_size / 2.0
        q = self.box_size / 4
        scene.add_object(Sphere(center=Vec3(q, floor_y + 3, q), radius=3, material=mats["glass"]))
        scene.add_object(Sphere(center=Vec3(-q, floor_y + 3, q), radius=3, material=mats["sphere_metal"]))
        cube2_center_y = floor_y + self.cube_size / 2.0 + self.cube_size
        cube2_top_y = cube2_center_y + self.cube_size / 2.0
        scene.add_object(Sphere(center=Vec3(0, cube2_top_y + 3.0, 0), radius=3.0, material=mats["glass"]))

    def _create_canvas(self, scene: Scene, mats) -> None:
        back_z = -self.box_size / 2.0
        bottom_y = -self.box_size / 2.0 + 0.5
        ang = math.radians(self.canvas_angle)
        half_w = self.canvas_width / 2.0
        bottom_z = back_z + 6.5 * self.canvas_depth
        top_z = bottom_z + self.canvas_height * math.cos(ang)
        top_y = bottom_y + self.canvas_height * math.sin(ang)
        bl, br = Vec3(-half_w, bottom_y, bottom_z), Vec3(half_w, bottom_y, bottom_z)
        tl, tr = Vec3(-half_w, top_y, top_z), Vec3(half_w, top_y, top_z)
        uv_bl, uv_br, uv_tl, uv_tr = (np.array(p) for p in ((0, 0), (1, 0), (0, 1), (1, 1)))
        scene.add_object(Triangle(bl, br, tr, uv_bl, uv_br, uv_tr, mats["canvas"]))
        scene.add_object(Triangle(bl, tr, tl, uv_bl, uv_tr, uv_tl, mats["canvas"]))

    def _create_lighting(self, scene: Scene) -> None:
        create_area_light(scene, center=Vec3(0, self.box_size / 2 - 1, 0),
                          u_vec=Vec3(1, 0, 0), v_vec=Vec3(0, 0, 1),
                          u_size=self.light_size, v_size=self.light_size, n_u=4, n_v=4)
