"""Synthetic scenes for the BASELINE configurations that the reference cannot build itself.

``heightfield_scene`` = BASELINE config 4: the Cornell walls and light samples unchanged,
cubes/spheres/canvas replaced by a 1001 x 501 height-field over x, z in [-14, 14] -> 1 000 000 triangles,
y = -15 + 4 + 3 sin(0.7 x) cos(0.9 z) + 0.5 N(0,1) with numpy.random.default_rng(1234); material colour
(0.8, 0.8, 0.8), diffuse 0.8, no texture.  The object-per-triangle API (``Scene.add_object(Triangle(...))``,
core/scene.py:35-36) cannot express this in reasonable time, so the mesh enters through the
``packer.TriangleMesh`` side door.
"""
from __future__ import annotations

import numpy as np

from .cornell import CustomSceneBuilder
from .packer import TriangleMesh
from .scene_api import Material, Scene, Vec3


def heightfield_mesh(nx: int = 1001, nz: int = 501, seed: int = 1234) -> TriangleMesh:
    x = np.linspace(-14.0, 14.0, nx)
    z = np.linspace(-14.0, 14.0, nz)
    X, Z = np.meshgrid(x, z, indexing="ij")
    rng = np.random.default_rng(seed)
    Y = -15.0 + 4.0 + 3.0 * np.sin(0.7 * X) * np.cos(0.9 * Z) + 0.5 * rng.standard_normal(X.shape)
    verts = np.stack([X, Y, Z], axis=-1).reshape(-1, 3)
    i, j = np.meshgrid(np.arange(nx - 1), np.arange(nz - 1), indexing="ij")
    v00 = (i * nz + j).reshape(-1)
    v10, v01, v11 = v00 + nz, v00 + 1, v00 + nz + 1
    faces = np.concatenate([np.stack([v00, v01, v11], 1), np.stack([v00, v11, v10], 1)])
    return TriangleMesh(verts, faces, Material(color=Vec3(0.8, 0.8, 0.8), diffuse=0.8))


def heightfield_scene(nx: int = 1001, nz: int = 501, seed: int = 1234):
    """-> (scene, builder): 5 Cornell walls + (nx-1)(nz-1)*2 triangles + the 4x4 light samples."""
    b = CustomSceneBuilder(texture_dir=False)
    scene = Scene()
    b._create_walls(scene, b._create_wall_materials())
    scene.objects.append(heightfield_mesh(nx, nz, seed))
    b._create_lighting(scene)
    scene.light_color = Vec3(0.7, 0.7, 0.7)
    scene.ambient = Vec3(0.5, 0.5, 0.5)
    return scene, b
