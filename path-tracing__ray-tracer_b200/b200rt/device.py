"""Device-resident scene: torch tensors as plain HBM buffers + the LBVH built by the library."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib
from .packer import PackedScene

TOP_NODES_DEFAULT = 512      # BVH nodes staged in shared memory per CTA (64 B each -> 32 KB)
SCAN_MAX_PRIMS = 64          # scenes this small scan all primitives for incoherent rays (b2rt_scene.scan_incoherent)


def require_cuda(device: Optional[torch.device] = None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("b200rt: no CUDA device is available (the B200 core has no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def current_stream_ptr(device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


def to_device(arr: np.ndarray, device, pinned: bool = True) -> torch.Tensor:
    """One H2D copy through pinned memory (the only way scene bytes reach HBM)."""
    t = torch.from_numpy(np.ascontiguousarray(arr))
    if t.numel() == 0:
        return torch.zeros(max(1, 4), dtype=t.dtype, device=device)[:0].contiguous()
    if pinned:
        t = t.pin_memory()
    return t.to(device, non_blocking=True)


class DeviceScene:
    """Uploads a ``PackedScene`` in the requested precision and builds its LBVH on the device."""

    def __init__(self, packed: PackedScene, precision: int = _lib.P_F32, device=None,
                 top_nodes: int = TOP_NODES_DEFAULT, ray_origin_extent: float = 0.0, textures_dev=None,
                 scan_max_prims: int = SCAN_MAX_PRIMS, occluder_hints: bool = True):
        self.lib = _lib.load()
        self.device = require_cuda(device)
        self.packed = packed
        self.precision = precision
        real = np.float64 if precision == _lib.P_F64 else np.float32
        dev = self.device
        with torch.cuda.device(dev):
            self.rect = to_device(packed.rect.astype(real), dev)
            self.sphere = to_device(packed.sphere.astype(real), dev)
            self.tri = to_device(packed.tri.astype(real), dev)
            self.shade = to_device(packed.shade.astype(real), dev)
            self.mat = to_device(packed.mat.astype(real), dev)
            self.lights = to_device(packed.lights.astype(real), dev)
            self.prim_mat = to_device(packed.prim_mat, dev)
            self.mat_tex = to_device(packed.mat_tex, dev)
            if textures_dev is not None:
                self.texels, self.tex_info = textures_dev
            else:
                self.texels = to_device(packed.texels.view(np.int32), dev)
                self.tex_info = to_device(packed.tex_info if packed.n_tex else np.zeros((1, 4), np.int32), dev)
            # the builder always consumes float32 geometry
            if precision == _lib.P_F64:
                g_rect = to_device(packed.rect.astype(np.float32), dev)
                g_sphere = to_device(packed.sphere.astype(np.float32), dev)
                g_tri = to_device(packed.tri.astype(np.float32), dev)
            else:
                g_rect, g_sphere, g_tri = self.rect, self.sphere, self.tri
            n = packed.n_prims
            need = C.c_size_t(0)
            _lib.check(self.lib.b2rt_lbvh_temp_bytes(n, C.byref(need)), "b2rt_lbvh_temp_bytes")
            temp = torch.empty(max(256, need.value), dtype=torch.uint8, device=dev)
            top_nodes = int(max(0, min(top_nodes, 1024)))
            self.nodes = torch.zeros(max(1, n - 1) * 16, dtype=torch.float32, device=dev)
            self.top = torch.zeros(max(1, top_nodes) * 16, dtype=torch.float32, device=dev)
            meta = (C.c_int32 * 3)()
            # pad covers float32 rounding of slab distances for rays that start up to ray_origin_extent away
            pad = 1e-5 * max(packed.max_abs_coordinate(), ray_origin_extent, 1e-3)
            self.box_pad = float(pad)
            _lib.check(self.lib.b2rt_lbvh_build(packed.n_rect, packed.n_sphere, packed.n_tri,
                                                g_rect.data_ptr(), g_sphere.data_ptr(), g_tri.data_ptr(),
                                                C.c_float(pad), self.nodes.data_ptr(), self.top.data_ptr(),
                                                top_nodes, meta, temp.data_ptr(), temp.numel(),
                                                current_stream_ptr(dev)), "b2rt_lbvh_build")
            self.n_top, self.root, self.n_internal = int(meta[0]), int(meta[1]), int(meta[2])
        s = _lib.SceneStruct()
        s.precision, s.semantics = precision, packed.semantics
        s.n_rect, s.n_sphere, s.n_tri = packed.n_rect, packed.n_sphere, packed.n_tri
        s.n_mat, s.n_tex, s.n_lights = packed.n_mat, packed.n_tex, packed.lights.shape[0]
        s.d_rect, s.d_sphere, s.d_tri, s.d_shade = (t.data_ptr() for t in (self.rect, self.sphere, self.tri, self.shade))
        s.d_prim_mat, s.d_mat, s.d_mat_tex = self.prim_mat.data_ptr(), self.mat.data_ptr(), self.mat_tex.data_ptr()
        s.d_texels, s.d_tex_info, s.d_lights = self.texels.data_ptr(), self.tex_info.data_ptr(), self.lights.data_ptr()
        s.d_bvh_nodes, s.d_bvh_top = self.nodes.data_ptr(), self.top.data_ptr()
        s.n_bvh_top, s.bvh_root = self.n_top, self.root
        s.scan_incoherent = 1 if 0 < packed.n_prims <= scan_max_prims else 0
        self.scan_prims = None
        s.n_scan_prims, s.d_scan_prims, s.d_occluder_hint = 0, None, None
        self.occluder_hint = None
        if s.scan_incoherent and precision == _lib.P_F32 and packed.semantics == 0:
            from .packer import build_scan_prims
            rec = build_scan_prims(packed)
            if 0 < rec.shape[0] // 4 <= 64:
                with torch.cuda.device(dev):
                    self.scan_prims = to_device(rec, dev)
                s.n_scan_prims, s.d_scan_prims = rec.shape[0] // 4, self.scan_prims.data_ptr()
                if occluder_hints and 0 < packed.lights.shape[0] <= 4096:
                    from .packer import build_occluder_hints
                    self.occluder_hint_host = build_occluder_hints(packed, rec)
                    with torch.cuda.device(dev):
                        self.occluder_hint = to_device(self.occluder_hint_host, dev)
                    s.d_occluder_hint = self.occluder_hint.data_ptr()
        self.struct = s

    def ref(self):
        return C.byref(self.struct)

    def h2d_bytes(self) -> int:
        ts = [self.rect, self.sphere, self.tri, self.shade, self.mat, self.lights, self.prim_mat, self.mat_tex,
              self.texels, self.tex_info]
        return int(sum(t.numel() * t.element_size() for t in ts))
