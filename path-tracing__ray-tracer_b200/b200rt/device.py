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


class _Staging:
    """Per-device pinned staging buffer, grown on demand and reused by every upload (cudaHostAlloc is slow:
    allocating pinned memory per array used to cost more than the copies themselves)."""

    _pool = {}

    @classmethod
    def get(cls, device, nbytes: int) -> torch.Tensor:
        key = str(device)
        buf, ev = cls._pool.get(key, (None, None))
        if ev is not None:
            ev.synchronize()                      # the previous async copy out of this buffer has finished
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, pin_memory=True)
        cls._pool[key] = (buf, None)
        return buf

    @classmethod
    def mark(cls, device) -> None:
        key = str(device)
        buf, _ = cls._pool[key]
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(device))
        cls._pool[key] = (buf, ev)


class Blob:
    """Collects host arrays, ships them with ONE pinned H2D copy, hands back 256-byte-aligned device views."""

    def __init__(self):
        self.items, self.size = [], 0

    def add(self, name: str, arr: np.ndarray) -> None:
        arr = np.ascontiguousarray(arr)
        off = (self.size + 255) & ~255
        self.items.append((name, arr, off))
        self.size = off + max(arr.nbytes, 16)

    def upload(self, device) -> dict:
        total = (self.size + 255) & ~255
        host = _Staging.get(device, total)
        hv = host.numpy()
        for _, arr, off in self.items:
            if arr.nbytes:
                hv[off:off + arr.nbytes] = arr.reshape(-1).view(np.uint8)
        dev = torch.empty(total, dtype=torch.uint8, device=device)
        dev.copy_(host[:total], non_blocking=True)
        _Staging.mark(device)
        self.host_image = bytes(hv[:total])      # kept so that a cached DeviceScene can re-send its inputs (reupload)
        out = {"_blob": dev}
        for name, arr, off in self.items:
            n = max(arr.nbytes, 16)
            view = dev[off:off + n].view(_TORCH_DTYPE[arr.dtype.str])
            out[name] = view[:arr.size] if arr.nbytes else view[:0]
        self.nbytes = total
        return out


_TORCH_DTYPE = {"<f4": torch.float32, "<f8": torch.float64, "<i4": torch.int32, "<u4": torch.int32, "|u1": torch.uint8,
                "<i8": torch.int64}


def to_device(arr: np.ndarray, device, pinned: bool = True) -> torch.Tensor:
    """One H2D copy through the pinned staging buffer."""
    b = Blob()
    b.add("x", arr)
    return b.upload(device)["x"]


class DeviceScene:
    """Uploads a ``PackedScene`` in the requested precision and builds its LBVH on the device."""

    def __init__(self, packed: PackedScene, precision: int = _lib.P_F32, device=None,
                 top_nodes: int = TOP_NODES_DEFAULT, ray_origin_extent: float = 0.0, textures_dev=None,
                 scan_max_prims: int = SCAN_MAX_PRIMS, occluder_hints: bool = True, ray_sort_min_prims: int = 4096,
                 scan_boxes: bool = True, surface_records: bool = True, rects_outside: bool = True,
                 lbvh_rotations: bool = True, prepare: str = "library", wide_nodes: bool = False,
                 quant_nodes: bool = False):
        """``prepare``: who derives the small-scene records (scan / box / surface records, occluder hints, bounds) —
        ``"library"`` = ``b2rt_scene_prepare_host`` inside ``libb200rt.so`` (what any C-ABI binder gets), ``"numpy"`` =
        the independent implementation in ``packer.py`` (kept as the cross-check; the CPU tests compare the two)."""
        self.lib = _lib.load()
        self.device = require_cuda(device)
        self.packed = packed
        self.precision = precision
        real = np.float64 if precision == _lib.P_F64 else np.float32
        dev = self.device
        with torch.cuda.device(dev):
            blob = Blob()
            for name in ("rect", "sphere", "tri", "shade", "mat", "lights"):
                blob.add(name, getattr(packed, name).astype(real))
            blob.add("prim_mat", packed.prim_mat)
            blob.add("mat_tex", packed.mat_tex)
            if precision == _lib.P_F64:          # the LBVH builder always consumes float32 geometry
                for name in ("rect", "sphere", "tri"):
                    blob.add("g_" + name, getattr(packed, name).astype(np.float32))
            if textures_dev is None:
                blob.add("texels", packed.texels.view(np.int32))
                blob.add("tex_info", packed.tex_info if packed.n_tex else np.zeros((1, 4), np.int32))
            self.scan_host = self.occluder_hint_host = None
            scan_ok = 0 < packed.n_prims <= scan_max_prims
            self.prepared_by = None
            if scan_ok and precision == _lib.P_F32 and packed.semantics == 0:
                if prepare == "library" and scan_max_prims <= 64:
                    lay, host = _library_records(self.lib, packed, occluder_hints, scan_boxes, surface_records)
                    if lay is not None and lay.scan_offset != _NONE:
                        nrec = lay.n_scan_prims + lay.n_scan_boxes
                        rec = np.frombuffer(host, dtype=np.float32, count=16 * nrec, offset=lay.scan_offset).reshape(-1, 4)
                        self.scan_host = rec[:4 * lay.n_scan_prims]
                        self.scan_boxes_host = rec[4 * lay.n_scan_prims:]
                        self.n_scan_loose = int(lay.n_scan_loose)
                        blob.add("scan", rec)
                        if lay.surface_offset != _NONE:
                            blob.add("surf", np.frombuffer(host, dtype=np.float32, count=20 * packed.n_prims,
                                                           offset=lay.surface_offset).reshape(-1, 4))
                        if lay.hint_offset != _NONE:
                            self.occluder_hint_host = np.frombuffer(host, dtype=np.int32, count=packed.lights.shape[0],
                                                                    offset=lay.hint_offset).copy()
                        self.prepared_by = "library"
                else:
                    self.scan_host, self.occluder_hint_host, self.n_scan_loose, self.scan_boxes_host = \
                        _small_scene_records(packed, occluder_hints, scan_boxes)
                    if self.scan_host is not None:
                        from .packer import build_surface_records
                        blob.add("scan", np.concatenate([self.scan_host, self.scan_boxes_host]))
                        if surface_records:
                            blob.add("surf", build_surface_records(packed))
                        self.prepared_by = "numpy"
            elif (not scan_ok and precision == _lib.P_F32 and packed.semantics == 0 and occluder_hints
                  and 0 < packed.n_rect <= 64 and packed.n_sphere <= 64 and 0 < packed.lights.shape[0] <= 4096):
                from .packer import build_occluder_hints, rect_scan_records
                self.occluder_hint_host = build_occluder_hints(packed, rect_scan_records(packed), generic=True)
            if self.occluder_hint_host is not None:
                blob.add("hint", self.occluder_hint_host)
            d = blob.upload(dev)
            self._blob, self.h2d_small, self._blob_host = d["_blob"], blob.nbytes, blob.host_image
            self.rect, self.sphere, self.tri, self.shade = d["rect"], d["sphere"], d["tri"], d["shade"]
            self.mat, self.lights, self.prim_mat, self.mat_tex = d["mat"], d["lights"], d["prim_mat"], d["mat_tex"]
            if textures_dev is not None:
                self.texels, self.tex_info = textures_dev
            else:
                self.texels, self.tex_info = d["texels"], d["tex_info"]
            if precision == _lib.P_F64:
                g_rect, g_sphere, g_tri = d["g_rect"], d["g_sphere"], d["g_tri"]
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
            # a few (room-sized) rectangles around a large mesh stay outside the hierarchy and are tested directly
            self.rects_outside = bool(rects_outside and not scan_ok and 0 < packed.n_rect <= 16 and n >= 4096)
            _lib.check(self.lib.b2rt_lbvh_build(packed.n_rect, packed.n_sphere, packed.n_tri,
                                                g_rect.data_ptr(), g_sphere.data_ptr(), g_tri.data_ptr(),
                                                C.c_float(pad), self.nodes.data_ptr(), self.top.data_ptr(),
                                                top_nodes, meta, temp.data_ptr(), temp.numel(),
                                                current_stream_ptr(dev),
                                                (1 if self.rects_outside else 0) | (0 if lbvh_rotations else 2)),
                       "b2rt_lbvh_build")
            self.n_top, self.root, self.n_internal = int(meta[0]), int(meta[1]), int(meta[2])
            # wide_nodes (off by default): 4-wide nodes for the persistent walk kernel of float32 scenes too large for the
            # record scan, derived on the device from the finished tree (128 B per node reference).  Measured on the
            # 1 M-triangle scene: 14.4 instead of 28.8 box steps per ray, identical hits, but 65.0 vs 62.4 ms per step in
            # the walk kernel (profiles/r2_c4_wide_nodes_ab.log, r2_c4_walk_kernel_analysis.md)
            self.wide = None
            if wide_nodes and precision == _lib.P_F32 and not scan_ok and self.n_internal > 0:
                wb = C.c_size_t(0)
                _lib.check(self.lib.b2rt_lbvh_wide_bytes(self.n_top, self.n_internal, C.byref(wb)), "b2rt_lbvh_wide_bytes")
                self.wide = torch.empty(wb.value, dtype=torch.uint8, device=dev)
                _lib.check(self.lib.b2rt_lbvh_widen(self.nodes.data_ptr(), self.top.data_ptr(), self.n_top, self.n_internal,
                                                    self.wide.data_ptr(), wb.value, current_stream_ptr(dev)),
                           "b2rt_lbvh_widen")
            # quant_nodes (off by default): 32 B nodes with both child boxes on a 16-bit grid over the scene bounds, for the
            # same kernel: two 16-byte loads per node instead of four (the walk is bound by the L1 data pipe)
            self.quant = None
            if quant_nodes and precision == _lib.P_F32 and not scan_ok and self.n_internal > 0:
                qb = C.c_size_t(0)
                _lib.check(self.lib.b2rt_lbvh_quant_bytes(self.n_top, self.n_internal, C.byref(qb)), "b2rt_lbvh_quant_bytes")
                self.quant = torch.empty(qb.value, dtype=torch.uint8, device=dev)
                glo, ghi = packed.bounds()
                slack = 4.0 * pad + 1e-5 * float(packed.max_abs_coordinate())      # node boxes = primitive boxes + pad
                lo3 = (C.c_float * 3)(*[float(v) - slack for v in glo])
                hi3 = (C.c_float * 3)(*[float(v) + slack for v in ghi])
                _lib.check(self.lib.b2rt_lbvh_quantize(self.nodes.data_ptr(), self.top.data_ptr(), self.n_top, self.n_internal,
                                                       lo3, hi3, self.quant.data_ptr(), qb.value, current_stream_ptr(dev)),
                           "b2rt_lbvh_quantize")
        s = _lib.new_scene_struct()
        s.precision, s.semantics = precision, packed.semantics
        s.n_rect, s.n_sphere, s.n_tri = packed.n_rect, packed.n_sphere, packed.n_tri
        s.n_mat, s.n_tex, s.n_lights = packed.n_mat, packed.n_tex, packed.lights.shape[0]
        s.d_rect, s.d_sphere, s.d_tri, s.d_shade = (t.data_ptr() for t in (self.rect, self.sphere, self.tri, self.shade))
        s.d_prim_mat, s.d_mat, s.d_mat_tex = self.prim_mat.data_ptr(), self.mat.data_ptr(), self.mat_tex.data_ptr()
        s.d_texels, s.d_tex_info, s.d_lights = self.texels.data_ptr(), self.tex_info.data_ptr(), self.lights.data_ptr()
        s.d_bvh_nodes, s.d_bvh_top = self.nodes.data_ptr(), self.top.data_ptr()
        s.d_bvh_wide = self.wide.data_ptr() if self.wide is not None else None
        s.d_bvh_quant = self.quant.data_ptr() if self.quant is not None else None
        s.n_bvh_top, s.bvh_root = self.n_top, self.root
        s.scan_incoherent = 1 if scan_ok else 0
        s.bvh_rects_outside = 1 if self.rects_outside else 0
        s.ray_sort_extent = float(packed.max_abs_coordinate()) if (not scan_ok and packed.n_prims >= ray_sort_min_prims) else 0.0
        s.n_scan_prims, s.d_scan_prims, s.d_occluder_hint, s.d_surface_records = 0, None, None, None
        blo, bhi = packed.bounds()
        for k in range(3):
            s.bounds_lo[k], s.bounds_hi[k] = float(blo[k]), float(bhi[k])
        if self.scan_host is not None:
            self.scan_prims = d["scan"]
            s.n_scan_prims, s.d_scan_prims = self.scan_host.shape[0] // 4, self.scan_prims.data_ptr()
            s.n_scan_loose, s.n_scan_boxes = self.n_scan_loose, self.scan_boxes_host.shape[0] // 4
            if "surf" in d:
                self.surface_records = d["surf"]
                s.d_surface_records = self.surface_records.data_ptr()
        if self.occluder_hint_host is not None:
            self.occluder_hint = d["hint"]
            s.d_occluder_hint = self.occluder_hint.data_ptr()
        self.struct = s

    def reupload(self, textures_dev=None) -> None:
        """Send the scene's host image to the device again (one pinned H2D copy into the same buffer) and point the struct
        at a fresh texture upload.  A renderer that finds the scene unchanged (value signature) keeps this object — its
        LBVH and derived records are functions of the same bytes — but every ``render()`` still moves its inputs."""
        n = len(self._blob_host)
        host = _Staging.get(self.device, n)
        host.numpy()[:n] = np.frombuffer(self._blob_host, dtype=np.uint8)
        self._blob.copy_(host[:n], non_blocking=True)
        _Staging.mark(self.device)
        if textures_dev is not None:
            self.texels, self.tex_info = textures_dev
            self.struct.d_texels, self.struct.d_tex_info = self.texels.data_ptr(), self.tex_info.data_ptr()

    def ref(self):
        return C.byref(self.struct)

    def h2d_bytes(self) -> int:
        return int(self.h2d_small)


_small_cache: dict = {}
_NONE = (1 << 64) - 1          # (size_t)-1: "not produced" in b2rt_prepare_layout


def _library_records(lib, packed: PackedScene, want_hints: bool, want_boxes: bool, want_surface: bool):
    """Small-scene records from ``b2rt_scene_prepare_host`` (host C++ inside the library), cached on the bytes of the
    packed streams like the numpy path -> (layout, host buffer)."""
    import hashlib
    f32 = {k: np.ascontiguousarray(getattr(packed, k), dtype=np.float32) for k in ("rect", "sphere", "tri", "shade", "mat", "lights")}
    pm, mt = np.ascontiguousarray(packed.prim_mat, dtype=np.int32), np.ascontiguousarray(packed.mat_tex, dtype=np.int32)
    h = hashlib.blake2b(digest_size=16)
    for a in (*f32.values(), pm, mt):
        h.update(a.tobytes())
    key = ("lib", h.hexdigest(), bool(want_hints), bool(want_boxes), bool(want_surface))
    if key not in _small_cache:
        need = C.c_size_t(0)
        n_lights = int(packed.lights.shape[0])
        lib.b2rt_scene_prepare_bytes(packed.n_rect, packed.n_sphere, packed.n_tri, n_lights, C.byref(need))
        out = np.zeros(need.value, dtype=np.uint8)
        lay = _lib.PrepareLayout()
        flags = (0 if want_boxes else 1) | (0 if want_surface else 2) | (0 if want_hints else 4)
        ptr = lambda a: a.ctypes.data_as(C.c_void_p)
        _lib.check(lib.b2rt_scene_prepare_host(packed.n_rect, packed.n_sphere, packed.n_tri, packed.n_mat, n_lights,
                                               ptr(f32["rect"]), ptr(f32["sphere"]), ptr(f32["tri"]), ptr(f32["shade"]),
                                               ptr(f32["mat"]), ptr(pm), ptr(mt), ptr(f32["lights"]), flags, ptr(out),
                                               out.nbytes, C.byref(lay)), "b2rt_scene_prepare_host")
        _small_cache[key] = (lay, out)
        if len(_small_cache) > 64:
            _small_cache.pop(next(iter(_small_cache)))
    return _small_cache[key]


def _small_scene_records(packed: PackedScene, want_hints: bool, want_boxes: bool = True):
    """Scan records (planar, loose count, boxes) + occluder hints, cached on the bytes of the packed geometry
    (pure functions of it) -> (planar records, hints, n_loose, box records)."""
    import hashlib
    from .packer import build_occluder_hints, build_scan_prims, group_scan_boxes
    h = hashlib.blake2b(digest_size=16)
    for a in (packed.rect, packed.sphere, packed.tri, packed.lights):
        h.update(np.ascontiguousarray(a).tobytes())
    key = (h.hexdigest(), bool(want_hints), bool(want_boxes))
    if key not in _small_cache:
        quads = []
        rec = build_scan_prims(packed, quads_out=quads)
        if not (0 < rec.shape[0] // 4 <= 64):
            _small_cache[key] = (None, None, 0, None)
        else:
            n_loose, boxes = rec.shape[0] // 4, np.zeros((0, 4), np.float32)
            if want_boxes:
                rec, n_loose, boxes = group_scan_boxes(rec, quads)
            hints = build_occluder_hints(packed, rec) if (want_hints and 0 < packed.lights.shape[0] <= 4096) else None
            _small_cache[key] = (rec, hints, n_loose, boxes)
        if len(_small_cache) > 64:
            _small_cache.pop(next(iter(_small_cache)))
    return _small_cache[key]
