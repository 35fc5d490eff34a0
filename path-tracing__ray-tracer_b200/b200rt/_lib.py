"""ctypes binding of ``libb200rt.so`` (the C ABI of ``include/b200rt.h``).

There is NO CPU fallback: if the library cannot be loaded the import of a renderer fails loudly
(``RuntimeError``), like the reference's ``_check_cuda_available`` (``cuda_path_tracer.py:741-746``).
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

P_F32, P_F64 = 0, 1
RNG_PCG, RNG_REFERENCE = 0, 1

EXPORTS = [
    "b2rt_last_error", "b2rt_version", "b2rt_device_info", "b2rt_lbvh_temp_bytes", "b2rt_lbvh_build",
    "b2rt_primary_hits", "b2rt_trace_rays", "b2rt_render_whitted_cpu", "b2rt_render_whitted_texture",
    "b2rt_path_workspace_bytes", "b2rt_render_path", "b2rt_resolve",
    "b2rt_profile_enable", "b2rt_profile_read", "b2rt_fp32_peak", "b2rt_reduce_resolve", "b2rt_expand_rgb8",
    "b2rt_scene_prepare_bytes", "b2rt_scene_prepare", "b2rt_scene_prepare_host",
    "b2rt_check_enabled", "b2rt_check_read", "b2rt_lbvh_wide_bytes", "b2rt_lbvh_widen",
    "b2rt_lbvh_quant_bytes", "b2rt_lbvh_quantize",
]
ABI_VERSION = 3


class PrepareLayout(C.Structure):
    """``struct b2rt_prepare_layout`` (include/b200rt.h)."""
    _fields_ = [("scan_offset", C.c_size_t), ("surface_offset", C.c_size_t), ("hint_offset", C.c_size_t),
                ("bytes_used", C.c_size_t), ("n_scan_prims", C.c_int32), ("n_scan_loose", C.c_int32),
                ("n_scan_boxes", C.c_int32), ("scan_incoherent", C.c_int32),
                ("bounds_lo", C.c_float * 3), ("bounds_hi", C.c_float * 3)]


def new_scene_struct():
    s = SceneStruct()
    s.struct_size, s.abi_version = C.sizeof(SceneStruct), ABI_VERSION
    return s


class SceneStruct(C.Structure):
    """``struct b2rt_scene`` (include/b200rt.h)."""
    _fields_ = [
        ("struct_size", C.c_uint32), ("abi_version", C.c_uint32),
        ("precision", C.c_int32), ("semantics", C.c_int32),
        ("n_rect", C.c_int32), ("n_sphere", C.c_int32), ("n_tri", C.c_int32),
        ("n_mat", C.c_int32), ("n_tex", C.c_int32), ("n_lights", C.c_int32),
        ("d_rect", C.c_void_p), ("d_sphere", C.c_void_p), ("d_tri", C.c_void_p), ("d_shade", C.c_void_p),
        ("d_prim_mat", C.c_void_p), ("d_mat", C.c_void_p), ("d_mat_tex", C.c_void_p),
        ("d_texels", C.c_void_p), ("d_tex_info", C.c_void_p), ("d_lights", C.c_void_p),
        ("d_bvh_nodes", C.c_void_p), ("d_bvh_top", C.c_void_p),
        ("n_bvh_top", C.c_int32), ("bvh_root", C.c_int32), ("scan_incoherent", C.c_int32), ("n_scan_prims", C.c_int32), ("d_scan_prims", C.c_void_p), ("d_occluder_hint", C.c_void_p), ("ray_sort_extent", C.c_float),
        ("n_scan_loose", C.c_int32), ("n_scan_boxes", C.c_int32), ("bvh_rects_outside", C.c_int32),
        ("d_surface_records", C.c_void_p), ("bounds_lo", C.c_float * 3), ("bounds_hi", C.c_float * 3),
        ("d_bvh_wide", C.c_void_p), ("d_bvh_quant", C.c_void_p),
    ]


_lib = None


def lib_path() -> str:
    return os.environ.get("B200RT_LIB", _build.LIB_PATH)


def load():
    """Load (building in-tree with nvcc first if the .so is missing or stale)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if "B200RT_LIB" not in os.environ and _build.is_stale():
        _build.build()
    if not os.path.isfile(path):
        raise RuntimeError(f"libb200rt.so not found at {path}; run `python -m b200rt.build` (needs nvcc). "
                           "b200rt has no CPU fallback.")
    lib = C.CDLL(path)
    lib.b2rt_last_error.restype = C.c_char_p
    vp, i32, i64, u64, dbl, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double, C.c_size_t
    SP = C.POINTER(SceneStruct)
    lib.b2rt_device_info.argtypes = [C.c_int, C.POINTER(i64)]
    lib.b2rt_lbvh_temp_bytes.argtypes = [i32, C.POINTER(sz)]
    lib.b2rt_lbvh_build.argtypes = [i32, i32, i32, vp, vp, vp, C.c_float, vp, vp, i32, C.POINTER(i32), vp, sz, vp, i32]
    lib.b2rt_lbvh_wide_bytes.argtypes = [i32, i32, C.POINTER(sz)]
    lib.b2rt_lbvh_widen.argtypes = [vp, vp, i32, i32, vp, sz, vp]
    lib.b2rt_lbvh_quant_bytes.argtypes = [i32, i32, C.POINTER(sz)]
    lib.b2rt_lbvh_quantize.argtypes = [vp, vp, i32, i32, C.POINTER(C.c_float), C.POINTER(C.c_float), vp, sz, vp]
    lib.b2rt_primary_hits.argtypes = [SP, C.POINTER(dbl), i32, i32, dbl, dbl, dbl, dbl, i32, vp, vp, vp]
    lib.b2rt_trace_rays.argtypes = [SP, i32, vp, vp, dbl, dbl, i32, i32, vp, vp, vp]
    lib.b2rt_render_whitted_cpu.argtypes = [SP, C.POINTER(dbl), i32, i32, vp, i32, C.POINTER(dbl), C.POINTER(dbl), vp, vp]
    lib.b2rt_render_whitted_texture.argtypes = [SP, C.POINTER(dbl), i32, i32, i32, i32, vp, vp, vp]
    lib.b2rt_path_workspace_bytes.argtypes = [i32, i32, i32, i32, i32, C.POINTER(sz)]
    lib.b2rt_render_path.argtypes = [SP, C.POINTER(dbl), i32, i32, i32, i64, i32, i32, i32, u64, i32, vp, vp, vp, vp, sz, vp, vp]
    lib.b2rt_resolve.argtypes = [i32, vp, i32, i32, dbl, i32, vp, vp]
    lib.b2rt_profile_enable.argtypes = [i32]
    lib.b2rt_profile_read.argtypes = [C.POINTER(dbl), C.POINTER(i64)]
    lib.b2rt_fp32_peak.argtypes = [i32, C.POINTER(dbl), vp]
    lib.b2rt_reduce_resolve.argtypes = [C.POINTER(vp), i32, i32, i32, i32, i32, dbl, i32, vp, vp, vp]
    lib.b2rt_expand_rgb8.argtypes = [vp, i64, vp, vp]
    lib.b2rt_check_read.argtypes = [C.POINTER(u64), C.POINTER(u64)]
    lib.b2rt_scene_prepare_bytes.argtypes = [i32, i32, i32, i32, C.POINTER(sz)]
    lib.b2rt_scene_prepare.argtypes = [SP, vp, sz, i32, vp]
    lib.b2rt_scene_prepare_host.argtypes = [i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, sz,
                                            C.POINTER(PrepareLayout)]
    for name in EXPORTS:
        if name not in ("b2rt_last_error",):
            getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().b2rt_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")


def dbl_array(values):
    arr = (C.c_double * len(values))(*[float(v) for v in values])
    return arr
