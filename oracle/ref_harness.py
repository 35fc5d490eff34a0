"""Container-only harness that runs the REAL reference (``/root/reference``) — test infrastructure.

Used by ``oracle/make_golden.py`` (to generate ``tests/golden/*.npz``) and by the
container-only pin tests.  It cannot travel to the GPU box (the reference is not
there), so nothing imported by ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may
import this module.

Three ways of running the reference:
  * ``cpu_raytracer``          — imported and called unmodified;
  * numba renderers, *njit*    — the renderer's own source text with ``cuda.jit`` mapped to
    ``numba.njit`` and the two thread-index lines replaced by arguments; every device
    function (``cuda_scene_hit``, ``cuda_trace_path`` ...) is the reference's code, compiled
    for the host;
  * the reference's own host packers (``_prepare_*``) via ``object.__new__`` (skips the
    device check in ``__init__``).
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import random
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("B200RT_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "renderers", "cuda_path_tracer.py"))


@contextlib.contextmanager
def _quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


@contextlib.contextmanager
def reference_cwd(texture_root: str | None = None):
    """chdir to where ``textures/<name>.jpg`` resolves (reference root, or a dir with stand-ins)."""
    old = os.getcwd()
    os.chdir(texture_root or REF_ROOT)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    for top in ("renderers", "core", "scene_builders"):          # a same-named foreign package must not shadow the reference
        m = sys.modules.get(top)
        if m is not None and not str(getattr(m, "__file__", "") or "").startswith(REF_ROOT):
            for k in [k for k in sys.modules if k == top or k.startswith(top + ".")]:
                del sys.modules[k]
    try:
        yield
    finally:
        os.chdir(old)


def write_synthetic_texture_dir(root: str) -> str:
    """Write the synthetic textures as lossless PNG bytes under ``root/textures/<name>.jpg``.

    PIL sniffs the format from the content, so the reference's ``Image.open("textures/blue.jpg")``
    decodes exactly the synthetic pixels.
    """
    from PIL import Image
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                    "path-tracing__ray-tracer_b200"))
    from b200rt.cornell import TEXTURE_SIZES, synthetic_texture
    tdir = os.path.join(root, "textures")
    os.makedirs(tdir, exist_ok=True)
    for name in TEXTURE_SIZES:
        Image.fromarray(synthetic_texture(name), "RGB").save(os.path.join(tdir, f"{name}.jpg"), format="PNG")
    return root


def build_reference_scene(seed: int = 0, aspect: float = 16 / 9, texture_root: str | None = None):
    """``random.seed(seed); CustomSceneBuilder().build_scene()`` with the reference's own classes."""
    with reference_cwd(texture_root), _quiet():
        mod = importlib.import_module("scene_builders.custom_scene_builder")
        random.seed(seed)
        builder = mod.CustomSceneBuilder()
        scene = builder.build_scene()
        camera = builder.create_camera(aspect)
    return scene, camera


def reference_cpu_renderer():
    with reference_cwd(), _quiet():
        return importlib.import_module("renderers.cpu_renderer").CPURenderer()


# ----------------------------------------------------------------------------- njit shim
class _CudaShim:
    """Stands in for ``numba.cuda`` in the renderer source: ``@cuda.jit`` / ``@cuda.jit(device=True)``."""

    @staticmethod
    def jit(*args, **kwargs):
        import numba
        if args and callable(args[0]) and not kwargs:
            return numba.njit(cache=False)(args[0])
        return lambda fn: numba.njit(cache=False)(fn)


def _load_njit_module(filename: str, kernel_name: str) -> types.ModuleType:
    path = os.path.join(REF_ROOT, "renderers", filename)
    src = open(path, encoding="utf-8").read()
    src = src.replace("from numba import cuda", "cuda = __cuda_shim__")
    src = src.replace("x = cuda.blockIdx.x * cuda.blockDim.x + cuda.threadIdx.x", "x = px")
    src = src.replace("y = cuda.blockIdx.y * cuda.blockDim.y + cuda.threadIdx.y", "y = py")
    # add (px, py) to the kernel signature
    head = f"def {kernel_name}(output,"
    assert head in src, kernel_name
    src = src.replace(head, f"def {kernel_name}(px, py, output,")
    src = "\n".join(l for l in src.split("\n") if not l.startswith("RendererFactory.register("))
    mod = types.ModuleType("ref_njit_" + filename[:-3])
    mod.__dict__["__cuda_shim__"] = _CudaShim
    with reference_cwd(), _quiet():
        exec(compile(src, path, "exec"), mod.__dict__)
    return mod


_cache: dict = {}


def njit_path_tracer() -> types.ModuleType:
    if "path" not in _cache:
        _cache["path"] = _load_njit_module("cuda_path_tracer.py", "cuda_path_trace_kernel")
    return _cache["path"]


def njit_texture_renderer() -> types.ModuleType:
    if "tex" not in _cache:
        _cache["tex"] = _load_njit_module("cuda_texture_renderer.py", "cuda_trace_kernel")
    return _cache["tex"]


def reference_pack(scene, camera, which: str = "path", texture_root: str | None = None):
    """Run the reference's own ``_prepare_*`` packers -> dict of numpy arrays."""
    mod = njit_path_tracer() if which == "path" else njit_texture_renderer()
    cls = mod.CUDAPathTracer if which == "path" else mod.CUDATextureRenderer
    r = object.__new__(cls)
    with reference_cwd(texture_root), _quiet():
        out = dict(scene=r._prepare_scene_data(scene), camera=r._prepare_camera_data(camera),
                   lights=r._prepare_light_data(scene))
        tex, info = r._prepare_texture_data(scene)
    out["tex"], out["tex_info"] = tex, info
    return out


def run_path_kernel(packed, width, height, spp, max_depth, frame_count=0):
    """The reference's ``cuda_path_trace_kernel`` for every pixel -> uint8 [H*W*3] (device row order)."""
    import numba
    mod = njit_path_tracer()
    kern = mod.cuda_path_trace_kernel

    @numba.njit(parallel=True)
    def drive(out, sd, cd, ld, td, ti, w, h, spp, md, fc):
        for py in numba.prange(h):
            for px in range(w):
                kern(px, py, out, sd, cd, ld, td, ti, w, h, spp, md, fc)

    out = np.zeros(width * height * 3, dtype=np.uint8)
    drive(out, packed["scene"], packed["camera"], packed["lights"], packed["tex"], packed["tex_info"],
          width, height, spp, max_depth, frame_count)
    return out


def run_path_float(packed, width, height, spp, max_depth, frame_count=0):
    """Per-pixel sum and sum-of-squares of the per-sample radiance, following
    ``cuda_path_trace_kernel`` (:28-46) but calling the reference's own device functions."""
    import numba
    mod = njit_path_tracer()
    rnd, xs, get_ray, trace = mod.cuda_random, mod.cuda_xorshift, mod.cuda_get_ray, mod.cuda_trace_path

    @numba.njit(parallel=True)
    def drive(s1, s2, sd, cd, ld, td, ti, w, h, spp, md, fc):
        for y in numba.prange(h):
            for x in range(w):
                rng = (x + y * w + fc * w * h) * 1103515245 + 12345
                for _ in range(spp):
                    u = (x + rnd(rng)) / w
                    v = (y + rnd(rng)) / h
                    rng = xs(rng)
                    o, d = get_ray(cd, u, v)
                    r, g, b = trace(sd, ld, td, ti, o, d, md, rng)
                    k = (y * w + x) * 3
                    s1[k] += r; s1[k + 1] += g; s1[k + 2] += b
                    s2[k] += r * r; s2[k + 1] += g * g; s2[k + 2] += b * b
                    rng = xs(rng)

    s1 = np.zeros(width * height * 3); s2 = np.zeros(width * height * 3)
    drive(s1, s2, packed["scene"], packed["camera"], packed["lights"], packed["tex"], packed["tex_info"],
          width, height, spp, max_depth, frame_count)
    return s1, s2


def run_texture_kernel(packed, width, height, spp, max_depth):
    """The reference's textured-Whitted ``cuda_trace_kernel`` -> uint8 [H*W*3] (device row order)."""
    import numba
    mod = njit_texture_renderer()
    kern = mod.cuda_trace_kernel

    @numba.njit(parallel=True)
    def drive(out, sd, cd, ld, td, ti, w, h, spp, md):
        for py in numba.prange(h):
            for px in range(w):
                kern(px, py, out, sd, cd, ld, td, ti, w, h, spp, md)

    out = np.zeros(width * height * 3, dtype=np.uint8)
    drive(out, packed["scene"], packed["camera"], packed["lights"], packed["tex"], packed["tex_info"],
          width, height, spp, max_depth)
    return out


def run_scene_hit(packed, origins, dirs, t_min=0.001, t_max=1000000.0):
    """The reference's ``cuda_scene_hit`` on explicit rays -> (hit[n], rec[n,19])."""
    import numba
    hitfn = njit_path_tracer().cuda_scene_hit

    @numba.njit
    def drive(sd, o, d, tmin, tmax, hit, rec):
        for i in range(o.shape[0]):
            h, t, p, n, m, uv = hitfn(sd, (o[i, 0], o[i, 1], o[i, 2]), (d[i, 0], d[i, 1], d[i, 2]), tmin, tmax)
            hit[i] = 1 if h else 0
            rec[i, 0] = t
            for k in range(3):
                rec[i, 1 + k] = p[k]; rec[i, 4 + k] = n[k]
            rec[i, 7] = uv[0]; rec[i, 8] = uv[1]
            for k in range(10):
                rec[i, 9 + k] = m[k]

    n = origins.shape[0]
    hit = np.zeros(n, dtype=np.int32); rec = np.zeros((n, 19))
    drive(packed["scene"], np.ascontiguousarray(origins, dtype=np.float64),
          np.ascontiguousarray(dirs, dtype=np.float64), t_min, t_max, hit, rec)
    return hit, rec
