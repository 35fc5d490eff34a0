"""Opt-in node formats of the large-scene walk kernel (profiles/r2_c4_walk_kernel_analysis.md): the
quantised 32 B nodes must CONTAIN the float32 boxes they replace and give the same closest hits, hence the same float
sums, as the default 64 B nodes.  (Sorted last on purpose: these formats are off by default.)"""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("no CUDA device", allow_module_level=True)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "path-tracing__ray-tracer_b200"))
from b200rt import _lib, renderer, scenes  # noqa: E402
from b200rt.device import DeviceScene  # noqa: E402
from b200rt.packer import pack_scene  # noqa: E402
from b200rt.scene_api import RenderSettings  # noqa: E402


def test_quantised_nodes_contain_the_float_boxes():
    """b2rt_lbvh_quantize: every child box, dequantised as base + q * cell, contains the float32 box of the same child with
    between ~1 and ~2.1 cells to spare on every face (the far-away placeholders of rectangles kept outside the hierarchy
    collapse into the last cell), and the references are copied unchanged."""
    scene, _ = scenes.heightfield_scene(nx=101, nz=51)              # 10 000 triangles + 5 room rectangles outside the tree
    ds = DeviceScene(pack_scene(scene, "numba"), _lib.P_F32, quant_nodes=True)
    assert ds.quant is not None and ds.struct.d_bvh_quant
    torch.cuda.synchronize()
    n_top, n_int = ds.n_top, ds.n_internal
    nodes = ds.nodes.cpu().numpy().reshape(-1, 16)[:n_int]
    top = ds.top.cpu().numpy().reshape(-1, 16)[:n_top]
    raw = ds.quant.cpu().numpy()
    base, cell = raw[:12].view(np.float32).astype(np.float64), raw[16:28].view(np.float32).astype(np.float64)
    q = raw[32:32 + 32 * (n_top + n_int)].view(np.uint32).reshape(-1, 8)
    rec = np.concatenate([top, nodes])                               # indexed by the child reference like the quantised array
    lo_f = np.stack([rec[:, 0:3], rec[:, 6:9]], 1).astype(np.float64)       # [ref, child, axis]
    hi_f = np.stack([rec[:, 3:6], rec[:, 9:12]], 1).astype(np.float64)
    w = q[:, :6].reshape(-1, 2, 3)
    lo_q, hi_q = base + (w & 0xffff) * cell, base + (w >> 16) * cell
    inside = np.abs(lo_f) < 1e30                                     # everything but the far-away placeholder boxes
    assert inside.mean() > 0.99
    m_lo, m_hi = ((lo_f - lo_q) / cell)[inside], ((hi_q - hi_f) / cell)[inside]
    assert m_lo.min() > 0.9 and m_hi.min() > 0.9, (m_lo.min(), m_hi.min())
    assert m_lo.max() < 2.2 and m_hi.max() < 2.2, (m_lo.max(), m_hi.max())
    assert np.array_equal(q[:, 6:8].view(np.int32), rec[:, 12:14].view(np.int32))
    assert np.all((w[~inside.all(axis=2)] >> 16) == 65535)           # placeholders: clamped into the last cell


def test_quantised_and_wide_walks_give_the_default_sums():
    scene, b = scenes.heightfield_scene(nx=101, nz=51)
    cam = b.create_camera(16 / 9)
    st = RenderSettings(256, 144, 8, 4)
    out = {}
    for name, kw in (("binary", {}), ("quant", {"quant_walk": True}), ("wide", {"wide_walk": True})):
        r = renderer.B200PathTracer(precision="f32", rng="pcg", seed=11, **kw)
        out[name] = r.render_accum(scene, cam, st)
    for name in ("quant", "wide"):
        assert np.array_equal(out[name][1][:4], out["binary"][1][:4]), name
        assert np.array_equal(out[name][0], out["binary"][0]), name
