"""``python -m b200rt.main``: the reference's own ``main.py`` (``/root/reference/main.py:11-20,24-44,75,90``) runs
UNMODIFIED with the B200 renderers registered into its ``RendererFactory``.

CPU: the shim's argument handling; with ``/root/reference`` mounted (container only) the reference CLI really runs
through the shim (its own ``cpu_raytracer``) and lists the ``b200_*`` renderers as ``--renderer`` choices.
GPU: inside ``baseline/_ref`` the shim's PNG equals what ``RendererFactory.create(...).render(...)`` returns.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "path-tracing__ray-tracer_b200")

from b200rt import main as shim  # noqa: E402


def _env():
    env = dict(os.environ)
    env["PYTHONPATH"] = PKG + os.pathsep + env.get("PYTHONPATH", "")
    return env


def test_split_args_takes_only_the_shims_flags():
    opts, rest = shim.split_args(["-r", "b200_path_tracer", "--gpus", "4", "-w", "64", "--seed=7",
                                  "--reference-root", "/x/y", "--height", "48"])
    assert opts == {"gpus": 4, "reference_root": "/x/y", "seed": 7}
    assert rest == ["-r", "b200_path_tracer", "-w", "64", "--height", "48"]
    with pytest.raises(SystemExit):
        shim.split_args(["--gpus"])


def test_path_samples_reach_the_b200_path_tracer():
    """main.py:49-54 hands --path-samples only to a renderer NAMED cuda_path_raytracer."""
    assert shim.translate_args(["-r", "b200_path_tracer", "--path-samples", "256"])[-2:] == ["--samples", "256"]
    assert shim.translate_args(["--renderer=b200_path_tracer"])[-2:] == ["--samples", "1024"]          # main.py:43 default
    explicit = ["-r", "b200_path_tracer", "-s", "16", "--path-samples", "256"]
    assert shim.translate_args(explicit) == explicit
    other = ["-r", "b200_texture_raytracer", "--path-samples", "256"]
    assert shim.translate_args(other) == other


def test_torchrun_command_shape():
    cmd = shim.torchrun_command(4, ["-r", "b200_path_tracer"], port=29511)
    assert cmd[1:4] == ["-m", "torch.distributed.run", "--nnodes=1"] and "--nproc-per-node=4" in cmd
    assert cmd[cmd.index("--master-addr") + 1] == "127.0.0.1" and cmd[cmd.index("--master-port") + 1] == "29511"
    assert cmd[-4:] == ["-m", "b200rt.main", "-r", "b200_path_tracer"]


@pytest.mark.reference
@pytest.mark.skipif(not shim.is_reference_root("/root/reference"), reason="reference not mounted")
def test_reference_main_runs_unmodified_through_the_shim(tmp_path):
    h = subprocess.run([sys.executable, "-m", "b200rt.main", "--reference-root", "/root/reference", "--help"],
                       capture_output=True, text=True, env=_env(), cwd=str(tmp_path), timeout=300)
    assert h.returncode == 0, h.stderr[-500:]
    for name in ("b200_path_tracer", "b200_texture_raytracer", "b200_raytracer", "cpu_raytracer"):
        assert name in h.stdout
    out = tmp_path / "cpu.png"
    r = subprocess.run([sys.executable, "-m", "b200rt.main", "--reference-root", "/root/reference", "--seed", "0",
                        "-r", "cpu_raytracer", "-w", "16", "--height", "12", "-s", "1", "-d", "1", "-o", "cpu.png"],
                       capture_output=True, text=True, env=_env(), cwd=str(tmp_path), timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    from PIL import Image
    assert Image.open(out).size == (16, 12)          # a relative --output lands in the caller's directory


REF = os.path.join(ROOT, "baseline", "_ref")


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.skipif(not shim.is_reference_root(REF), reason="no reference checkout under baseline/_ref")
@pytest.mark.parametrize("name,flags", [("b200_texture_raytracer", ["-s", "4", "-d", "6"]),
                                        ("b200_path_tracer", ["--path-samples", "16", "-d", "6"])])
def test_shim_png_equals_factory_render(tmp_path, name, flags):
    if _gpu_count() < 1:
        pytest.skip("no CUDA device")
    W, H = 96, 72
    out = tmp_path / "shim.png"
    r = subprocess.run([sys.executable, "-m", "b200rt.main", "--seed", "0", "-r", name, "-w", str(W), "--height", str(H),
                        *flags, "-o", str(out)], capture_output=True, text=True, env=_env(), cwd=REF, timeout=900)
    assert r.returncode == 0, (r.stdout[-300:], r.stderr[-800:])
    from PIL import Image
    got = np.asarray(Image.open(out).convert("RGB"))
    # the same thing without main.py: the reference's factory, the reference's scene builder, render()
    code = (
        "import os, random, sys; sys.path.insert(0, os.getcwd()); sys.path.insert(0, %r)\n"
        "import b200rt.renderer\n"
        "from renderers.base_renderer import RendererFactory\n"
        "from scene_builders.custom_scene_builder import CustomSceneBuilder\n"
        "from core.scene import RenderSettings\n"
        "random.seed(0); b = CustomSceneBuilder(); scene = b.build_scene(); cam = b.create_camera(%d / %d)\n"
        "RendererFactory.create(%r).render(scene, cam, RenderSettings(%d, %d, %d, 6)).save(%r)\n"
    ) % (PKG, W, H, name, W, H, 4 if name == "b200_texture_raytracer" else 16, str(tmp_path / "direct.png"))
    d = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=REF, timeout=900)
    assert d.returncode == 0, d.stderr[-800:]
    want = np.asarray(Image.open(tmp_path / "direct.png").convert("RGB"))
    assert got.shape == (H, W, 3) and np.array_equal(got, want)


@pytest.mark.gpu
@pytest.mark.skipif(not shim.is_reference_root(REF), reason="no reference checkout under baseline/_ref")
def test_shim_gpus_flag_splits_the_samples(tmp_path):
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    W, H = 96, 72
    base = [sys.executable, "-m", "b200rt.main", "--seed", "0", "-r", "b200_path_tracer", "-w", str(W), "--height", str(H),
            "--path-samples", "64", "-d", "6"]
    a = subprocess.run(base + ["-o", str(tmp_path / "one.png")], capture_output=True, text=True, env=_env(), cwd=REF, timeout=900)
    b = subprocess.run(base + ["--gpus", "2", "-o", str(tmp_path / "two.png")], capture_output=True, text=True, env=_env(), cwd=REF, timeout=900)
    assert a.returncode == 0 and b.returncode == 0, (a.stderr[-500:], b.stderr[-800:])
    from PIL import Image
    one = np.asarray(Image.open(tmp_path / "one.png").convert("RGB")).astype(int)
    two = np.asarray(Image.open(tmp_path / "two.png").convert("RGB")).astype(int)
    # same global samples (counter-based RNG keyed by pixel and global sample index): only the float32 summation order
    # differs, i.e. at most the last bit of a quantised byte
    assert np.abs(one - two).max() <= 1 and (one != two).mean() < 0.01
