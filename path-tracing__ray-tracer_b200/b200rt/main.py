"""Run the reference's OWN ``main.py``, unmodified, with the B200 renderers registered.

    cd <reference checkout> && python -m b200rt.main -r b200_path_tracer -w 1920 --height 1080 --path-samples 1024 -d 8
    python -m b200rt.main --reference-root baseline/_ref -r b200_texture_raytracer            # its default image
    python -m b200rt.main --gpus 8 -r b200_path_tracer --path-samples 4096 -w 3840 --height 2160 -d 8

What the shim does, in this order (reference lines: ``main.py:11-20`` renderer imports, ``:24-44`` flags,
``:26`` ``choices=RendererFactory.list_available()``, ``:75`` ``RendererFactory.create``, ``:90`` ``render``,
``:104-108`` throughput line):

1. puts the reference checkout first on ``sys.path`` and makes it the working directory (the scene builder opens
   its textures by relative path, ``custom_scene_builder.py:77-86``);
2. imports ``b200rt.renderer`` — ``b200rt.plugin`` finds the reference's ``renderers.base_renderer`` and registers
   ``b200_path_tracer`` / ``b200_texture_raytracer`` / ``b200_raytracer`` into the REFERENCE's ``RendererFactory``,
   so they appear in ``--renderer``'s choices;
3. imports the reference's ``main`` module as it is and calls ``main.main()``.

Flags added on top of the reference's (consumed here, never seen by ``main.py``):
  ``--gpus N``            re-executes under ``torch.distributed.run`` with one rank per GPU (samples split + NCCL reduce);
                          ranks other than 0 render their share and skip saving / showing
  ``--reference-root D``  where the checkout is (default: the working directory, then ``$B200RT_REFERENCE``,
                          then ``baseline/_ref`` next to this repository)
  ``--seed S``            ``random.seed(S)`` before the scene is built (the BVH shuffle, ``core/acceleration.py:13``)
``main.py:49-54`` only gives ``--path-samples`` to a renderer literally named ``cuda_path_raytracer``; for
``-r b200_path_tracer`` the shim passes the ``--path-samples`` value as ``--samples`` so the flag means the same.
"""
from __future__ import annotations

import importlib
import os
import random
import socket
import sys
from typing import List, Optional, Tuple

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def is_reference_root(path: str) -> bool:
    return bool(path) and os.path.isfile(os.path.join(path, "main.py")) and \
        os.path.isfile(os.path.join(path, "renderers", "base_renderer.py"))


def find_reference_root(explicit: Optional[str] = None) -> str:
    cands = [explicit] if explicit else [os.getcwd(), os.environ.get("B200RT_REFERENCE", ""),
                                         os.path.join(REPO, "baseline", "_ref"), "/root/reference"]
    for c in cands:
        if c and is_reference_root(c):
            return os.path.abspath(c)
    raise SystemExit("b200rt.main: no reference checkout found (tried: %s); pass --reference-root"
                     % ", ".join(repr(c) for c in cands if c))


def split_args(argv: List[str]) -> Tuple[dict, List[str]]:
    """Takes the shim's own flags out of ``argv`` -> (options, the flags main.py will see)."""
    opts = {"gpus": 1, "reference_root": None, "seed": None}
    rest, i = [], 0
    while i < len(argv):
        a = argv[i]
        key = a.split("=", 1)[0]
        if key in ("--gpus", "--reference-root", "--seed"):
            if "=" in a:
                val = a.split("=", 1)[1]
            else:
                i += 1
                if i >= len(argv):
                    raise SystemExit(f"b200rt.main: {key} needs a value")
                val = argv[i]
            name = key[2:].replace("-", "_")
            opts[name] = val if name == "reference_root" else int(val)
        else:
            rest.append(a)
        i += 1
    return opts, rest


def _flag_value(argv: List[str], names) -> Optional[str]:
    for i, a in enumerate(argv):
        if a in names and i + 1 < len(argv):
            return argv[i + 1]
        for n in names:
            if n.startswith("--") and a.startswith(n + "="):
                return a.split("=", 1)[1]
    return None


def translate_args(argv: List[str]) -> List[str]:
    """``-r b200_path_tracer``: ``--path-samples`` (default 1024, ``main.py:43``) becomes ``--samples`` unless
    ``--samples`` was given explicitly — ``main.py:49-54`` keys that flag on the reference renderer's name."""
    name = _flag_value(argv, ("--renderer", "-r"))
    if name != "b200_path_tracer" or _flag_value(argv, ("--samples", "-s")) is not None:
        return list(argv)
    spp = _flag_value(argv, ("--path-samples",)) or "1024"
    return list(argv) + ["--samples", spp]


def torchrun_command(gpus: int, argv: List[str], port: Optional[int] = None) -> List[str]:
    if port is None:
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
    return [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={gpus}",
            "--master-addr", "127.0.0.1", "--master-port", str(port), "-m", "b200rt.main", *argv]


class _NullImage:
    """What ``render()`` hands ``main.py`` on ranks other than 0: ``save`` / ``show`` do nothing."""
    size = (0, 0)

    def save(self, *a, **k):
        pass

    def show(self, *a, **k):
        pass


def run(argv: Optional[List[str]] = None) -> int:
    opts, rest = split_args(list(sys.argv[1:] if argv is None else argv))
    root = find_reference_root(opts["reference_root"])
    if opts["gpus"] > 1 and "RANK" not in os.environ:
        env = dict(os.environ)
        pkg_parent = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        env["PYTHONPATH"] = pkg_parent + os.pathsep + env.get("PYTHONPATH", "")
        fwd = ["--reference-root", root] + (["--seed", str(opts["seed"])] if opts["seed"] is not None else []) + rest
        import subprocess
        return subprocess.call(torchrun_command(opts["gpus"], fwd), env=env, cwd=os.getcwd())

    out_flag = _flag_value(rest, ("--output", "-o"))
    if out_flag and not os.path.isabs(out_flag):             # the working directory is about to change
        out_abs = os.path.abspath(out_flag)
        rest = [out_abs if a == out_flag else a for a in rest]
    os.chdir(root)
    if root in sys.path:
        sys.path.remove(root)
    sys.path.insert(0, root)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch
        import torch.distributed as td
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        td.init_process_group("nccl")

    from . import plugin
    if plugin._ref is None:
        raise SystemExit(f"b200rt.main: {root} has no importable renderers.base_renderer")
    from . import renderer  # noqa: F401  registers the b200_* renderers into the REFERENCE's factory
    factory = plugin.RendererFactory

    if os.environ.get("B200RT_SHOW", "0") != "1":            # main.py:124-127 tries image.show(): no viewer on a server
        from PIL import Image
        Image.Image.show = lambda self, *a, **k: None
    if rank != 0:
        create = factory.create.__func__ if hasattr(factory.create, "__func__") else factory.create

        def create_quiet(cls, name, **kwargs):
            r = create(cls, name, **kwargs)
            render = r.render

            def render_rank(scene, camera, settings):
                img = render(scene, camera, settings)
                return img if img is not None else _NullImage()
            r.render = render_rank
            return r
        factory.create = classmethod(create_quiet)

    ref_main = importlib.import_module("main")               # the reference's main.py, as it is
    if not hasattr(ref_main, "main") or os.path.dirname(os.path.abspath(ref_main.__file__)) != root:
        raise SystemExit("b200rt.main: imported a `main` module that is not the reference's")
    if opts["seed"] is not None:
        random.seed(opts["seed"])
    sys.argv = [os.path.join(root, "main.py")] + translate_args(rest)
    try:
        ref_main.main()
    finally:
        if world > 1:
            import torch.distributed as td
            if td.is_initialized():
                td.destroy_process_group()
    return 0


if __name__ == "__main__":
    raise SystemExit(run())
