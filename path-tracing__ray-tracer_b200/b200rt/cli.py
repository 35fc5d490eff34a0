"""Command line with the flags of the reference's ``main.py:24-44`` (``--renderer --scene --width --height
--samples --depth --output --path-samples``) plus ``--seed``, ``--precision`` and ``--stats``.

    python -m b200rt.cli -r b200_path_tracer -w 1920 --height 1080 --path-samples 1024 -d 8 -o cornell.png

Unlike the reference, the throughput line reports COUNTED rays (closest-hit + shadow rays actually traced),
not the nominal ``W*H*spp*depth`` of ``main.py:104-108``.  Run under ``torchrun`` to split the samples over GPUs.
"""
from __future__ import annotations

import argparse
import os
import random
import time


def main(argv=None) -> int:
    from . import renderer  # noqa: F401  (registers the renderers)
    from .cornell import CustomSceneBuilder
    from .plugin import RendererFactory
    from .scene_api import RenderSettings

    ap = argparse.ArgumentParser(description="b200rt renderer CLI (flags of the reference main.py)")
    ap.add_argument("--renderer", "-r", choices=[n for n in RendererFactory.list_available() if n.startswith("b200")],
                    default="b200_texture_raytracer")
    ap.add_argument("--scene", choices=["original", "custom"], default="custom")
    ap.add_argument("--width", "-w", type=int, default=2000)
    ap.add_argument("--height", type=int, default=1500)
    ap.add_argument("--samples", "-s", type=int, default=25)
    ap.add_argument("--depth", "-d", type=int, default=16)
    ap.add_argument("--output", "-o", default="output.png")
    ap.add_argument("--path-samples", type=int, default=1024)
    ap.add_argument("--seed", type=int, default=0, help="scene BVH shuffle seed and path-tracer RNG seed")
    ap.add_argument("--precision", choices=["f32", "f64"], default="f32")
    ap.add_argument("--stats", action="store_true")
    args = ap.parse_args(argv)

    rank = int(os.environ.get("RANK", "0"))
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch
        import torch.distributed as td
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        td.init_process_group("nccl")
    spp = args.path_samples if args.renderer == "b200_path_tracer" else args.samples       # main.py:49-54
    settings = RenderSettings(args.width, args.height, spp, args.depth)
    random.seed(args.seed)
    builder = CustomSceneBuilder()
    scene = builder.build_scene()
    camera = builder.create_camera(args.width / args.height)
    kwargs = dict(precision=args.precision)
    if args.renderer == "b200_path_tracer":
        kwargs["seed"] = args.seed
    r = RendererFactory.create(args.renderer, **kwargs)
    t0 = time.time()
    image = r.render(scene, camera, settings)
    elapsed = time.time() - t0
    if rank == 0 and image is not None:
        image.save(args.output)
        print(f"{args.renderer}: {args.width}x{args.height}, {spp} spp, depth {args.depth}: {elapsed:.3f} s -> {args.output}")
        st = r.last_stats
        if "closest_rays" in st:
            rays = st["closest_rays"] + st["shadow_rays"]
            print(f"  kernels {st['kernel_s'] * 1e3:.1f} ms, {st['paths'] / st['kernel_s'] / 1e6:.0f} Mpaths/s, "
                  f"{rays / st['kernel_s'] / 1e6:.0f} Mrays/s counted ({rays / max(1, st['paths']):.2f} rays/path)")
        if args.stats:
            print("  " + ", ".join(f"{k}={v}" for k, v in st.items()))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
