"""Launched by torchrun from the tests: renders 8 spp split over the ranks (one NCCL / gloo reduce) and, on rank 0,
the same 8 spp alone; writes "ok <max rel diff>" or "FAIL ..." to argv[2].  argv[1] = backend (nccl)."""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "path-tracing__ray-tracer_b200")):
    sys.path.insert(0, p)

import numpy as np
import torch
import torch.distributed as td

from b200rt import renderer
from b200rt.cornell import CustomSceneBuilder
from b200rt.scene_api import RenderSettings

backend, out_path = sys.argv[1], sys.argv[2]
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
td.init_process_group(backend, device_id=dev)
random.seed(0)
b = CustomSceneBuilder(texture_dir=False)
scene = b.build_scene()
W, H, SPP, D = 320, 180, 8, 8
cam = b.create_camera(W / H)
rs = renderer.B200PathTracer(precision="f32", seed=5, device=dev)
split, cnt = rs.render_accum(scene, cam, RenderSettings(W, H, SPP, D))
img_split = rs.render(scene, cam, RenderSettings(W, H, SPP, D))       # fused reduce + resolve over peer memory (when available)
msg = "ok"
if rank == 0:
    r1 = renderer.B200PathTracer(precision="f32", seed=5, device=dev, distributed=False)
    one, cnt1 = r1.render_accum(scene, cam, RenderSettings(W, H, SPP, D))
    img_one = r1.render(scene, cam, RenderSettings(W, H, SPP, D))
    a, c = split[..., :3].astype(np.float64), one[..., :3].astype(np.float64)
    rel = np.abs(a - c) / np.maximum(np.abs(c), 1e-3)
    dimg = np.abs(np.asarray(img_split).astype(int) - np.asarray(img_one).astype(int))
    ok = np.allclose(a, c, rtol=1e-5, atol=1e-6) and int(cnt1[0]) == W * H * SPP and dimg.max() <= 1
    msg = ("ok %.3e image diff %d fused %s" % (rel.max(), dimg.max(), rs._symm is not None)) if ok else \
          ("FAIL max rel %.3e, paths %d, image diff %d" % (rel.max(), int(cnt1[0]), dimg.max()))
    with open(out_path, "w") as f:
        f.write(msg)
td.barrier()
td.destroy_process_group()
sys.exit(0 if msg.startswith("ok") else 1)
