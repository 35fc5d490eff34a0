"""BASELINE config 5: 4K (3840x2160) Cornell box, 4096 spp, depth 8, spp split across the ranks + one NCCL
reduce.  Launch with torchrun (one rank per GPU); rank 0 prints one JSON line."""
import json, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "path-tracing__ray-tracer_b200"))
import torch, torch.distributed as td
from b200rt import dist, renderer
from b200rt.cornell import CustomSceneBuilder
from b200rt.scene_api import RenderSettings

W, H, SPP, D = 3840, 2160, int(os.environ.get("C5_SPP", "4096")), 8
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    td.init_process_group("nccl", device_id=dev)
random.seed(0); b = CustomSceneBuilder(texture_dir=False); scene = b.build_scene(); cam = b.create_camera(W / H)
r = renderer.B200PathTracer(precision="f32", device=dev)
st = r.prepare(scene, cam, RenderSettings(W, H, SPP, D))
def frame():
    st["accum"].zero_(); r.accumulate(st); dist.reduce_to_root(st["accum"])
    if rank == 0: r.resolve(st)
r_small = r.prepare(scene, cam, RenderSettings(W, H, 8 * world, D)); r.accumulate(r_small); torch.cuda.synchronize()
if world > 1: td.barrier(device_ids=[local])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); frame(); e1.record(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev)
cnt = st["counters"].clone()
if world > 1:
    td.all_reduce(t, op=td.ReduceOp.MAX); td.all_reduce(cnt, op=td.ReduceOp.SUM)
t0 = time.perf_counter(); img = r.render(scene, cam, RenderSettings(W, H, SPP, D)); wall = time.perf_counter() - t0
if rank == 0:
    c = cnt.cpu().numpy(); s = float(t.item())
    img.save(os.path.join(ROOT, "gpurun_out", f"config5_{world}gpu.png")) if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else None
    print(json.dumps({"config": f"C5 4K Cornell {SPP} spp depth 8", "n_gpus": world, "frame_s": s, "paths": int(c[0]),
                      "Mpaths_per_s": c[0] / s / 1e6, "Mrays_per_s": (c[1] + c[2]) / s / 1e6, "e2e_render_s": wall,
                      "spp_per_gpu": st["spp_local"], "wave": st["wave"]}))
if world > 1: td.destroy_process_group()
