"""BASELINE configs[3] ("config 4"): instanced ~20M-triangle city, 3840x2160, 256 effective spp, samples sharded across the
GPUs of one box (frame indices round-robin, per-GPU SUM buffers) + ONE NCCL all-reduce of the accumulation image.
STRONG scaling: the 256 samples are divided among the ranks.  Run under torchrun for N > 1:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_config4.py [--spp 256]

Rank 0 prints one JSON line: total ms (device events, max over ranks, all-reduce inside), Mrays/s, reduce ms, and an image
checksum-of-sums (the mean radiance must not depend on N beyond fp32 summation order)."""
import argparse, json, os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--spp", type=int, default=256)
ap.add_argument("--width", type=int, default=3840)
ap.add_argument("--height", type=int, default=2160)
a = ap.parse_args()
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dist = None
if world > 1:
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from helios_b200 import abi, api, multi_gpu, scenes
s = scenes.city_scene(width=a.width, height=a.height)
ctx = api.Context(s.width, s.height, device=local)
t0 = time.time()
ctx.load_scene(s)
ctx.synchronize()
t_load = time.time() - t0
ctx.set_accum_mode(abi.ACCUM_SUM)
assert a.spp % world == 0
mine = multi_gpu.frame_indices(rank, world, a.spp // world)
pcs = [s.push_constants(f) for f in mine]
acc = torch.as_tensor(multi_gpu.DeviceArray(ctx.accum_device_ptr(), s.width * s.height * 4), device=f"cuda:{local}")
for pc in pcs[:2]:  # warm-up (also NCCL)
    ctx.render_frame(pc)
ctx.synchronize()
if dist is not None:
    multi_gpu.all_reduce_sum(acc.clone(), dist)
    torch.cuda.synchronize()
    dist.barrier()
ctx.accum_clear(); ctx.reset_counters(); ctx.synchronize(); torch.cuda.synchronize()
ctx.event_record(0)
for pc in pcs:
    ctx.render_frame(pc)
ctx.event_record(1)
ctx.synchronize()
r0 = torch.cuda.Event(enable_timing=True); r1 = torch.cuda.Event(enable_timing=True)
r0.record()
if dist is not None:
    multi_gpu.all_reduce_sum(acc, dist)
r1.record()
torch.cuda.synchronize()
ms_render, ms_reduce = ctx.event_elapsed_ms(0, 1), r0.elapsed_time(r1)
c = ctx.counters()
rays = float(c["extension_rays"] + c["shadow_rays"])
img = ctx.tonemap(1.0, abi.TONE_MAP_ACES, sample_scale=1.0 / a.spp)
mean = float(acc.view(-1, 4)[:, :3].double().mean().item()) / a.spp
if dist is not None:
    t = torch.tensor([ms_render, ms_reduce, ms_render + ms_reduce], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    r = torch.tensor([rays], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(r, op=dist.ReduceOp.SUM)
    ms_render, ms_reduce, ms_total, rays = float(t[0]), float(t[1]), float(t[2]), float(r[0])
else:
    ms_total = ms_render + ms_reduce
if rank == 0:
    print(json.dumps({"config": "city 19.8M instanced triangles (1023 instances of 32 meshes + ground), 3840x2160, depth 8", "n_gpus": world, "spp_total": a.spp,
                      "spp_per_gpu": a.spp // world, "ms_total": round(ms_total, 2), "ms_render": round(ms_render, 2), "ms_allreduce": round(ms_reduce, 3),
                      "ms_per_spp": round(ms_total / a.spp, 4), "mrays_s": round(rays / ms_total / 1e3, 1), "rays_total": rays, "mean_radiance": mean,
                      "image_mean_rgb8": float(img[..., :3].mean()), "scene_load_s": round(t_load, 2), "scaling": "strong"}))
ctx.close()
if dist is not None:
    dist.destroy_process_group()
