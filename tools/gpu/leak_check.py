"""device memory before / after many context + mesh life cycles (the stream-ordered build pool may keep a bounded amount)
    python tools/gpu/leak_check.py"""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import numpy as np
import torch
from helios_b200 import scenes, api, abi

def free_mb():
    torch.cuda.synchronize()
    return torch.cuda.mem_get_info()[0] / 2**20

s = scenes.foliage_scene(n_clusters=2000, cards_per_cluster=20, width=320, height=180, ground_grid=16, tex_size=64)
torch.zeros(1, device="cuda")
log = []
for it in range(40):
    ctx = api.Context(s.width, s.height)
    handles = ctx.load_scene(s)
    ctx.render(s, 3)
    if it % 3 == 0:
        ctx.set_option(abi.OPT_FRAMES_IN_FLIGHT, 8); ctx.render(s, 9); ctx.resize(400, 300)
    if it % 2:
        for h in handles:
            ctx.destroy_mesh(h)
    ctx.close()
    log.append(round(free_mb(), 1))
print(json.dumps({"free_mb_after_each_cycle": log[:3] + ["..."] + log[-5:], "drift_mb_cycle_5_to_40": round(log[4] - log[-1], 1)}))
# one context, many mesh create / destroy cycles
ctx = api.Context(64, 64)
m = s.meshes[1]
log = []
for it in range(60):
    h = ctx.create_mesh(m.vertices, m.indices, m.submeshes)
    ctx.destroy_mesh(h)
    log.append(round(free_mb(), 1))
ctx.close()
print(json.dumps({"mesh_cycles_free_mb": log[:3] + ["..."] + log[-3:], "drift_mb_cycle_5_to_60": round(log[4] - log[-1], 1)}))
