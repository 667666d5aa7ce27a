"""which pixel / frame of a full-size config differs between two BVH builds (HL_OPT_SAH_CLUSTER a / b)?
    python tools/gpu/find_tree_diff.py city 0 2 [frames]"""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import numpy as np
from helios_b200 import api, scenes, abi

name, ca, cb = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
frames = int(sys.argv[4]) if len(sys.argv) > 4 else 3
s = {"city": scenes.city_scene, "foliage": scenes.foliage_scene, "terrain": scenes.terrain_scene}[name]()
ctxs = []
for c in (ca, cb):
    ctx = api.Context(s.width, s.height)
    ctx.set_option(abi.OPT_SAH_CLUSTER, c)
    ctx.load_scene(s)
    ctxs.append(ctx)
for f in range(frames):
    imgs, cnt = [], []
    for ctx in ctxs:
        ctx.accum_clear(); ctx.reset_counters()
        ctx.set_accum_mode(abi.ACCUM_SUM)
        ctx.render_frame(s.push_constants(f))
        imgs.append(ctx.read_accum()); c = ctx.counters(); cnt.append((int(c["extension_rays"]), int(c["shadow_rays"])))
    d = np.argwhere((imgs[0] != imgs[1]).any(-1))
    print(json.dumps({"frame": f, "rays": cnt, "differing_pixels": d.tolist()[:10], "values": [[imgs[0][y, x].tolist(), imgs[1][y, x].tolist()] for y, x in d[:10]]}), flush=True)
