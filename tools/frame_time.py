"""ms/frame, Mrays/s and per-stage times of one BASELINE config on the GPU, plus an image checksum (so that
tuning variants can be checked for identical results).  python tools/frame_time.py [terrain|foliage|city|cornell]"""
import sys, json, hashlib, os
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from helios_b200 import scenes, api

name = sys.argv[1] if len(sys.argv) > 1 else "terrain"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 32
s = {"terrain": lambda: scenes.terrain_scene(), "foliage": lambda: scenes.foliage_scene(), "city": lambda: scenes.city_scene(),
     "foliage_small": lambda: scenes.foliage_scene(n_clusters=5000), "city_1080": lambda: scenes.city_scene(width=1920, height=1080),
     "cornell": lambda: scenes.cornell_box(1024, 1024)}[name]()
ctx = api.Context(s.width, s.height)
if os.environ.get('HL_SAH_CLUSTER') is not None:
    ctx.set_option(4, int(os.environ['HL_SAH_CLUSTER']))  # HL_OPT_SAH_CLUSTER
handles = ctx.load_scene(s)
ctx.set_option(3, int(os.environ.get('HL_PIPELINE', '1')))  # HL_OPT_PIPELINE
if os.environ.get('HL_TAIL_THRESHOLD') is not None: ctx.set_option(1, int(os.environ['HL_TAIL_THRESHOLD']))
if os.environ.get('HL_GRAPH') is not None: ctx.set_option(6, int(os.environ['HL_GRAPH']))
if os.environ.get('HL_SLOTS') is not None: ctx.set_option(5, int(os.environ['HL_SLOTS']))
if os.environ.get('HL_TAIL_START') is not None: ctx.set_option(2, int(os.environ['HL_TAIL_START']))
pcs = [s.push_constants(f) for f in range(1, frames + 5)]
ctx.accum_clear()
for pc in pcs[:4]: ctx.render_frame(pc)
ctx.reset_counters(); ctx.synchronize()
ctx.event_record(0)
for pc in pcs[4:]: ctx.render_frame(pc)
ctx.event_record(1)
ms = ctx.event_elapsed_ms(0, 1) / frames
c = ctx.counters()
rays = float(c["extension_rays"] + c["shadow_rays"]) / frames
acc = ctx.read_accum()
out = {"scene": name, "pipeline": int(os.environ.get('HL_PIPELINE', '1')), "tris": int(s.num_triangles), "sah_cluster": os.environ.get('HL_SAH_CLUSTER', 'default'), "build_ms": [round(float(ctx.mesh_build_stats(h)["ms_build"]), 3) for h in handles[:3]], "sah": [round(float(ctx.mesh_build_stats(h)["sah_cost"]), 2) for h in handles[:3]], "ms_per_frame": round(ms, 4), "mrays_s": round(rays / ms / 1e3, 1), "rays_per_frame": rays,
       "sha": hashlib.sha256(acc.tobytes()).hexdigest()[:16]}
ctx.set_profiling(True)
st = {k: 0.0 for k in ("ms_generate", "ms_extend", "ms_shade", "ms_connect", "ms_resolve", "ms_frame")}
for pc in pcs[4:12]:
    ctx.render_frame(pc)
    c = ctx.counters()
    for k in st: st[k] += float(c[k]) / 8
out["stages"] = {k: round(v, 4) for k, v in st.items()}
print(json.dumps(out))
ctx.close()
