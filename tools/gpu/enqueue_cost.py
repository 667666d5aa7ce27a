"""host-side cost of enqueuing a frame vs the frame's GPU time (is the CPU the limit at small extents?)
    python tools/gpu/enqueue_cost.py"""
import sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from helios_b200 import scenes, api

import os
for w in (256, 512, 1024):
    s = scenes.cornell_box(w, w)
    ctx = api.Context(w, w)
    ctx.load_scene(s)
    if os.environ.get('HL_GRAPH') is not None: ctx.set_option(6, int(os.environ['HL_GRAPH']))
    pcs = [s.push_constants(f) for f in range(1, 301)]
    for pc in pcs[:20]:
        ctx.render_frame(pc)
    ctx.synchronize()
    ctx.reset_counters()
    ctx.event_record(0)
    t0 = time.perf_counter()
    for pc in pcs[20:]:
        ctx.render_frame(pc)
    t1 = time.perf_counter()
    ctx.event_record(1)
    gpu_ms = ctx.event_elapsed_ms(0, 1) / 280
    c = ctx.counters()
    rays = float(c["extension_rays"] + c["shadow_rays"]) / 280
    print(json.dumps({"extent": w, "host_enqueue_us_per_frame": round((t1 - t0) / 280 * 1e6, 1), "gpu_ms_per_frame": round(gpu_ms, 4), "mrays_s": round(rays / gpu_ms / 1e3, 1)}))
    ctx.close()
