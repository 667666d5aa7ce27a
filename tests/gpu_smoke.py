"""quick GPU bring-up script (not a pytest): prints parity and timing numbers"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from helios_b200 import scenes, api
from oracle import oracle

def ids_cmp(g, r):
    return [int((a.view(np.uint32) != b.view(np.uint32)).sum()) for a, b in zip(g, r)]

s = scenes.cornell_box(128, 128)
ctx = api.Context(s.width, s.height); ctx.load_scene(s)
o = oracle.OracleScene(s)
pc = s.push_constants(1)
print("cornell id mismatches", ids_cmp(ctx.trace_primary_ids(pc), o.trace_primary_ids(pc)))
a = ctx.render(s, 9); b = o.render(9)
d = np.abs(a - b)[..., :3]
print("cornell radiance maxdiff", d.max(), "px>1e-4", int((d.max(-1) > 1e-4).sum()), "mean", a[..., :3].mean(), b[..., :3].mean(), ctx.counters(), o.counters)
ctx.close()

for n in (100_000, 1_000_000):
    s = scenes.triangle_soup(n, 1920, 1080)
    ctx = api.Context(s.width, s.height); t = time.time(); h = ctx.load_scene(s); print("soup", n, "load s", time.time() - t, ctx.mesh_build_stats(h[0]))
    pc = s.push_constants(1)
    g = ctx.trace_primary_ids(pc)
    ctx.set_profiling(True)
    ctx.render_frame(s.push_constants(0)); ctx.render_frame(s.push_constants(1)); ctx.synchronize()
    print("  frame counters", ctx.counters())
    if n <= 100_000:
        o = oracle.OracleScene(s)
        print("  id mismatches", ids_cmp(g, o.trace_primary_ids(pc)))
    ctx.close()

s = scenes.terrain_scene()
print("terrain tris", s.num_triangles)
ctx = api.Context(s.width, s.height); t = time.time(); h = ctx.load_scene(s); print("terrain load s", time.time() - t, ctx.mesh_build_stats(h[0]))
ctx.set_profiling(True)
for f in range(4):
    ctx.render_frame(s.push_constants(f)); ctx.synchronize()
    c = ctx.counters(); print("  frame", f, c)
img = ctx.tonemap()
from PIL import Image
Path("gpurun_out").mkdir(exist_ok=True)
Image.fromarray(img[..., :3]).save("gpurun_out/terrain.png")
ctx.close()
