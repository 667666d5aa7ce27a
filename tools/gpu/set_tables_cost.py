"""cost of hl_scene_set_tables (table upload + TLAS rebuild) = what an interactive transform edit pays per change
(the reference rebuilds its TLAS too, renderer.cpp:147-168).  python tools/gpu/set_tables_cost.py"""
import sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import numpy as np
from helios_b200 import scenes, api

for n in (30, 255, 1023):
    s = scenes.city_scene(n_instances=n, n_meshes=8, width=640, height=360, floors=(2, 4), detail=(1, 3))
    ctx = api.Context(s.width, s.height)
    handles = ctx.load_scene(s)
    mh = [handles[int(i["mesh_index"])] for i in s.instances]
    inst = s.instances.copy()
    ctx.synchronize()
    ts = []
    for k in range(20):
        inst["model_matrix"][:, 12] += 0.001  # nudge every instance along x (column-major translation)
        t0 = time.perf_counter()
        ctx.set_tables(s.materials, inst, mh, s.submesh_info, s.lights)
        ctx.synchronize()
        ts.append(time.perf_counter() - t0)
    ctx.render_frame(s.push_constants(1))
    ctx.synchronize()
    print(json.dumps({"instances": len(inst), "set_tables_ms_median": round(1e3 * float(np.median(ts)), 3), "min": round(1e3 * min(ts), 3)}))
    ctx.close()
