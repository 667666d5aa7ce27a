"""device time of the BVH build for one soup size, three builds in a row (first = cold pool): python tools/build_time.py N"""
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from helios_b200 import api, scenes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
s = scenes.triangle_soup(n, 64, 64)
ctx = api.Context(64, 64)
if os.environ.get('HL_SAH_CLUSTER') is not None:
    ctx.set_option(4, int(os.environ['HL_SAH_CLUSTER']))  # HL_OPT_SAH_CLUSTER
m = s.meshes[0]
for k in range(3):
    h = ctx.create_mesh(m.vertices, m.indices, m.submeshes)
    st = ctx.mesh_build_stats(h)
    print(n, "build", k, "ms", round(float(st["ms_build"]), 3), "nodes", int(st["wide_nodes"]), "sah", round(float(st["sah_cost"]), 2))
    ctx.destroy_mesh(h)
ctx.close()
