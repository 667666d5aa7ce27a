"""small end-to-end renders for compute-sanitizer (memcheck / racecheck / initcheck): builder, both BVH levels, any-hit,
all light types, tail kernel, graphs, debug views.  compute-sanitizer --tool memcheck python tools/gpu/sanitize.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import numpy as np
from helios_b200 import scenes, api, abi

for mk in (lambda: scenes.cornell_box(64, 48),
           lambda: scenes.foliage_scene(n_clusters=60, cards_per_cluster=12, width=64, height=36, ground_grid=8, tex_size=32),
           lambda: scenes.city_scene(n_instances=30, n_meshes=3, width=64, height=36, floors=(2, 4), detail=(1, 3)),
           lambda: scenes.terrain_scene(grid=40, n_spheres=6, sphere_level=1, width=64, height=36, textured=True),
           lambda: scenes.triangle_soup(30_000, 64, 36)):  # the two-level re-split with a few dozen treelets (k_treelets)
    s = mk()
    ctx = api.Context(s.width, s.height)
    ctx.load_scene(s)
    ctx.set_option(abi.OPT_TAIL_START, 2)
    ctx.set_option(abi.OPT_TAIL_THRESHOLD, 1 << 20)
    acc = ctx.render(s, 6)
    ctx.set_option(abi.OPT_CUDA_GRAPH, 0)
    acc2 = ctx.render(s, 6)
    assert np.array_equal(acc, acc2)
    ctx.tonemap(1.0, abi.TONE_MAP_ACES)
    ctx.trace_primary_ids(s.push_constants(1))
    ctx.render_output_buffer(s.push_constants(1), abi.OUTPUT_BUFFER_NORMALS)
    ctx.gather_debug_rays(s.push_constants(1, pixel_coord=(s.width // 2, s.height // 2)), 16)
    ctx.resize(48, 32)
    ctx.close()
    print(s.name, "ok", float(acc[..., :3].mean()))
