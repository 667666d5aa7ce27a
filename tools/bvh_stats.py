"""Traversal statistics of the builder's BVH, measured with the kernel-logic emulator (tests/emul) on the CPU:
node visits and leaf triangle tests per extension ray, per bounce.  Used to judge builder changes (SAH
refinement, collapse policy) before spending GPU minutes.  Debug tooling; not imported by the package.

    python tools/bvh_stats.py terrain --width 480 --height 270
"""
from __future__ import annotations

import argparse
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from helios_b200 import scenes, sky  # noqa: E402
from tests.emul import emul  # noqa: E402


def make_scene(name, w, h):
    if name == "terrain":
        return scenes.terrain_scene(width=w, height=h)
    if name == "terrain_small":
        return scenes.terrain_scene(grid=200, n_spheres=8, sphere_level=2, width=w, height=h)
    if name == "foliage":
        return scenes.foliage_scene(n_clusters=5000, width=w, height=h)
    if name == "city":
        return scenes.city_scene(n_instances=255, n_meshes=8, width=w, height=h)
    if name == "soup":
        return scenes.triangle_soup(1_000_000, width=w, height=h)
    if name == "cornell":
        return scenes.cornell_box(w, h)
    raise SystemExit(f"unknown scene {name}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("scene")
    ap.add_argument("--width", type=int, default=480)
    ap.add_argument("--height", type=int, default=270)
    ap.add_argument("--depths", type=int, default=4)
    a = ap.parse_args()
    s = make_scene(a.scene, a.width, a.height)
    faces = None
    if s.env_cube is None and getattr(s, "sun_direction", None) is not None:
        faces = emul.sky_bake(sky.sky_coefficients(s.sun_direction), s.sun_direction, 64)
    import os
    if os.environ.get("HL_EMUL_SAH_TOP"):
        emul.lib().em_set_sah_top(C.c_int(int(os.environ["HL_EMUL_SAH_TOP"])))
    t0 = time.time()
    e = emul.EmulScene(s, sky_faces=faces)
    t1 = time.time()
    n = s.width * s.height
    log = np.zeros((16, n), np.uint32)
    emul.lib().em_set_node_log(log.ctypes.data_as(C.c_void_p), C.c_size_t(n))
    accum = np.zeros((s.height, s.width, 4), np.float32)
    e.render_frame(s.push_constants(1), accum)
    emul.lib().em_set_node_log(None, C.c_size_t(0))
    print(f"{a.scene}: {s.num_triangles} triangles, build {t1 - t0:.1f}s, mesh0 stats (tris, wide nodes, binary) = {e.mesh_stats(0)}")
    emul.lib().em_mesh_sah.restype = C.c_float
    print("  SAH cost per mesh (DP, per unit root area):", [round(float(emul.lib().em_mesh_sah(e.h, C.c_int(k))), 2) for k in range(len(s.meshes))])
    tot_n = tot_l = tot_r = 0
    for d in range(a.depths):
        v = log[d]
        m = v != 0
        if not m.any():
            break
        nodes, leaves = (v[m] & 0xFFFF).astype(np.float64), (v[m] >> 16).astype(np.float64)
        tot_n += nodes.sum(); tot_l += leaves.sum(); tot_r += m.sum()
        print(f"  depth {d}: rays {m.sum():8d}  nodes/ray {nodes.mean():6.2f} (p99 {np.percentile(nodes, 99):5.0f})  tri tests/ray {leaves.mean():6.2f}")
    print(f"  all: nodes/ray {tot_n / tot_r:.2f}  tri tests/ray {tot_l / tot_r:.2f}")


if __name__ == "__main__":
    main()
