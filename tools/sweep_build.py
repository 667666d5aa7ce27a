"""BASELINE configs[4] ("config 5" in SURVEY.md §8d): BVH build + primary-ray hit-ID sweep over uniform triangle soups.
Per size: GPU build time (second build = warm allocator), primary Mrays/s (1920x1080 pinhole, device only), and hit
parity against the CPU oracle (instance, geometry, primitive AND t,u,v bit patterns).  One JSON line per size.

    python tools/sweep_build.py [--sizes 100000,300000,...] [--no-oracle-above N] [--out profiles/r01_config5_sweep.jsonl]
"""
import argparse, json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from helios_b200 import api, scenes

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="100000,300000,1000000,3000000,10000000,20000000,50000000")
ap.add_argument("--no-oracle-above", type=int, default=50_000_000)
ap.add_argument("--out", default="")
a = ap.parse_args()
out = open(a.out, "w") if a.out else None
for n in [int(x) for x in a.sizes.split(",")]:
    t0 = time.time()
    s = scenes.triangle_soup(n)  # 1920x1080, seed = n
    t_gen = time.time() - t0
    ctx = api.Context(s.width, s.height)
    m = s.meshes[0]
    h = ctx.create_mesh(m.vertices, m.indices, m.submeshes)
    cold = float(ctx.mesh_build_stats(h)["ms_build"])
    ctx.destroy_mesh(h)
    t0 = time.time()
    handles = ctx.load_scene(s)
    ctx.synchronize()
    t_load = time.time() - t0
    st = ctx.mesh_build_stats(handles[0])
    pc = s.push_constants(1)
    ctx.trace_primary_device_only(pc)
    ctx.synchronize()
    reps = 10
    ctx.event_record(0)
    for _ in range(reps):
        ctx.trace_primary_device_only(pc)
    ctx.event_record(1)
    ms = ctx.event_elapsed_ms(0, 1) / reps
    rec = {"triangles": n, "build_ms": round(float(st["ms_build"]), 3), "build_ms_first": round(cold, 3), "build_mtris_per_s": round(n / float(st["ms_build"]) / 1e3, 1),
           "wide_nodes": int(st["wide_nodes"]), "bvh_mbytes": round((int(st["bytes_nodes"]) + int(st["bytes_triangles"])) / 1e6, 1),
           "primary_ms": round(ms, 4), "primary_mrays_s": round(s.width * s.height / ms / 1e3, 1), "host_generate_s": round(t_gen, 1), "upload_and_build_s": round(t_load, 2)}
    g = ctx.trace_primary_ids(pc)
    rec["hit_fraction"] = round(float((g[0] != 0xFFFFFFFF).mean()), 4)
    ctx.close()
    if n <= a.no_oracle_above:
        from oracle import oracle
        t0 = time.time()
        o = oracle.OracleScene(s)
        r = o.trace_primary_ids(pc)
        rec["oracle_s"] = round(time.time() - t0, 1)
        ids_bad = (g[0] != r[0]) | (g[1] != r[1]) | (g[2] != r[2])
        tuv_bad = np.zeros_like(ids_bad)
        for k in (3, 4, 5):
            tuv_bad |= g[k].view(np.uint32) != r[k].view(np.uint32)
        rec["id_mismatches"] = int(ids_bad.sum())
        rec["tuv_bit_mismatches"] = int((tuv_bad & ~ids_bad).sum())
        rec["rays_compared"] = int(ids_bad.size)
        del o
    line = json.dumps(rec)
    print(line, flush=True)
    if out:
        out.write(line + "\n"); out.flush()
    del s, m, g
