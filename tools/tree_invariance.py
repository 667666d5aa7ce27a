"""Do two differently built BVHs (HL_OPT_SAH_CLUSTER a / b) return the same closest hit for every ray?
Secondary-like rays (origins exactly ON the scene's surfaces, random directions, tmin 1e-4 as in rchit:523) are
traced through both trees; every disagreement is printed with the oracle's brute-force answer for that ray.
    python tools/tree_invariance.py city 0 2 [million_rays]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from helios_b200 import api, scenes

name, ca, cb = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
total = int(float(sys.argv[4]) * 1e6) if len(sys.argv) > 4 else 64_000_000
s = {"city": lambda: scenes.city_scene(width=960, height=540), "city_small": lambda: scenes.city_scene(n_instances=40, n_meshes=4, width=960, height=540, floors=(2, 5), detail=(1, 3)),
     "terrain": lambda: scenes.terrain_scene(width=960, height=540), "foliage": lambda: scenes.foliage_scene(width=960, height=540)}[name]()
ctxs = []
for c in (ca, cb):
    ctx = api.Context(s.width, s.height)
    ctx.set_option(4, c)
    ctx.load_scene(s)
    ctxs.append(ctx)
# surface points: first hits of random rays from the camera position
rng = np.random.default_rng(1)
n0 = 2_000_000
d0 = rng.normal(size=(n0, 3)).astype(np.float32)
d0 /= np.linalg.norm(d0, axis=1, keepdims=True).astype(np.float32)
r0 = np.empty((n0, 8), np.float32)
r0[:, 0:3], r0[:, 3], r0[:, 4:7], r0[:, 7] = np.asarray(s.camera.position, np.float32), 1e-3, d0, 1e4
h0 = ctxs[0].trace_rays(r0)
ok = h0.view(np.uint32)[:, 3] != 0xFFFFFFFF
P = (r0[ok, 0:3] + d0[ok] * h0[ok, 0:1]).astype(np.float32)
print(f"{name}: {s.num_triangles} triangles, {len(P)} surface points", flush=True)
chunk, bad, done = 8_000_000, [], 0
while done < total:
    idx = rng.integers(0, len(P), chunk)
    dirs = rng.normal(size=(chunk, 3)).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True).astype(np.float32)
    rays = np.empty((chunk, 8), np.float32)
    rays[:, 0:3], rays[:, 3], rays[:, 4:7], rays[:, 7] = P[idx], 1e-4, dirs, 1e4
    ha, hb = ctxs[0].trace_rays(rays), ctxs[1].trace_rays(rays)
    m = np.nonzero((ha.view(np.uint32) != hb.view(np.uint32)).any(axis=1))[0]
    for k in m[:20]:
        bad.append((rays[k].copy(), ha[k].copy(), hb[k].copy()))
    done += chunk
    print(f"  {done/1e6:.0f}M rays, {len(m)} disagreements in this chunk", flush=True)
if bad:
    from oracle import oracle as O
    orc = O.OracleScene(s, brute_force=True)
    R = np.stack([b[0] for b in bad])
    ho = orc.trace_rays(R)
    for (r, a, b), h in zip(bad, ho):
        print("ray", r.tolist())
        print("   tree a:", a.view(np.uint32)[3:6].tolist(), a[:3].tolist(), " tree b:", b.view(np.uint32)[3:6].tolist(), b[:3].tolist(), " brute force:", h.view(np.uint32)[3:6].tolist(), h[:3].tolist())
for c in ctxs:
    c.close()
