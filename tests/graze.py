"""Adversarial rays for the conservative node test (hl_bvh.h intersect_children): large wall triangles among clutter, and
rays that cross a wall's plane within a hair of one of its edges at grazing angles (0.01 .. 10 degrees) from 3 .. 600 units
away.  The fp32 triangle test accepts some of them although the ray passes the triangle's box at a distance that grows
with 1 / cos(theta): a traversal whose box test is not conservative enough loses those hits in one tree and not in
another.  With the node-test slack at 2^-19 D this set loses 44 of 2e6 hits (LBVH tree), at 2^-18: 3, from 2^-17 on: 0;
the library uses 2^-16.  Used by tests/test_emul_parity.py (device logic on the CPU vs brute force) and
tests/test_gpu_edges.py."""
import numpy as np

from helios_b200 import abi, scenes

def graze_scene(seed=3, n_walls=40, n_clutter=3000):
    """large axis-aligned and tilted wall quads (two triangles each, sharing a diagonal) + small clutter triangles"""
    rng = np.random.default_rng(seed)
    s = scenes.triangle_soup(n_clutter, 64, 36, seed=seed)
    m = s.meshes[0]
    pos = m.vertices["position"][:, :3].copy() * 200.0 - 100.0          # clutter spread over [-100, 100]^3
    quads = []
    for k in range(n_walls):
        c = rng.uniform(-80, 80, 3); size = rng.uniform(5, 60)
        if k % 2 == 0:   # axis-aligned wall
            ax = k % 3; u = np.roll(np.array([0, 1.0, 0]), ax); v = np.roll(np.array([0, 0, 1.0]), ax)
        else:
            u = rng.normal(size=3); u /= np.linalg.norm(u); v = np.cross(u, rng.normal(size=3)); v /= np.linalg.norm(v)
        p = [c - u*size - v*size, c + u*size - v*size, c + u*size + v*size, c - u*size + v*size]
        quads += [p[0], p[1], p[2], p[0], p[2], p[3]]
    allpos = np.concatenate([pos, np.array(quads)]).astype(np.float32)
    n = len(allpos)
    v = np.zeros(n, abi.VERTEX); v["position"][:, :3] = allpos; v["normal"][:, 1] = 1; v["tangent"][:, 0] = 1; v["bitangent"][:, 2] = 1
    idx = np.arange(n, dtype=np.uint32)
    subs = np.zeros(1, abi.SUBMESH); subs[0] = (0, n, n, 1)
    s.meshes[0] = scenes.MeshData(v, idx, subs, [0])
    s.submesh_info = [scenes.submesh_table(s.meshes[0])]
    return s, np.array(quads, np.float32).reshape(-1, 3, 3)

def graze_rays(tris, n, seed=5):
    """rays that cross a wall triangle's plane within a hair of one of its edges, at grazing angles, from far away"""
    rng = np.random.default_rng(seed)
    t = tris[rng.integers(0, len(tris), n)].astype(np.float64)
    e = rng.integers(0, 3, n)
    a = t[np.arange(n), e]; b = t[np.arange(n), (e + 1) % 3]; c = t[np.arange(n), (e + 2) % 3]
    w = rng.random(n)[:, None]
    edge_pt = a * (1 - w) + b * w
    inward = (c - edge_pt); inward /= np.linalg.norm(inward, axis=1, keepdims=True)
    target = edge_pt + inward * (rng.normal(size=(n, 1)) * 2e-4)          # +- a hair across the edge
    nrm = np.cross(b - a, c - a); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    tang = rng.normal(size=(n, 3)); tang -= nrm * (tang * nrm).sum(1, keepdims=True); tang /= np.linalg.norm(tang, axis=1, keepdims=True)
    ang = np.radians(10 ** rng.uniform(-2, 1, n))[:, None]               # 0.01 .. 10 degrees off the plane
    d = tang * np.cos(ang) + nrm * np.sin(ang) * rng.choice([-1, 1], (n, 1))
    dist = (10 ** rng.uniform(0.5, 2.8, n))[:, None]                       # 3 .. 600 units away
    o = target - d * dist
    rays = np.concatenate([o, np.full((n, 1), 1e-3), d, np.full((n, 1), 1e4)], 1).astype(np.float32)
    return rays
