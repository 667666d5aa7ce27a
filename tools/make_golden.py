#!/usr/bin/env python3
"""Generates tests/golden/oracle_golden.json from the CPU oracle (the reference ships no
golden vectors, so these pin the oracle against regressions; integer entries are exact, float entries are
stored as IEEE bit patterns and compared exactly or to 1e-6 where libm transcendentals are involved)."""
import ctypes as C
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from helios_b200 import scenes  # noqa: E402
from oracle import oracle  # noqa: E402

L = oracle.lib()


def p(a):
    return a.ctypes.data_as(C.c_void_p)


def bits(a):
    return [int(x) for x in np.asarray(a, np.float32).view(np.uint32).ravel()]


g = {}
# RNG
res, st = np.zeros(8, np.uint32), np.zeros(16, np.uint32)
L.or_rng_sequence(C.c_uint32(1), C.c_uint32(2), C.c_uint32(8), p(res), p(st))
g["xoroshiro64star_from_1_2"] = {"results": res.tolist(), "states": st.tolist()}
g["wang_hash"] = {str(s): int(L.or_rng_hash(C.c_uint32(s))) for s in (0, 1, 61, 0xFFFFFFFF)}
o2 = np.zeros(2, np.uint32)
g["rng_init"] = {}
for x, y, f in ((0, 0, 0), (1919, 1079, 63), (5, 7, 1)):
    L.or_rng_init(C.c_uint32(x), C.c_uint32(y), C.c_uint32(f), p(o2))
    g["rng_init"][f"{x},{y},{f}"] = o2.tolist()
fl = np.zeros(8, np.float32)
L.or_next_floats(C.c_uint32(o2[0]), C.c_uint32(o2[1]), C.c_uint32(8), p(fl))
g["next_float_bits_after_init_5_7_1"] = bits(fl)
# camera
cam = {}
for ap in (0.0, 0.1):
    s = scenes.cornell_box(512, 512, aperture_radius=ap)
    pc = np.ascontiguousarray(s.push_constants(3))
    for px, py in ((0, 0), (511, 511), (200, 300)):
        out = np.zeros(6, np.float32)
        L.or_generate_ray(p(pc), C.c_uint32(px), C.c_uint32(py), p(out))
        cam[f"ap{ap}_{px}_{py}"] = bits(out)
g["camera_rays_cornell512_frame3"] = cam
# BRDF
brdf = []
tuples = [
    ((0, 1, 0), (0, 1, 0), (0, 1, 0), 0.5, 0.0, (0.8, 0.8, 0.8)),
    ((0, 1, 0), (0.6, 0.8, 0), (-0.6, 0.8, 0), 0.2, 1.0, (0.9, 0.6, 0.2)),
    ((0, 1, 0), (0.0, 0.1, 0.995), (0.3, 0.9, -0.316), 0.05, 0.0, (0.1, 0.2, 0.3)),
    ((0, 0, 1), (0.1, 0.2, 0.97), (0.5, -0.5, 0.707), 1.0, 0.5, (1.0, 1.0, 1.0)),
    ((0, 1, 0), (0.7, -0.1, 0.7), (0.0, 1.0, 0.0), 0.3, 0.0, (0.5, 0.5, 0.5)),
    ((0.577, 0.577, 0.577), (0, 1, 0), (1, 0, 0), 0.7, 0.2, (0.2, 0.7, 0.4)),
]
for n, wo, wi, r, m, alb in tuples:
    out = np.zeros(4, np.float32)
    a = [np.asarray(v, np.float32) for v in (n, wo, wi, alb)]
    L.or_evaluate_uber(p(a[0]), p(a[1]), p(a[2]), C.c_float(r), C.c_float(m), p(a[3]), p(out))
    brdf.append({"in": [list(map(float, n)), list(map(float, wo)), list(map(float, wi)), r, m, list(map(float, alb))], "out_bits": bits(out)})
g["evaluate_uber_pdf"] = brdf
smp = []
for sx, sy in ((1, 2), (123456789, 987654321), (0xDEADBEEF, 0x12345678), (42, 4242)):
    out = np.zeros(7, np.float32)
    n, wo, alb = (np.asarray(v, np.float32) for v in ((0, 1, 0), (0.3, 0.9, 0.316), (0.7, 0.6, 0.5)))
    L.or_sample_uber(p(n), p(wo), C.c_float(0.35), C.c_float(0.0), p(alb), C.c_uint32(sx), C.c_uint32(sy), p(out))
    smp.append({"state": [sx, sy], "out_bits": bits(out)})
g["sample_uber_rough0.35"] = smp
# tone map
acc = np.zeros((1, 5, 4), np.float32)
acc[0, :, :3] = np.array([0.0, 0.18, 0.5, 1.0, 4.0])[:, None]
g["tonemap"] = {}
for op in (0, 1):
    for ex in (1.0, 2.0):
        g["tonemap"][f"op{op}_exp{ex}"] = oracle.tonemap(acc, ex, op)[0, :, 0].tolist()
# sky
import math

el = math.radians(45.0)
sun = np.array([math.cos(el) * 0.6, math.sin(el), math.cos(el) * 0.8], np.float32)
sun /= np.linalg.norm(sun)
cf = oracle.sky_coeffs(sun)
g["sky_coeffs_sun45"] = {"sun": bits(sun), "coeffs_bits": bits(cf)}
sky = oracle.sky_bake(cf, sun, 8)
g["sky_texels_size8"] = {f"{f},{j},{i}": bits(sky[f, j, i, :3]) for f, j, i in ((0, 1, 1), (1, 6, 2), (2, 4, 4), (3, 0, 0), (4, 2, 5), (5, 7, 7))}
# scenes
s = scenes.cornell_box(64, 64)
o = oracle.OracleScene(s, brute_force=True)
ids = o.trace_primary_ids(s.push_constants(1))
g["cornell64_primary_frame1"] = {
    "sha256_inst_geom_prim": hashlib.sha256(b"".join(a.tobytes() for a in ids[:3])).hexdigest(),
    "sha256_tuv_bits": hashlib.sha256(b"".join(a.tobytes() for a in ids[3:])).hexdigest(),
    "geometry_histogram": {str(int(k)): int(v) for k, v in zip(*np.unique(ids[1], return_counts=True))},
}
acc = o.render(17)
g["cornell64_16spp"] = {"mean_rgb": [float(x) for x in acc[..., :3].mean((0, 1))], "extension_rays": int(o.counters[0]), "shadow_rays": int(o.counters[1])}
# white furnace: env = 1, albedo 1, only the env light -> documents the env double count
fs = scenes.cornell_box(32, 32)
fs.materials["albedo"][:] = [1, 1, 1, 1]
fs.materials["emissive"][:] = 0
from helios_b200 import abi

fs.lights = np.zeros(1, abi.LIGHT)
fs.lights[0]["light_data0"] = [abi.LIGHT_ENVIRONMENT_MAP, 0, 0, 0]
fs.env_cube = (2, np.ones((6, 2, 2, 4), np.float32))
fo = oracle.OracleScene(fs)
raw = np.zeros((32, 32, 4), np.float32)
tot = np.zeros(3)
for f in range(1, 9):
    a = np.zeros((32, 32, 4), np.float32)
    fo.render_frame(fs.push_constants(f), a, raw_L=raw)
    tot += raw[..., :3].mean((0, 1))
g["white_furnace_mean_raw_L_8spp"] = [float(x) for x in tot / 8]
out = ROOT / "tests" / "golden" / "oracle_golden.json"
out.write_text(json.dumps(g, indent=1))
print("wrote", out)
