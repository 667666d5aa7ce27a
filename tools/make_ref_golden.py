#!/usr/bin/env python3
"""Generates tests/golden/ref_glsl_golden.npz by running THE REFERENCE'S OWN SHADERS (oracle/_ref/
libhelios_glsl_ref.so = /root/reference/src/engine/shader/*.glsl compiled as C++ by oracle/ref_glsl/, plus the host
functions of gfx/hosek_wilkie_sky_model.cpp) in this container.  The fixture travels to machines that have neither
/root/reference nor the prebuilt library, where tests/test_ref_glsl.py checks the restatement (oracle/) and the
CUDA path against it.  Only ref_* entry points produce the values stored here; the inputs are the seeded scene
generators of helios_b200/scenes.py (GOLDEN_SCENES below is imported by the tests so both sides build the same
inputs).

    python tools/make_ref_golden.py
"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from helios_b200 import scenes  # noqa: E402

GOLDEN_SCENES = {
    "cornell": lambda: scenes.cornell_box(48, 48),
    "cornell_lens_bias": lambda: scenes.cornell_box(40, 40, aperture_radius=0.1, shadow_ray_bias=1e-3),
    "soup": lambda: scenes.triangle_soup(3000, 64, 36),
    "terrain": lambda: scenes.terrain_scene(grid=40, n_spheres=6, sphere_level=1, width=64, height=36),
    "terrain_textured": lambda: scenes.terrain_scene(grid=24, n_spheres=4, sphere_level=1, width=64, height=36, textured=True),
    "foliage": lambda: scenes.foliage_scene(n_clusters=60, cards_per_cluster=12, width=64, height=36, ground_grid=8, tex_size=32),
    "city": lambda: scenes.city_scene(n_instances=30, n_meshes=3, width=64, height=36, floors=(2, 4), detail=(1, 3)),
}
GOLDEN_FRAMES = 3
GOLDEN_SKY_SIZE = 32
RNG_SEEDS = [(1, 2), (123456789, 987654321), (0xDEADBEEF, 0x12345678), (42, 4242)]
SKY_CASES = [((0.0, 0.7071068, 0.7071068), 4.0, 0.1, 1.15), ((0.3, 0.2, -0.9327379), 2.5, 0.3, 1.15), ((0.0, 1.0, 0.0), 7.75, 0.0, 0.0), ((0.6, -0.1, 0.7937254), 4.0, 0.1, 1.15)]


RAYDEBUG_RAYS, RAYDEBUG_BOUNCES, RAYDEBUG_FRAME = 24, 5, 3


def raydebug_push_constants(s, oracle_scene):
    """the push constants of a gather_debug_rays launch through a pixel of `s` that sees geometry"""
    hit = oracle_scene.trace_primary_ids(s.push_constants(RAYDEBUG_FRAME))[0].reshape(s.height, s.width) != 0xFFFFFFFF
    ys, xs = np.nonzero(hit)
    k = len(ys) // 2
    return s.push_constants(RAYDEBUG_FRAME, pixel_coord=(int(xs[k]), s.height - int(ys[k])), max_ray_bounces=RAYDEBUG_BOUNCES)


def brdf_cases():
    rng = np.random.default_rng(11)
    out = []
    for _ in range(64):
        n, wo, wi = (rng.normal(size=3).astype(np.float32) for _ in range(3))
        n, wo, wi = n / np.linalg.norm(n), wo / np.linalg.norm(wo), wi / np.linalg.norm(wi)
        if n @ wo < 0:
            wo = -wo
        out.append((n.astype(np.float32), wo.astype(np.float32), wi.astype(np.float32), np.float32(rng.random()), np.float32(rng.random()), rng.random(3).astype(np.float32), int(rng.integers(1, 2**32)), int(rng.integers(1, 2**32))))
    return out


def tonemap_image():
    rng = np.random.default_rng(5)
    img = (rng.random((9, 13, 4)).astype(np.float32) ** 3 * 4).astype(np.float32)
    img[..., 3] = 1
    return img


def p(a):
    return a.ctypes.data_as(C.c_void_p)


def main():
    from oracle import oracle

    L = oracle.ref_lib()
    if L is None:
        raise SystemExit("needs oracle/_ref/libhelios_glsl_ref.so (i.e. /root/reference)")
    g = {}
    for name, mk in GOLDEN_SCENES.items():
        s = mk()
        r = oracle.GlslRefScene(s, sky_size=GOLDEN_SKY_SIZE)
        a = np.zeros((s.height, s.width, 4), np.float32)
        a[..., 3] = 1
        for f in range(GOLDEN_FRAMES):
            r.render_frame(s.push_constants(f), a)
        g[f"frame/{name}"] = a
        g[f"rays/{name}"] = r.counters.copy()
    for sx, sy in RNG_SEEDS:
        res, st, fl = np.zeros(16, np.uint32), np.zeros(32, np.uint32), np.zeros(16, np.float32)
        L.ref_rng_sequence(C.c_uint32(sx), C.c_uint32(sy), C.c_uint32(16), p(res), p(st))
        L.ref_next_floats(C.c_uint32(sx), C.c_uint32(sy), C.c_uint32(16), p(fl))
        g[f"rng/{sx},{sy}/results"], g[f"rng/{sx},{sy}/states"], g[f"rng/{sx},{sy}/floats"] = res, st, fl
    init = np.zeros((4, 2), np.uint32)
    for k, (x, y, f) in enumerate(((0, 0, 0), (1919, 1079, 63), (5, 7, 1), (3839, 2159, 255))):
        L.ref_rng_init(C.c_uint32(x), C.c_uint32(y), C.c_uint32(f), p(init[k]))
    g["rng/init"] = init
    ev, sm = np.zeros((64, 4), np.float32), np.zeros((64, 7), np.float32)
    for k, (n, wo, wi, ro, me, alb, sx, sy) in enumerate(brdf_cases()):
        L.ref_evaluate_uber(p(n), p(wo), p(wi), C.c_float(ro), C.c_float(me), p(alb), p(ev[k]))
        L.ref_sample_uber(p(n), p(wo), C.c_float(ro), C.c_float(me), p(alb), C.c_uint32(sx), C.c_uint32(sy), p(sm[k]))
    g["brdf/evaluate"], g["brdf/sample"] = ev, sm
    for k, (d, tb, al, ny) in enumerate(SKY_CASES):
        d = np.asarray(d, np.float32)
        cf = np.zeros(40, np.float32)
        L.ref_sky_coeffs(p(d), C.c_float(tb), C.c_float(al), C.c_float(ny), p(cf))
        g[f"sky/coeffs/{k}"] = cf
        if k < 2:
            g[f"sky/bake/{k}"] = oracle.ref_sky_bake(cf, d, 16)
    img = tonemap_image()
    for op in (0, 1, 2):
        g[f"tonemap/{op}"] = oracle.ref_tonemap(img, 0.8, op)
    # the RAY_DEBUG_VIEW build of the same shaders (second library): segment vertices of gather_debug_rays
    if oracle.ref_debug_lib() is not None:
        for name, mk in GOLDEN_SCENES.items():
            s = mk()
            r = oracle.GlslRefDebugScene(s, sky_size=GOLDEN_SKY_SIZE)
            v, n = r.gather_debug_rays(raydebug_push_constants(s, r), RAYDEBUG_RAYS, max_vertices=4096)
            assert n == len(v)
            g[f"raydebug/{name}"] = v
    out = ROOT / "tests" / "golden" / "ref_glsl_golden.npz"
    np.savez_compressed(out, **g)
    print(f"wrote {out} ({out.stat().st_size} bytes, {len(g)} arrays)")


if __name__ == "__main__":
    main()
