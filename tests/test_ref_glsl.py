"""Pins for the oracle: the restatement (oracle/helios_oracle.cpp) against THE REFERENCE'S OWN SHADERS.

oracle/_ref/libhelios_glsl_ref.so is /root/reference/src/engine/shader/{random,sampling,common,brdf,path_trace_rgen,
path_trace_rchit,path_trace_rahit,path_trace_rmiss}.glsl, path_trace_shadow.{rchit,rmiss}, tone_map.frag,
procedural_sky.frag and the host functions of gfx/hosek_wilkie_sky_model.cpp, compiled as C++ by oracle/ref_glsl/
(lexical rewrite at build time; sources never enter the repository).  Two layers:

* side by side (needs the library: built here from /root/reference, or the prebuilt copy that travels to the GPU
  box) — same scene object, same driver (traversal, texture units), shader logic from the reference vs restated:
  every pixel of every frame bit-identical, ray counts identical;
* golden (always runs) — tests/golden/ref_glsl_golden.npz holds outputs of the reference shaders generated here
  by tools/make_ref_golden.py; the restatement must reproduce them bit for bit wherever it runs.
"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))
import make_ref_golden as G  # noqa: E402

GOLD = np.load(ROOT / "tests" / "golden" / "ref_glsl_golden.npz")


def p(a):
    return a.ctypes.data_as(C.c_void_p)


def same_bits(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))


@pytest.fixture(scope="module")
def ref(oracle_mod):
    L = oracle_mod.ref_lib()
    if L is None:
        pytest.skip("oracle/_ref/libhelios_glsl_ref.so not available (no /root/reference, no prebuilt copy)")
    return L


# ---- golden layer: restatement vs committed outputs of the reference shaders -----------------------------------
@pytest.mark.parametrize("name", list(G.GOLDEN_SCENES))
def test_restatement_reproduces_reference_frames(name, oracle_mod):
    s = G.GOLDEN_SCENES[name]()
    o = oracle_mod.OracleScene(s, sky_size=G.GOLDEN_SKY_SIZE)
    a = np.zeros((s.height, s.width, 4), np.float32)
    a[..., 3] = 1
    for f in range(G.GOLDEN_FRAMES):
        o.render_frame(s.push_constants(f), a)
    assert same_bits(a, GOLD[f"frame/{name}"]), f"max |diff| {np.abs(a - GOLD[f'frame/{name}']).max()}"
    assert np.array_equal(o.counters, GOLD[f"rays/{name}"])  # extension rays, shadow rays


def test_restatement_reproduces_reference_rng(oracle_mod):
    L = oracle_mod.lib()
    for sx, sy in G.RNG_SEEDS:
        res, st, fl = np.zeros(16, np.uint32), np.zeros(32, np.uint32), np.zeros(16, np.float32)
        L.or_rng_sequence(C.c_uint32(sx), C.c_uint32(sy), C.c_uint32(16), p(res), p(st))
        L.or_next_floats(C.c_uint32(sx), C.c_uint32(sy), C.c_uint32(16), p(fl))
        assert np.array_equal(res, GOLD[f"rng/{sx},{sy}/results"]) and np.array_equal(st, GOLD[f"rng/{sx},{sy}/states"])
        assert same_bits(fl, GOLD[f"rng/{sx},{sy}/floats"])
    init = np.zeros((4, 2), np.uint32)
    for k, (x, y, f) in enumerate(((0, 0, 0), (1919, 1079, 63), (5, 7, 1), (3839, 2159, 255))):
        L.or_rng_init(C.c_uint32(x), C.c_uint32(y), C.c_uint32(f), p(init[k]))
    assert np.array_equal(init, GOLD["rng/init"])


def test_restatement_reproduces_reference_brdf(oracle_mod):
    L = oracle_mod.lib()
    ev, sm = np.zeros((64, 4), np.float32), np.zeros((64, 7), np.float32)
    for k, (n, wo, wi, ro, me, alb, sx, sy) in enumerate(G.brdf_cases()):
        L.or_evaluate_uber(p(n), p(wo), p(wi), C.c_float(ro), C.c_float(me), p(alb), p(ev[k]))
        L.or_sample_uber(p(n), p(wo), C.c_float(ro), C.c_float(me), p(alb), C.c_uint32(sx), C.c_uint32(sy), p(sm[k]))
    assert same_bits(ev, GOLD["brdf/evaluate"]) and same_bits(sm, GOLD["brdf/sample"])


def test_restatement_reproduces_reference_sky_and_tonemap(oracle_mod):
    for k, (d, tb, al, ny) in enumerate(G.SKY_CASES):
        d = np.asarray(d, np.float32)
        cf = oracle_mod.sky_coeffs(d, tb, al, ny)
        assert same_bits(cf, GOLD[f"sky/coeffs/{k}"])  # also pins helios_b200/data/hosek_rgb_v1_4a.f64
        if k < 2:
            assert same_bits(oracle_mod.sky_bake(cf, d, 16), GOLD[f"sky/bake/{k}"])
    img = G.tonemap_image()
    for op in (0, 1, 2):
        assert np.array_equal(oracle_mod.tonemap(img, 0.8, op), GOLD[f"tonemap/{op}"])


def test_product_sky_fit_matches_reference_host_code():
    """helios_b200/sky.py (the Python host's coefficient fit) against the reference's host functions (golden)"""
    from helios_b200 import sky

    for k, (d, tb, al, ny) in enumerate(G.SKY_CASES):
        cf = sky.sky_coefficients(np.asarray(d, np.float32), turbidity=tb, albedo=al, normalized_sun_y=ny)
        np.testing.assert_allclose(np.asarray(cf, np.float32).ravel(), GOLD[f"sky/coeffs/{k}"], rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("name", list(G.GOLDEN_SCENES))
def test_restatement_reproduces_reference_ray_debug_view(name, oracle_mod):
    """gather_debug_rays: the restatement of the RAY_DEBUG_VIEW blocks against the committed output of the reference's shaders
    compiled with that define (same vertices, same order: both run the launch sequentially, depth first)"""
    s = G.GOLDEN_SCENES[name]()
    o = oracle_mod.OracleScene(s, sky_size=G.GOLDEN_SKY_SIZE)
    v, n = o.gather_debug_rays(G.raydebug_push_constants(s, o), G.RAYDEBUG_RAYS, max_vertices=4096)
    gold = GOLD[f"raydebug/{name}"]
    assert n == len(gold) and same_bits(v, gold)
    assert name in ("soup",) or n > 0


# ---- side-by-side layer: needs the reference-GLSL library ---------------------------------------------------------
def test_golden_file_is_current(ref, oracle_mod):
    """the committed fixture is what the reference shaders produce today (guards against a stale .npz)"""
    s = G.GOLDEN_SCENES["foliage"]()
    r = oracle_mod.GlslRefScene(s, sky_size=G.GOLDEN_SKY_SIZE)
    a = np.zeros((s.height, s.width, 4), np.float32)
    a[..., 3] = 1
    for f in range(G.GOLDEN_FRAMES):
        r.render_frame(s.push_constants(f), a)
    assert same_bits(a, GOLD["frame/foliage"])


@pytest.mark.parametrize("name", list(G.GOLDEN_SCENES))
def test_side_by_side_frames_bit_identical(name, ref, oracle_mod):
    s = G.GOLDEN_SCENES[name]()
    r = oracle_mod.GlslRefScene(s, sky_size=G.GOLDEN_SKY_SIZE, brute_force=(name == "cornell"))
    a = np.zeros((s.height, s.width, 4), np.float32)
    a[..., 3] = 1
    b = a.copy()
    for f in (0, 1, 2, 17):
        pc = s.push_constants(f)
        r.render_frame(pc, a)
        r.restated_frame(pc, b)
        assert same_bits(a, b), f"{name} frame {f}: max |diff| {np.abs(a - b).max()}"
    assert np.array_equal(r.counters, r.restated_counters)


@pytest.mark.parametrize("name", ["cornell", "terrain_textured", "foliage", "city"])
def test_side_by_side_ray_debug_view(name, oracle_mod):
    """the reference's shaders built with -DRAY_DEBUG_VIEW (oracle/_ref/libhelios_glsl_ref_raydebug.so) vs the restatement in
    the same library: every vertex bit-identical, for several pixels, frames and bounce limits"""
    if oracle_mod.ref_debug_lib() is None:
        pytest.skip("oracle/_ref/libhelios_glsl_ref_raydebug.so not available")
    s = G.GOLDEN_SCENES[name]()
    r = oracle_mod.GlslRefDebugScene(s, sky_size=G.GOLDEN_SKY_SIZE)
    total = 0
    for frame, bounces, px in ((0, 1, (3, 5)), (2, 4, (s.width // 2, s.height // 2)), (9, 8, (s.width // 3, s.height - 4)), (4, 6, (s.width - 2, 2 * s.height // 3))):
        pc = s.push_constants(frame, pixel_coord=px, max_ray_bounces=bounces)
        a, na = r.gather_debug_rays(pc, 48, max_vertices=8192)
        b, nb = r.restated_debug_rays(pc, 48, max_vertices=8192)
        assert na == nb and same_bits(a, b), f"{name} frame {frame}: {na} vs {nb}"
        total += na
    assert total > 0


def test_side_by_side_tiles_and_depths(ref, oracle_mod):
    """launch rectangles (path_integrator.cpp:312-336) and max_ray_bounces 1..5 through both paths"""
    from helios_b200 import scenes

    s = scenes.cornell_box(40, 28)
    r = oracle_mod.GlslRefScene(s)
    for depth in (1, 2, 5):
        a = np.zeros((s.height, s.width, 4), np.float32)
        b = a.copy()
        for tile in ((0, 0), (16, 0), (32, 0), (0, 16), (16, 16), (32, 16)):
            pc = s.push_constants(2, tile=tile, max_ray_bounces=depth)
            r.render_frame(pc, a, launch=(16, 16))
            r.restated_frame(pc, b, launch=(16, 16))
        assert same_bits(a, b) and a[..., :3].max() > 0


def test_side_by_side_units_random(ref, oracle_mod):
    rng = np.random.default_rng(7)
    for _ in range(500):
        n, wo, wi = (rng.normal(size=3).astype(np.float32) for _ in range(3))
        n, wo, wi = n / np.linalg.norm(n), wo / np.linalg.norm(wo), wi / np.linalg.norm(wi)
        alb, ro, me = rng.random(3).astype(np.float32), float(rng.random()), float(rng.random())
        sx, sy = int(rng.integers(1, 2**32)), int(rng.integers(1, 2**32))
        a, b = np.zeros(4, np.float32), np.zeros(4, np.float32)
        ref.or_evaluate_uber(p(n), p(wo), p(wi), C.c_float(ro), C.c_float(me), p(alb), p(a))
        ref.ref_evaluate_uber(p(n), p(wo), p(wi), C.c_float(ro), C.c_float(me), p(alb), p(b))
        assert same_bits(a, b)
        a, b = np.zeros(7, np.float32), np.zeros(7, np.float32)
        ref.or_sample_uber(p(n), p(wo), C.c_float(ro), C.c_float(me), p(alb), C.c_uint32(sx), C.c_uint32(sy), p(a))
        ref.ref_sample_uber(p(n), p(wo), C.c_float(ro), C.c_float(me), p(alb), C.c_uint32(sx), C.c_uint32(sy), p(b))
        assert same_bits(a, b)
        a, b = np.zeros(2, np.uint32), np.zeros(2, np.uint32)
        x, y, f = int(rng.integers(0, 4096)), int(rng.integers(0, 4096)), int(rng.integers(0, 100000))
        ref.or_rng_init(C.c_uint32(x), C.c_uint32(y), C.c_uint32(f), p(a))
        ref.ref_rng_init(C.c_uint32(x), C.c_uint32(y), C.c_uint32(f), p(b))
        assert np.array_equal(a, b)
        assert ref.or_rng_hash(C.c_uint32(sx)) == ref.ref_rng_hash(C.c_uint32(sx))


def test_side_by_side_sky_host_random(ref, oracle_mod):
    rng = np.random.default_rng(3)
    for i in range(200):
        d = rng.normal(size=3).astype(np.float32)
        d /= np.linalg.norm(d)
        tb, al, ny = float(rng.uniform(1, 10)), float(rng.uniform(0, 1)), (1.15, 0.0, 0.8)[i % 3]
        b = np.zeros(40, np.float32)
        ref.ref_sky_coeffs(p(d), C.c_float(tb), C.c_float(al), C.c_float(ny), p(b))
        assert same_bits(oracle_mod.sky_coeffs(d, tb, al, ny), b)


def test_gen_rewrite_rules():
    """the lexical rewrite of oracle/ref_glsl/gen.py on synthetic GLSL (no reference text involved)"""
    sys.path.insert(0, str(ROOT / "oracle" / "ref_glsl"))
    import gen

    out = gen.rewrite(
        '#version 460\n#extension GL_EXT_ray_tracing : require\n#include "a.glsl"\n'
        "layout (set = 1, binding = 0, std430) readonly buffer VB \n{\n    Vertex data[];\n} Vs[];\n"
        "layout(push_constant) uniform PC\n{\n    mat4 m;\n    float f;\n} u_pc;\n"
        "layout (set = 4, binding = 0) uniform sampler2D s_T[];\n"
        "layout(location = 1) rayPayloadEXT P p_X;\nhitAttributeEXT vec2 attr;\n"
        "float f(in vec3 a, out float b, inout RNG r) { return 2.0 * a.x + 1.0f + 1e-3 + 3; }\n"
        "vec2 g(inout RNG r) { return vec2(h(r), max(vec2(0.5), h(r)).x); }\n")
    assert "#version" not in out and "#extension" not in out and '#include "a.glsl.inc"' in out
    assert "struct VB" in out and "Vertex* data;" in out and "static VB* Vs;" in out
    assert "struct PC" in out and "static PC u_pc;" in out and "static sampler2D_array s_T;" in out
    assert "static thread_local P p_X;" in out and "static thread_local vec2 attr;" in out
    assert "float f(vec3 a, float& b, RNG& r) { return 2.0f * a.x + 1.0f + 1e-3f + 3; }" in out
    assert "return vec2{h(r), max(vec2{0.5f}, h(r)).x};" in out
