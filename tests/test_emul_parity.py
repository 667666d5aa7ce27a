"""Device-logic parity on the CPU: the HL_HD headers the CUDA kernels are built from (helios_b200/csrc/hl_*.h),
compiled with g++ by tests/emul and driven in wavefront order, against the oracle.  This is what can be checked
without a GPU: builder (Morton / radix tree / collapse / quantisation), traversal and tie rule, any-hit, shading,
RNG order, accumulation.  The same comparisons run against the real kernels in test_gpu_parity.py."""
import numpy as np
import pytest

from helios_b200 import abi, scenes
from helios_b200.sky import sky_coefficients


@pytest.fixture(scope="module")
def emul():
    from tests.emul import emul as e

    e.build()
    return e


def pair(scene, oracle_mod, emul, sky_size=32, **kw):
    cf = sky_coefficients(scene.sun_direction) if scene.sun_direction is not None else None
    o = oracle_mod.OracleScene(scene, sky_coeffs_override=cf, sky_size=sky_size)
    faces = oracle_mod.sky_bake(cf, scene.sun_direction, sky_size) if cf is not None else None
    return o, emul.EmulScene(scene, sky_faces=faces, **kw)


def assert_ids_equal(a, b):
    for x, y in zip(a, b):
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32))


SCENES = {
    "cornell": lambda: scenes.cornell_box(48, 48),
    "cornell_lens_bias": lambda: scenes.cornell_box(40, 40, aperture_radius=0.1, shadow_ray_bias=1e-3),
    "soup": lambda: scenes.triangle_soup(3000, 96, 54),
    "terrain": lambda: scenes.terrain_scene(grid=40, n_spheres=6, sphere_level=1, width=96, height=54),
    "terrain_textured": lambda: scenes.terrain_scene(grid=24, n_spheres=4, sphere_level=1, width=64, height=36, textured=True),
    "foliage": lambda: scenes.foliage_scene(n_clusters=60, cards_per_cluster=12, width=96, height=54, ground_grid=8, tex_size=32),
    "city": lambda: scenes.city_scene(n_instances=30, n_meshes=3, width=96, height=54, floors=(2, 4), detail=(1, 3)),
}


@pytest.mark.parametrize("name", list(SCENES))
def test_primary_ids_bit_exact(name, oracle_mod, emul):
    s = SCENES[name]()
    o, e = pair(s, oracle_mod, emul)
    for frame in (0, 5):
        pc = s.push_constants(frame)
        assert_ids_equal(o.trace_primary_ids(pc), e.trace_primary_ids(pc))


@pytest.mark.parametrize("name", list(SCENES))
def test_radiance_matches(name, oracle_mod, emul):
    s = SCENES[name]()
    o, e = pair(s, oracle_mod, emul)
    a, b = o.render(4), e.render(4)
    # identical discrete decisions; only the summation order of the per-bounce terms differs
    assert np.abs(a - b)[..., :3].max() < 2e-6
    assert int(o.counters[0]) == int(e.counters[0])  # extension rays identical
    assert int(e.counters[1]) <= int(o.counters[1])  # black shadow terms are skipped


def test_two_level_equals_single_level(oracle_mod, emul):
    s = scenes.triangle_soup(2000, 64, 36)
    o, e1 = pair(s, oracle_mod, emul)
    _, e2 = pair(s, oracle_mod, emul, force_two_level=True)
    pc = s.push_constants(2)
    assert_ids_equal(e1.trace_primary_ids(pc), e2.trace_primary_ids(pc))
    assert_ids_equal(o.trace_primary_ids(pc), e2.trace_primary_ids(pc))


@pytest.mark.parametrize("flags", [0, 1, 3])
def test_random_rays_and_flags(flags, oracle_mod, emul):
    s = scenes.foliage_scene(n_clusters=40, cards_per_cluster=10, width=32, height=18, ground_grid=6, tex_size=16)
    o, e = pair(s, oracle_mod, emul)
    rng = np.random.default_rng(flags)
    n = 4000
    org = (rng.random((n, 3)) - 0.5) * np.array([50, 4, 50]) + np.array([0, 4, 0])
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[::50, 0] = 0.0  # axis-parallel components
    d[::77, 1] = 1e-12
    rays = np.concatenate([org, np.full((n, 1), 1e-3), d, np.full((n, 1), 30.0)], 1).astype(np.float32)
    a, b = o.trace_rays(rays, flags), e.trace_rays(rays, flags)
    if flags & 2:  # terminate on first hit: only hit / miss is defined
        assert np.array_equal(np.isinf(a[:, 0]), np.isinf(b[:, 0]))
    else:
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert np.isfinite(a[:, 0]).mean() > 0.05


def test_triangle_postponing_does_not_change_hits(oracle_mod, emul):
    """the GPU step scheduler may put a leaf group on the stack and keep descending (hl_bvh.h trav_postpone);
    forced on every opportunity here, single- and two-level, the hits stay bit-identical to the oracle"""
    s = scenes.triangle_soup(20000, 96, 54)
    o, e1 = pair(s, oracle_mod, emul)
    _, e2 = pair(s, oracle_mod, emul, force_two_level=True)
    pc = s.push_constants(3)
    ref = o.trace_primary_ids(pc)
    emul.lib().em_set_force_postpone(1)
    try:
        assert_ids_equal(ref, e1.trace_primary_ids(pc))
        assert_ids_equal(ref, e2.trace_primary_ids(pc))
    finally:
        emul.lib().em_set_force_postpone(0)


@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 5, 9, 33])
def test_tiny_meshes(n, oracle_mod, emul):
    s = scenes.triangle_soup(max(n, 1), 48, 27, seed=100 + n)
    if n == 0:  # empty geometry: a submesh with zero triangles
        s.meshes[0].submeshes[0]["index_count"] = 0
    o, e = pair(s, oracle_mod, emul)
    pc = s.push_constants(1)
    assert_ids_equal(o.trace_primary_ids(pc), e.trace_primary_ids(pc))


def test_tiled_launch_equals_full_frame(oracle_mod, emul):
    """PathIntegrator tiles (path_integrator.cpp:312-336): tile offsets in launch_id_size.xy, 128^2 launches"""
    s = scenes.cornell_box(80, 56)
    _, e = pair(s, oracle_mod, emul)
    full = np.zeros((s.height, s.width, 4), np.float32)
    e.render_frame(s.push_constants(1), full)
    tiled = np.zeros_like(full)
    for ty in range(0, s.height, 32):
        for tx in range(0, s.width, 32):
            e.render_frame(s.push_constants(1, tile=(tx, ty)), tiled, launch=(32, 32))
    assert np.array_equal(full, tiled)


def test_sum_mode_equals_mean(oracle_mod, emul):
    s = scenes.cornell_box(32, 32)
    _, e = pair(s, oracle_mod, emul)
    mean = e.render(5)
    acc = np.zeros_like(mean)
    for f in range(1, 5):
        e.render_frame(s.push_constants(f), acc, accum_mode=abi.ACCUM_SUM)
    assert np.allclose(acc[..., :3] / 4.0, mean[..., :3], atol=1e-6)


def test_tonemap_and_sky_match_oracle(oracle_mod, emul):
    rng = np.random.default_rng(1)
    acc = (rng.random((20, 30, 4)) * 2).astype(np.float32)
    import ctypes as C

    for op in (0, 1):
        out = np.zeros((20, 30, 4), np.uint8)
        emul.lib().em_tonemap(acc.ctypes.data_as(C.c_void_p), C.c_uint32(30), C.c_uint32(20), C.c_float(1.5), C.c_int(op), C.c_float(1.0), out.ctypes.data_as(C.c_void_p))
        assert np.array_equal(out, oracle_mod.tonemap(acc, 1.5, op))
    sun = np.array([0.3, 0.8, 0.52], np.float32)
    sun /= np.linalg.norm(sun)
    cf = sky_coefficients(sun)
    assert np.allclose(emul.sky_bake(cf, sun, 16), oracle_mod.sky_bake(cf, sun, 16), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("name", ["cornell", "terrain_textured", "foliage", "city"])
def test_debug_output_buffers(name, oracle_mod, emul):
    """albedo / normals / roughness / metallic / emissive views (debug_visualization.frag:144-161)"""
    s = SCENES[name]()
    o, e = pair(s, oracle_mod, emul)
    pc = s.push_constants(1)
    hit = o.trace_primary_ids(pc)[0].reshape(s.height, s.width) != abi.MISS_ID
    for which in range(5):
        a, b = o.output_buffer(pc, which), e.output_buffer(pc, which)
        assert np.array_equal(a[~hit], np.broadcast_to(np.float32([0, 0, 0, 1]), a[~hit].shape))  # the pass's clear colour
        assert np.all(a[..., 3] == 1.0)
        # same hits (bit-exact above) and the same fetch; contraction differences only in the interpolated normal
        assert np.abs(a - b).max() < 2e-6, (which, np.abs(a - b).max())
    rough = o.output_buffer(pc, abi.OUTPUT_BUFFER_ROUGHNESS)
    mats = s.materials
    if all(int(m["texture_indices0"][2]) == -1 for m in mats):  # constant roughness: the view shows the table's value, un-floored
        vals = {np.float32(m["roughness_metallic"][0]) for m in mats}
        assert set(np.unique(rough[hit][:, 0])) <= vals


@pytest.mark.parametrize("name", ["cornell", "terrain", "foliage", "city"])
def test_ray_debug_view(name, oracle_mod, emul):
    """PathIntegrator::gather_debug_rays — the RAY_DEBUG_VIEW pipeline (no Russian roulette, one line segment
    per secondary ray through one pixel)"""
    s = SCENES[name]()
    o, e = pair(s, oracle_mod, emul)
    # a pixel that sees geometry (the push constant holds (x, H - y): path_integrator.cpp:146)
    hit = o.trace_primary_ids(s.push_constants(3))[0].reshape(s.height, s.width) != abi.MISS_ID
    ys, xs = np.nonzero(hit)
    k = len(ys) // 2
    pc = s.push_constants(3, pixel_coord=(int(xs[k]), s.height - int(ys[k])), max_ray_bounces=5)
    a, na = o.gather_debug_rays(pc, 40)
    b, nb = e.gather_debug_rays(pc, 40)
    assert na == nb and na % 2 == 0 and na > 0
    assert np.allclose(a, b, rtol=1e-5, atol=1e-5)  # same hits, same random numbers; fp contraction differences only
    seg = a.reshape(-1, 2, 8)
    assert np.array_equal(seg[:, 0, 4:], seg[:, 1, 4:]) and np.all(seg[..., 3] == 1.0) and np.all(seg[..., 7] == 1.0)
    assert np.all((seg[..., 4:7] >= 0.5) & (seg[..., 4:7] < 1.0))  # rgen:193-195: next_float * 0.5 + 0.5
    # no Russian roulette: a path that keeps hitting geometry leaves max_ray_bounces - 1 segments, each starting where the
    # previous one ended (the indirect ray starts at the hit position, rchit:521)
    colours, counts = np.unique(seg[:, 0, 4:7], axis=0, return_counts=True)
    assert counts.max() <= 4 and len(colours) <= 40
    # capacity: the count keeps growing past the buffer like the reference's draw argument, writes stop
    c, nc = e.gather_debug_rays(pc, 40, max_vertices=6)
    assert nc == nb and len(c) == 6 and np.array_equal(c, b[:6])


@pytest.mark.parametrize("bounces", [0, 1, 2])
def test_bounce_limits(bounces, oracle_mod, emul):
    """max_ray_bounces = 0 still traces and shades the primary ray (rgen:205, rchit:575 only guards the indirect ray)"""
    s = SCENES["cornell"]()
    o, e = pair(s, oracle_mod, emul)
    a, b = o.render(3, max_ray_bounces=bounces), e.render(3, max_ray_bounces=bounces)
    assert np.abs(a - b)[..., :3].max() < 2e-6 and a[..., :3].max() > 0


def test_grazing_rays_never_lose_an_edge_hit(oracle_mod, emul):
    """tests/graze.py: 300 000 adversarial grazing rays; the device traversal (both tree builds, one- and two-level) and the
    oracle's BVH must return the brute-force closest hit for every one of them, bit for bit"""
    import ctypes as C

    from tests import graze

    s, tris = graze.graze_scene()
    rays = graze.graze_rays(tris, 300_000)
    brute = oracle_mod.OracleScene(s, brute_force=True).trace_rays(rays, 0)
    assert np.isfinite(brute[:, 0]).mean() > 0.9
    assert np.array_equal(oracle_mod.OracleScene(s).trace_rays(rays, 0).view(np.uint32), brute.view(np.uint32))
    try:
        for top, two_level in ((0, False), (2, False), (2, True)):
            emul.lib().em_set_sah_top(C.c_int(top))
            h = emul.EmulScene(s, force_two_level=two_level).trace_rays(rays, 0)
            bad = (h.view(np.uint32) != brute.view(np.uint32)).any(1)
            assert not bad.any(), f"tree {top}, two-level {two_level}: {int(bad.sum())} of {len(rays)} grazing rays differ from brute force"
    finally:
        emul.lib().em_set_sah_top(C.c_int(2))  # HL_DEFAULT_SAH_CLUSTER


@pytest.mark.parametrize("tex_size", [1, 3, 5, 7])
def test_odd_texture_extents(tex_size, oracle_mod, emul):
    """REPEAT addressing without integer modulo (hl_tex.h texture_footprint) on 1x1 and non-power-of-two textures, through
    the any-hit alpha fetch and the albedo fetch"""
    s = scenes.foliage_scene(n_clusters=30, cards_per_cluster=8, width=48, height=27, ground_grid=4, tex_size=tex_size)
    o, e = pair(s, oracle_mod, emul)
    pc = s.push_constants(1)
    assert_ids_equal(o.trace_primary_ids(pc), e.trace_primary_ids(pc))
    assert np.abs(o.render(3) - e.render(3)).max() < 2e-6


def hostile_rays(n=2000, seed=0):
    """zero / NaN / denormal / huge directions, infinite and NaN origins and intervals: a traversal must terminate on them"""
    rng = np.random.default_rng(seed)
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = rng.normal(size=(n, 3)) * 30
    rays[:, 3] = 1e-3
    rays[:, 4:7] = rng.normal(size=(n, 3))
    rays[:, 7] = 1e4
    rays[0::10, 4:7] = 0.0
    rays[1::10, 4] = np.nan
    rays[2::10, 0] = np.inf
    rays[3::10, 4:7] *= 1e-30
    with np.errstate(over="ignore"):
        rays[4::10, 4:7] *= np.float32(1e30)
        rays[7::10, 0:3] *= np.float32(1e20)
    rays[5::10, 7] = np.inf
    rays[6::10, 3] = np.nan
    return rays


@pytest.mark.timeout(120)
def test_hostile_rays_terminate_and_match(oracle_mod, emul):
    s = SCENES["city"]()
    o, e = pair(s, oracle_mod, emul)
    rays = hostile_rays()
    a, b = o.trace_rays(rays, 0), e.trace_rays(rays, 0)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def extreme_vertex_scenes():
    """finite but absurd and non-finite vertex positions: the builder must stay structurally valid (areas overflow to +inf
    from about 1e19 on) and every triangle that can be hit must still be found"""

    def soup(mod):
        s = scenes.triangle_soup(500, 64, 36, seed=1)
        with np.errstate(all="ignore"):
            mod(s.meshes[0].vertices["position"])
        return s

    def one(p):
        p[90, :3] = 1e30

    def tri(p):
        p[90:93, :3] = 1e30

    def neg(p):
        p[120:123, :3] = -3e38

    def mixed(p):
        p[3, 0], p[30, 1], p[60, 2] = np.nan, np.inf, -np.inf
        p[90:93, :3], p[120:123, :3] = 1e30, -3e38

    return {"one_vertex_1e30": soup(one), "triangle_1e30": soup(tri), "triangle_-3e38": soup(neg), "nan_inf_huge_mixed": soup(mixed)}


@pytest.mark.timeout(120)
@pytest.mark.parametrize("name", ["one_vertex_1e30", "triangle_1e30", "triangle_-3e38", "nan_inf_huge_mixed"])
def test_extreme_vertices(name, oracle_mod, emul):
    """(a triangle with one vertex at 1e30 used to make the collapse DP choose a leaf for a subtree of more than three
    triangles — inf <= inf — and the leaf writer ran past its array)"""
    s = extreme_vertex_scenes()[name]
    o = oracle_mod.OracleScene(s, brute_force=True)
    e = emul.EmulScene(s)
    pc = s.push_constants(1)
    with np.errstate(all="ignore"):
        a, b = o.trace_primary_ids(pc), e.trace_primary_ids(pc)
    assert_ids_equal(a, b)
    assert int((a[0] != abi.MISS_ID).sum()) > 100


def _set_instance_matrix(s, k, M):
    s.instances["model_matrix"][k] = M.T.reshape(16).astype(np.float32)  # column-major
    N4 = np.eye(4)
    N4[:3, :3] = np.linalg.inv(M[:3, :3]).T
    s.instances["normal_matrix"][k] = N4.T.reshape(16).astype(np.float32)


def instance_transform_scenes():
    """instanced city variants for the two-level traversal: non-uniform scale / shear / mirror / arbitrary linear maps,
    every instance at the same place (tie rule across instances), and the whole scene far from the origin"""
    rng = np.random.default_rng(4)
    base = lambda: scenes.city_scene(n_instances=12, n_meshes=3, width=96, height=54, floors=(1, 3), detail=(1, 2))  # noqa: E731
    out = {}
    s = base()
    for k in range(1, len(s.instances)):
        M = np.array(s.instances["model_matrix"][k], np.float64).reshape(4, 4).T
        A = np.eye(4)
        if k % 4 == 0:
            A[:3, :3] = np.diag([1.7, 0.3, 0.9])
        elif k % 4 == 1:
            A[0, 1], A[2, 0] = 0.6, -0.4
        elif k % 4 == 2:
            A[:3, :3] = np.diag([-1.0, 1.0, 1.0])
        else:
            A[:3, :3] = rng.normal(size=(3, 3)) * 0.8
        _set_instance_matrix(s, k, M @ A)
    out["non_rigid"] = s
    s = base()
    for k in range(2, len(s.instances)):
        for f in ("model_matrix", "normal_matrix", "mesh_index"):
            s.instances[f][k] = s.instances[f][1]
    s.submesh_info = [s.submesh_info[0]] + [s.submesh_info[1]] * (len(s.instances) - 1)
    out["coincident"] = s
    s = base()
    off = np.array([3e5, -2e5, 1e5])
    for k in range(len(s.instances)):
        M = np.array(s.instances["model_matrix"][k], np.float64).reshape(4, 4).T
        M[:3, 3] += off
        _set_instance_matrix(s, k, M)
    s.camera.position = (np.asarray(s.camera.position, np.float64) + off).astype(np.float32)
    out["far_from_origin"] = s
    return out


@pytest.mark.parametrize("name", ["non_rigid", "coincident", "far_from_origin"])
def test_instance_transform_edge_cases(name, oracle_mod, emul):
    s = instance_transform_scenes()[name]
    cf = sky_coefficients(s.sun_direction)
    o = oracle_mod.OracleScene(s, brute_force=True, sky_coeffs_override=cf, sky_size=32)
    e = emul.EmulScene(s, sky_faces=oracle_mod.sky_bake(cf, s.sun_direction, 32))
    for f in (1, 2):
        pc = s.push_constants(f)
        a = o.trace_primary_ids(pc)
        assert_ids_equal(a, e.trace_primary_ids(pc))
        if name == "coincident":
            hit = a[0] != abi.MISS_ID
            assert not np.any(a[0][hit] > 1)  # of the coincident copies the smallest instance id wins
    assert np.abs(o.render(3) - e.render(3)).max() < 2e-6


def hostile_shading_scenes():
    """out-of-range material parameters and degenerate punctual lights on the Cornell box (finite values only: with an
    infinite light intensity the reference computes inf * 0 for an occluded sample, which is outside defined behaviour)"""
    base = lambda: scenes.cornell_box(48, 48)  # noqa: E731
    out = {}
    for tag, rough, metal, alb in (("rough0_metal1", 0.0, 1.0, (1, 1, 1, 1)), ("albedo_gt_1", 1.0, 0.0, (3, 2, 5, 1)), ("albedo_0", 0.5, 0.5, (0, 0, 0, 1)),
                                   ("albedo_negative", 0.3, 0.2, (-1, 0.5, 0.5, 1)), ("rough5_metal_neg", 5.0, -1.0, (0.5, 0.5, 0.5, 1))):
        s = base()
        s.materials["roughness_metallic"][:3, 0], s.materials["roughness_metallic"][:3, 1], s.materials["albedo"][:3] = rough, metal, alb
        out[tag] = s
    for tag, rows in (("point_radius_0", [scenes.point_light((0, 1.0, 0), intensity=5.0, radius=0.0)]), ("point_radius_50", [scenes.point_light((0, 1.0, 0), intensity=5.0, radius=50.0)]),
                      ("spot_inner_eq_outer", [scenes.spot_light((0, 1.9, 0), forward=(0, 1, 0), intensity=50.0, radius=0.1, inner_deg=25.0, outer_deg=25.0)]),
                      ("spot_zero_direction", [scenes.spot_light((0, 1.9, 0), forward=(0, 0, 0), intensity=50.0, radius=0.1, inner_deg=10.0, outer_deg=60.0)]),
                      ("point_on_floor", [scenes.point_light((0.0, 0.0, 0.0), intensity=5.0)])):
        s = base()
        s.lights = scenes._stack(rows, abi.LIGHT)
        out[tag] = s
    return out


HOSTILE_SHADING = ["rough0_metal1", "albedo_gt_1", "albedo_0", "albedo_negative", "rough5_metal_neg", "point_radius_0", "point_radius_50", "spot_inner_eq_outer", "spot_zero_direction", "point_on_floor"]


@pytest.mark.parametrize("name", HOSTILE_SHADING)
def test_hostile_materials_and_lights(name, oracle_mod, emul):
    s = hostile_shading_scenes()[name]
    o, e = oracle_mod.OracleScene(s, brute_force=True), emul.EmulScene(s)
    with np.errstate(all="ignore"):
        a, b = o.render(4), e.render(4)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    assert np.abs(np.nan_to_num(a) - np.nan_to_num(b)).max() < 2e-6
    assert int(o.counters[0]) == int(e.counters[0])


def hostile_light_geometry_scenes():
    out = {}
    s = scenes.cornell_box(48, 48)
    m = s.meshes[0]
    i0 = int(m.submeshes[-1]["base_index"])  # the emissive quad; only its triangle 0 is ever sampled
    m.vertices["position"][m.indices[i0 + 1], :3] = m.vertices["position"][m.indices[i0], :3]  # zero-area sampled triangle: pdf guard max(1e-4, cos * A)
    out["area_light_degenerate"] = s
    s = scenes.cornell_box(48, 48)
    m = s.meshes[0]
    sub = m.submeshes[-1]
    for k in range(int(sub["index_count"])):
        m.vertices["position"][m.indices[int(sub["base_index"]) + k], 1] = 0.0  # light quad in the floor plane
    out["area_light_in_floor_plane"] = s
    s = scenes.terrain_scene(grid=24, n_spheres=4, sphere_level=1, width=48, height=27)
    faces = np.full((6, 8, 8, 4), 0.5, np.float32)
    faces[0, :, :, :3], faces[4, 0, 0, :3], faces[3] = -1.0, 1e30, 0.0  # negative, huge and black texels (finite)
    s.env_cube, s.sun_direction = (8, faces), None
    out["env_map_hostile_finite"] = s
    s = scenes.foliage_scene(n_clusters=30, cards_per_cluster=8, width=48, height=27, ground_grid=4, tex_size=4)
    fmt, w, h, data = s.textures[0]
    d = np.array(data).copy().reshape(h, w, 4)
    d[..., 3] = np.array([[25, 26, 25, 26]] * 4, np.uint8)  # 25/255 = 0.098, 26/255 = 0.102: around the any-hit threshold 0.1
    s.textures[0] = (fmt, w, h, d)
    out["alpha_around_threshold"] = s
    return out


@pytest.mark.parametrize("name", ["area_light_degenerate", "area_light_in_floor_plane", "env_map_hostile_finite", "alpha_around_threshold"])
def test_hostile_light_geometry_and_texels(name, oracle_mod, emul):
    s = hostile_light_geometry_scenes()[name]
    o, e = oracle_mod.OracleScene(s, brute_force=True), emul.EmulScene(s)
    with np.errstate(all="ignore"):
        a, b = o.render(4), e.render(4)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    assert np.abs(np.nan_to_num(a) - np.nan_to_num(b)).max() < 2e-6
    assert int(o.counters[0]) == int(e.counters[0])


def moved_instances(s, seed, frac=0.5, shift=6.0):
    """a copy of the scene's instance table with a random subset of the instances translated / rotated about y"""
    import math

    rng = np.random.default_rng(seed)
    inst = s.instances.copy()
    for i in range(len(inst)):
        if rng.random() > frac:
            continue
        M = inst["model_matrix"][i].reshape(4, 4).T.astype(np.float64)
        ang = rng.random() * 2 * math.pi
        c, sn = math.cos(ang), math.sin(ang)
        R = np.array([[c, 0, sn, 0], [0, 1, 0, 0], [-sn, 0, c, 0], [0, 0, 0, 1]], np.float64)
        T = np.eye(4)
        T[:3, 3] = (rng.random(3) - 0.5) * np.array([shift, 0.3 * shift, shift])
        new = scenes.make_instance(T @ M @ R, int(inst["mesh_index"][i]))
        inst["model_matrix"][i], inst["normal_matrix"][i] = new["model_matrix"], new["normal_matrix"]
    return inst


def test_instance_tree_refit_equals_rebuild(emul):
    """moving instances REFIT the instance tree (hl_build.h refit_node_box / refit_requantize) instead of
    rebuilding it (reference: TLAS created ALLOW_UPDATE, scene.cpp:797, rebuilt per change, renderer.cpp:147-168).  After three
    rounds of random moves the refitted tree must give exactly the hits and the image of a scene built from scratch with
    the same transforms, and both must equal the oracle."""
    import copy

    from oracle import oracle

    s = scenes.city_scene(n_instances=40, n_meshes=4, width=64, height=36, floors=(2, 4), detail=(1, 2))
    e = emul.EmulScene(s)
    cur = s
    for rnd in range(3):
        inst = moved_instances(cur, seed=10 + rnd)
        e.update_instances(inst)
        cur = copy.copy(cur)
        cur.instances = inst
        fresh = emul.EmulScene(cur)
        pc = cur.push_constants(1)
        a, b = e.trace_primary_ids(pc), fresh.trace_primary_ids(pc)
        for x, y in zip(a, b):
            assert np.array_equal(x.view(np.uint32), y.view(np.uint32))
        o = oracle.OracleScene(cur)
        r = o.trace_primary_ids(pc)
        for x, y in zip(a[:3], r[:3]):
            assert np.array_equal(x, y)
        assert np.array_equal(e.render(3), fresh.render(3))
    assert (a[0] != 0xFFFFFFFF).mean() > 0.3


def test_traversal_stack_overflow_is_counted_and_memory_safe(tmp_path):
    """hl_bvh.h TravStack: an entry that does not fit is dropped and COUNTED (hl_get_counters fails loudly on the GPU while the
    count is non-zero), an instance whose sentinel would not fit is skipped — nothing dropped is ever popped.  Forced here by
    compiling the device headers with a 3- and a 2-entry stack in a subprocess: the render must
    finish, report overflows, and with the normal 64-entry stack the same scene must report none."""
    import subprocess
    import sys
    import textwrap
    from pathlib import Path

    prog = textwrap.dedent(
        """
        import sys, ctypes as C
        sys.path.insert(0, %r)
        import numpy as np
        from helios_b200 import scenes
        from tests.emul import emul
        s = scenes.city_scene(n_instances=30, n_meshes=3, width=48, height=27, floors=(2, 5), detail=(2, 3))
        e = emul.EmulScene(s)
        L = emul.lib()
        L.em_stack_overflows.restype = C.c_uint64
        L.em_stack_overflows(C.c_int(1))
        a = e.render(2)
        print(int(L.em_stack_overflows(C.c_int(0))), int(np.isfinite(a).all()))
        """
        % str(Path(__file__).resolve().parent.parent)
    )
    import os

    def run(flags):
        env = dict(os.environ, OMP_NUM_THREADS="1")
        if flags:
            env["HL_EMUL_CXXFLAGS"] = flags
        else:
            env.pop("HL_EMUL_CXXFLAGS", None)
        r = subprocess.run([sys.executable, "-c", prog], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        return [int(x) for x in r.stdout.split()]

    lost, finite = run("-DHL_STACK_FAST=2 -DHL_STACK_SPILL=1")  # 3 entries: an instance fits (sentinel + 2), deeper pushes are dropped
    assert lost > 0 and finite == 1
    lost, finite = run("-DHL_STACK_FAST=1 -DHL_STACK_SPILL=1")  # 2 entries: no instance can be entered (its sentinel would not fit)
    assert lost > 0 and finite == 1
    lost, finite = run("")
    assert lost == 0 and finite == 1
