"""GPU: edge cases of the path against the oracle through the C ABI — empty and tiny meshes, coincident and degenerate
triangles (tie rule), no lights, the instance limit of the reference (scene.h:14-17), odd image extents and clipped
tiles, bounce limits, axis-parallel rays, context resize, frames-in-flight settings."""
import numpy as np
import pytest

from helios_b200 import abi, scenes

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


@pytest.fixture(scope="module")
def api():
    from helios_b200 import api as a

    return a


def ids_equal(g, r):
    return all(np.array_equal(a.view(np.uint32), b.view(np.uint32)) for a, b in zip(g, r))


@pytest.mark.parametrize("n", [0, 1, 2, 3, 5, 9, 33])
def test_tiny_meshes(n, api, oracle_mod):
    s = scenes.triangle_soup(max(n, 1), 48, 27, seed=100 + n)
    if n == 0:  # empty geometry: a submesh with zero triangles
        s.meshes[0].submeshes[0]["index_count"] = 0
    ctx = api.Context(s.width, s.height)
    ctx.load_scene(s)
    o = oracle_mod.OracleScene(s, brute_force=True)
    pc = s.push_constants(1)
    assert ids_equal(ctx.trace_primary_ids(pc), o.trace_primary_ids(pc))
    a, b = ctx.render(s, 3), o.render(3)
    assert np.abs(a - b)[..., :3].max() < 1e-4
    ctx.close()


def test_coincident_and_degenerate_triangles(api, oracle_mod):
    """equal t -> the lexicographically smallest (instance, geometry, primitive) wins, whatever the tree; zero-area
    triangles are never hit"""
    s = scenes.triangle_soup(300, 64, 36, seed=7)
    m = s.meshes[0]
    v, idx = m.vertices, m.indices
    # triangles 100..199 become exact copies of 0..99 (same vertices through the index buffer); 200..249 collapse to a point / a line
    idx[300:600] = idx[0:300]
    v["position"][idx[600:675], :3] = v["position"][idx[600], :3]
    for t in range(225, 250):
        v["position"][idx[3 * t + 2], :3] = 0.5 * (v["position"][idx[3 * t], :3] + v["position"][idx[3 * t + 1], :3])
    ctx = api.Context(s.width, s.height)
    ctx.load_scene(s)
    o = oracle_mod.OracleScene(s, brute_force=True)
    for frame in (0, 1, 2):
        pc = s.push_constants(frame)
        g, r = ctx.trace_primary_ids(pc), o.trace_primary_ids(pc)
        assert ids_equal(g, r)
        prim = g[2][g[0] != abi.MISS_ID]
        assert not np.any((prim >= 100) & (prim < 200)), "a duplicate with the larger primitive id won a tie"
        assert not np.any((prim >= 200) & (prim < 225)), "a point-sized triangle was hit"
    ctx.close()


def test_no_lights(api, oracle_mod):
    """num_lights = 0: next_uint(rng, 0) = 0 indexes an empty table; the term is multiplied by num_lights = 0 (SURVEY C-4)"""
    s = scenes.cornell_box(64, 64)
    s.lights = s.lights[:0]
    ctx = api.Context(s.width, s.height)
    ctx.load_scene(s)
    o = oracle_mod.OracleScene(s, brute_force=True)
    a, b = ctx.render(s, 4), o.render(4)
    assert np.isfinite(a).all() and np.abs(a - b)[..., :3].max() < 1e-4
    assert a[..., :3].max() > 0  # the emissive quad is still seen directly
    ctx.close()


def test_instance_limit_1024(api, oracle_mod):
    """MAX_SCENE_MESH_INSTANCE_COUNT = 1024 (include/resource/scene.h:14): a full table of instances of two small meshes"""
    s = scenes.city_scene(n_instances=1023, n_meshes=2, width=96, height=54, floors=(1, 2), detail=(1, 1))  # + the ground instance
    assert len(s.instances) == 1024
    ctx = api.Context(s.width, s.height)
    ctx.load_scene(s)
    o = oracle_mod.OracleScene(s)
    pc = s.push_constants(1)
    g, r = ctx.trace_primary_ids(pc), o.trace_primary_ids(pc)
    same = (g[0] == r[0]) & (g[1] == r[1]) & (g[2] == r[2])
    assert same.mean() >= 1 - 1e-4
    assert len(np.unique(g[0][g[0] != abi.MISS_ID])) > 20
    ctx.close()


@pytest.mark.parametrize("w,h", [(1, 1), (3, 5), (257, 3), (130, 129)])
def test_odd_extents_and_clipped_tiles(w, h, api, oracle_mod):
    s = scenes.cornell_box(w, h)
    ctx = api.Context(w, h)
    ctx.load_scene(s)
    o = oracle_mod.OracleScene(s, brute_force=True)
    pc = s.push_constants(1)
    assert ids_equal(ctx.trace_primary_ids(pc), o.trace_primary_ids(pc))
    full = ctx.render(s, 3)
    assert np.abs(full - o.render(3))[..., :3].max() < 1e-4
    # 128^2 tiles like PathIntegrator::compute_tile_coords; the last row / column of tiles hangs over the image edge
    ctx.accum_clear()
    for f in range(3):
        for ty in range(0, h, 128):
            for tx in range(0, w, 128):
                ctx.render_frame(s.push_constants(f, tile=(tx, ty)), launch=(128, 128))
    assert np.array_equal(ctx.read_accum(), full)
    ctx.close()


@pytest.mark.parametrize("bounces", [0, 1, 2, 64, 100])
def test_bounce_limits(bounces, api, oracle_mod):
    """max_ray_bounces = 0 still traces the primary ray (the raygen shader does, rgen:205); 1 = direct light only; more
    than 64 is refused (HL_ERR_LIMIT; the reference's UI offers 1..8)"""
    s = scenes.cornell_box(48, 48)
    ctx = api.Context(s.width, s.height)
    ctx.load_scene(s)
    o = oracle_mod.OracleScene(s, brute_force=True)
    if bounces > 64:
        from helios_b200._lib import HeliosError

        with pytest.raises(HeliosError, match="max_ray_bounces > 64"):
            ctx.render_frame(s.push_constants(0, max_ray_bounces=bounces))
        ctx.close()
        return
    ctx.accum_clear()
    acc = np.zeros((s.height, s.width, 4), np.float32)
    acc[..., 3] = 1.0
    for f in range(3):
        pc = s.push_constants(f, max_ray_bounces=bounces)
        ctx.render_frame(pc)
        o.render_frame(pc, acc)  # in place
    a = ctx.read_accum()
    assert np.abs(a - acc)[..., :3].max() < 1e-4, float(np.abs(a - acc)[..., :3].max())
    ctx.close()


@pytest.mark.parametrize("flags", [0, 1, 3])
def test_axis_parallel_and_grazing_rays(flags, api, oracle_mod):
    s = scenes.foliage_scene(n_clusters=40, cards_per_cluster=10, width=32, height=18, ground_grid=6, tex_size=16)
    ctx = api.Context(s.width, s.height)
    ctx.load_scene(s)
    o = oracle_mod.OracleScene(s)
    rng = np.random.default_rng(10 + flags)
    n = 20000
    org = (rng.random((n, 3)) - 0.5) * np.array([50, 4, 50]) + np.array([0, 4, 0])
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[::7] = np.eye(3)[rng.integers(0, 3, len(d[::7]))] * rng.choice([-1.0, 1.0], (len(d[::7]), 1))  # exactly along an axis
    d[1::50, 0] = 0.0
    d[2::77, 1] = 1e-12
    d[3::91, 1] = -0.0
    org[4::13, 1] = 0.0  # origins on the ground plane
    rays = np.concatenate([org, np.full((n, 1), 1e-3), d, np.full((n, 1), 30.0)], 1).astype(np.float32)
    rays[5::101, 7] = 0.0  # empty interval: tmax < tmin
    a, b = ctx.trace_rays(rays, flags), o.trace_rays(rays, flags)
    if flags & 2:  # terminate on first hit: only hit / miss is defined
        assert np.array_equal(np.isinf(a[:, 0]), np.isinf(b[:, 0]))
    else:
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert np.isfinite(a[:, 0]).mean() > 0.05
    ctx.close()


def test_resize_and_frames_in_flight(api, oracle_mod):
    """hl_context_resize (Renderer::on_window_resize) and HL_OPT_FRAMES_IN_FLIGHT 1..8: same image for every setting,
    ray counters survive a change of the setting, bad values are refused"""
    from helios_b200._lib import HeliosError

    s0 = scenes.cornell_box(40, 40)
    ctx = api.Context(s0.width, s0.height)
    ctx.load_scene(s0)
    ctx.render(s0, 2)
    s = scenes.cornell_box(96, 64)
    ctx.resize(s.width, s.height)
    ref = None
    total = 0
    ctx.reset_counters()
    for n in (4, 1, 8, 2, 3):
        ctx.set_option(abi.OPT_FRAMES_IN_FLIGHT, n)
        a = ctx.render(s, 9)
        total += 9
        if ref is None:
            ref = a
            assert np.abs(a - oracle_mod.OracleScene(s, brute_force=True).render(9))[..., :3].max() < 1e-4
        else:
            assert np.array_equal(a, ref), n
        c = ctx.counters()
        assert int(c["extension_rays"]) >= total * s.width * s.height
    for bad in (0, 9, -1):
        with pytest.raises(HeliosError):
            ctx.set_option(abi.OPT_FRAMES_IN_FLIGHT, bad)
    ctx.close()


def test_cuda_graph_replay_follows_every_setting(api):
    """HL_OPT_CUDA_GRAPH: the bounce loop replayed from a per-slot CUDA graph must be re-captured whenever something its
    launches carry by value changes — bounce limit, shadow bias, number of lights, the scene tables, the extent, the tail
    settings — and give the image of the plain launch sequence"""

    def run(ctx, s, n, **kw):
        ctx.accum_clear()
        for f in range(n):
            ctx.render_frame(s.push_constants(f, **kw))
        return ctx.read_accum()

    a = scenes.cornell_box(64, 48)
    b = scenes.foliage_scene(n_clusters=40, cards_per_cluster=10, width=64, height=48, ground_grid=6, tex_size=16)
    g, p = api.Context(64, 48), api.Context(64, 48)
    p.set_option(abi.OPT_CUDA_GRAPH, 0)
    steps = [(a, {}), (a, {"max_ray_bounces": 3}), (a, {"max_ray_bounces": 3, "shadow_ray_bias": 1e-3}), (b, {}), (b, {"max_ray_bounces": 2}), (a, {})]
    loaded = None
    for s, kw in steps:
        if s is not loaded:
            g.load_scene(s), p.load_scene(s)
            loaded = s
        assert np.array_equal(run(g, s, 9, **kw), run(p, s, 9, **kw)), kw
    g.set_option(abi.OPT_TAIL_START, 1), p.set_option(abi.OPT_TAIL_START, 1)
    assert np.array_equal(run(g, a, 9), run(p, a, 9))
    g.resize(80, 40), p.resize(80, 40)
    a2 = scenes.cornell_box(80, 40)
    assert np.array_equal(run(g, a2, 9), run(p, a2, 9))
    assert g.kernel_launches() == p.kernel_launches()  # graph nodes are counted like plain launches
    g.close(), p.close()


def test_grazing_rays_never_lose_an_edge_hit(api, oracle_mod):
    """tests/graze.py on the GPU: 2e6 adversarial grazing rays through the LBVH tree and the SAH re-split tree; both must
    return the oracle's hit (its BVH is checked against brute force in tests/test_emul_parity.py) bit for bit"""
    from tests import graze

    s, tris = graze.graze_scene()
    rays = graze.graze_rays(tris, 2_000_000)
    ref = oracle_mod.OracleScene(s).trace_rays(rays, 0)
    for cluster in (0, 2):
        ctx = api.Context(s.width, s.height)
        ctx.set_option(abi.OPT_SAH_CLUSTER, cluster)
        ctx.load_scene(s)
        h = ctx.trace_rays(rays, 0)
        bad = (h.view(np.uint32) != ref.view(np.uint32)).any(1)
        assert not bad.any(), f"SAH cluster {cluster}: {int(bad.sum())} of {len(rays)} grazing rays differ"
        ctx.close()


@pytest.mark.timeout(120)
def test_hostile_rays_terminate_and_match(api, oracle_mod):
    """zero / NaN / denormal / huge directions, infinite origins, NaN intervals (checked on the CPU emulation first:
    tests/test_emul_parity.py): the kernels terminate and return the oracle's answer"""
    from tests.test_emul_parity import hostile_rays

    s = scenes.city_scene(n_instances=30, n_meshes=3, width=96, height=54, floors=(2, 4), detail=(1, 3))
    ctx = api.Context(s.width, s.height)
    ctx.load_scene(s)
    rays = hostile_rays(20000)
    a, b = ctx.trace_rays(rays, 0), oracle_mod.OracleScene(s).trace_rays(rays, 0)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    ctx.close()


@pytest.mark.timeout(120)
@pytest.mark.parametrize("name", ["one_vertex_1e30", "triangle_1e30", "triangle_-3e38", "nan_inf_huge_mixed"])
def test_extreme_vertices(name, api, oracle_mod):
    """finite but absurd (1e30, -3e38) and non-finite vertex positions: the GPU builder stays valid for both tree builds and
    the hits equal the oracle's brute force (checked on the CPU emulation first, tests/test_emul_parity.py)"""
    from tests.test_emul_parity import extreme_vertex_scenes

    s = extreme_vertex_scenes()[name]
    pc = s.push_constants(1)
    with np.errstate(all="ignore"):
        ref = oracle_mod.OracleScene(s, brute_force=True).trace_primary_ids(pc)
    for cluster in (2, 0):
        ctx = api.Context(s.width, s.height)
        ctx.set_option(abi.OPT_SAH_CLUSTER, cluster)
        ctx.load_scene(s)
        assert ids_equal(ctx.trace_primary_ids(pc), ref), cluster
        ctx.close()


@pytest.mark.parametrize("name", ["non_rigid", "coincident", "far_from_origin"])
def test_instance_transform_edge_cases(name, api, oracle_mod):
    """two-level traversal with non-uniform scale / shear / mirrored / arbitrary instance transforms, coincident instances
    (tie rule across instances) and a scene 3e5 units from the origin: hits equal the oracle's brute force"""
    from tests.test_emul_parity import instance_transform_scenes

    s = instance_transform_scenes()[name]
    ctx = api.Context(s.width, s.height)
    ctx.load_scene(s)
    o = oracle_mod.OracleScene(s, brute_force=True)
    for f in (1, 2):
        pc = s.push_constants(f)
        assert ids_equal(ctx.trace_primary_ids(pc), o.trace_primary_ids(pc))
    rays = np.concatenate([np.random.default_rng(1).normal(size=(20000, 3)) * 20 + np.asarray(s.camera.position), np.full((20000, 1), 1e-3),
                           np.random.default_rng(2).normal(size=(20000, 3)), np.full((20000, 1), 1e4)], 1).astype(np.float32)
    assert np.array_equal(ctx.trace_rays(rays, 0).view(np.uint32), o.trace_rays(rays, 0).view(np.uint32))
    ctx.close()


@pytest.mark.parametrize("name", ["rough0_metal1", "albedo_gt_1", "albedo_negative", "rough5_metal_neg", "point_radius_50", "spot_inner_eq_outer", "spot_zero_direction", "point_on_floor"])
def test_hostile_materials_and_lights(name, api, oracle_mod):
    """out-of-range material parameters and degenerate punctual lights: same NaN pattern and radiance as the oracle"""
    from tests.test_emul_parity import hostile_shading_scenes

    s = hostile_shading_scenes()[name]
    ctx = api.Context(s.width, s.height)
    ctx.load_scene(s)
    o = oracle_mod.OracleScene(s, brute_force=True)
    with np.errstate(all="ignore"):
        a, b = ctx.render(s, 4), o.render(4)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    d = np.abs(np.nan_to_num(a) - np.nan_to_num(b))[..., :3].max(-1)
    assert (d > 1e-4).mean() < 0.01 and d.max() < 0.2, (int((d > 1e-4).sum()), float(d.max()))
    ctx.close()


def test_scene_tables_reject_indices_a_shader_would_follow_out_of_bounds(api):
    """hl_scene_set_tables validates every index the kernels dereference: material texture slots, and the
    instance / material / primitive range of area-light rows (sample_light, path_trace_rchit.glsl:376-451 reads them from
    LightData).  A bad one must come back as HL_ERR_INVALID_ARGUMENT — not as a sticky cudaErrorIllegalAddress — and the
    context must keep working afterwards."""
    from helios_b200._lib import HeliosError

    s = scenes.cornell_box(64, 64)
    ctx = api.Context(s.width, s.height)
    handles = ctx.load_scene(s)
    good = ctx.render(s, 2)
    meshes = [handles[int(i["mesh_index"])] for i in s.instances]

    def attempt(materials=None, lights=None):
        with pytest.raises(HeliosError) as e:
            ctx.set_tables(s.materials if materials is None else materials, s.instances, meshes, s.submesh_info, s.lights if lights is None else lights)
        assert e.value.status == 1, e.value  # HL_ERR_INVALID_ARGUMENT

    for slot, field in [(0, "texture_indices0"), (1, "texture_indices0"), (2, "texture_indices0"), (3, "texture_indices0"), (0, "texture_indices1")]:
        for bad in (0, 7, -2):  # the scene has no textures: any index but -1 is stale
            m = s.materials.copy()
            m[field][1, slot] = bad
            attempt(materials=m)
    area = [k for k, l in enumerate(s.lights) if int(l["light_data0"][0]) == abi.LIGHT_AREA]
    assert area
    k = area[0]
    n_tris = len(s.meshes[0].indices) // 3
    for field, comp, bad in [("light_data0", 1, 5.0), ("light_data0", 1, -1.0), ("light_data0", 2, 4096.0), ("light_data0", 3, float(n_tris)), ("light_data1", 2, float(n_tris + 1)),
                             ("light_data0", 3, float("nan")), ("light_data1", 2, 3e9)]:
        l = s.lights.copy()
        l[field][k, comp] = bad
        attempt(lights=l)
    # the failed calls changed nothing a frame depends on: the old tables are still installed and render the same image
    ctx.set_tables(s.materials, s.instances, meshes, s.submesh_info, s.lights)
    assert np.array_equal(ctx.render(s, 2), good)
    ctx.close()


def test_textures_clear_makes_material_indices_stale(api):
    from helios_b200._lib import HeliosError

    s = scenes.terrain_scene(grid=24, n_spheres=2, sphere_level=1, width=64, height=36, textured=True)
    ctx = api.Context(s.width, s.height)
    handles = ctx.load_scene(s)
    ctx.render(s, 1)
    ctx._chk(ctx.lib.hl_textures_clear(ctx.h))
    with pytest.raises(HeliosError):
        ctx.set_tables(s.materials, s.instances, [handles[int(i["mesh_index"])] for i in s.instances], s.submesh_info, s.lights)
    ctx.close()


def test_stack_overflow_fails_hl_get_counters_loudly(tmp_path):
    """VERDICT r1 / ADVICE: a traversal that runs out of stack must not be silent.  tests/_variants/libhelios_b200_stack3.so is the
    product library compiled with a 3-entry stack (__graft_entry__.build()); in a subprocess (HELIOS_B200_LIB) a city render
    overflows, finishes without a CUDA error, and hl_get_counters returns HL_ERR_LIMIT with the count in the message; after
    hl_reset_counters and a scene that fits (one shallow mesh) the call succeeds again."""
    import os
    import subprocess
    import sys
    import textwrap
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    lib = root / "tests" / "_variants" / "libhelios_b200_stack3.so"
    if not lib.exists():
        pytest.skip("test variant not built (python __graft_entry__.py)")
    prog = textwrap.dedent(
        """
        import sys
        sys.path.insert(0, %r)
        import numpy as np
        from helios_b200 import api, scenes
        from helios_b200._lib import HeliosError
        s = scenes.city_scene(n_instances=30, n_meshes=3, width=96, height=54, floors=(2, 5), detail=(2, 3))
        ctx = api.Context(s.width, s.height)
        ctx.load_scene(s)
        a = ctx.render(s, 2)
        assert np.isfinite(a).all()
        try:
            ctx.counters()
            print("no error")
        except HeliosError as e:
            print("status", e.status, "overflows" in str(e))
        ctx.reset_counters()
        ctx.synchronize()
        ctx.counters()
        print("reset ok")
        ctx.close()
        """
        % str(root)
    )
    r = subprocess.run([sys.executable, "-c", prog], capture_output=True, text=True, env=dict(os.environ, HELIOS_B200_LIB=str(lib)), timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "status 5 True" in r.stdout and "reset ok" in r.stdout, r.stdout  # HL_ERR_LIMIT


def test_treelet_fallback_fit_builds_the_same_tree():
    """hl_builder.cu k_treelets keeps the cost rows of a treelet's warp-built nodes in shared memory (HL_TREELET_ROWS of them) and
    fits a treelet that has more through global memory.  tests/_variants/libhelios_b200_rows4.so (4 rows: nearly every treelet
    takes that path) must build the same trees — wide-node count and SAH cost — and return the same hits as the product library."""
    import json
    import os
    import subprocess
    import sys
    import textwrap
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    lib = root / "tests" / "_variants" / "libhelios_b200_rows4.so"
    if not lib.exists():
        pytest.skip("test variant not built (python __graft_entry__.py)")
    prog = textwrap.dedent(
        """
        import sys, json, hashlib
        sys.path.insert(0, %r)
        import numpy as np
        from helios_b200 import api, scenes
        out = []
        for s in (scenes.triangle_soup(120_000, 160, 90), scenes.foliage_scene(n_clusters=2000, width=160, height=90)):
            ctx = api.Context(s.width, s.height)
            handles = ctx.load_scene(s)
            st = [ctx.mesh_build_stats(h) for h in handles]
            ids = ctx.trace_primary_ids(s.push_constants(1))
            out.append({"nodes": [int(x["wide_nodes"]) for x in st], "sah": [float(x["sah_cost"]) for x in st],
                        "hits": hashlib.sha256(b"".join(np.ascontiguousarray(a).tobytes() for a in ids)).hexdigest()})
            ctx.close()
        print("RESULT " + json.dumps(out))
        """
        % str(root)
    )
    res = []
    plain = {k: v for k, v in os.environ.items() if k != "HELIOS_B200_LIB"}
    for env in (plain, dict(plain, HELIOS_B200_LIB=str(lib))):
        r = subprocess.run([sys.executable, "-c", prog], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        res.append(json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][0][7:]))
    assert res[0] == res[1], f"{res[0]} vs {res[1]}"
