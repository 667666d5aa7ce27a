"""GPU: BASELINE configs at their FULL sizes, checked through size-independent properties (the oracle needs minutes per
frame at these sizes; the small-size tests compare against it directly):

* scheduling invariance — frames in flight, CUDA-graph replay, tail kernel, tile decomposition: same accumulation image,
  bit for bit, and the same ray counts;
* tree invariance — plain LBVH vs binned-SAH re-split (different BVHs): same primary hits (ids and t, u, v) and same image;
* spp sharding linearity — the per-"rank" SUM buffers of a round-robin frame split add up to the single-GPU sum
  (what the one NCCL all-reduce of the path computes), to fp32 summation order;
* range — accumulated radiance stays inside [0, 1] (RADIANCE_CLAMP_COLOR, rgen:219) with alpha 1."""
import numpy as np
import pytest

from helios_b200 import abi, multi_gpu, scenes

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]
FRAMES = 6


@pytest.fixture(scope="module")
def terrain():
    return scenes.terrain_scene()  # configs[1]: 1M triangles, 1920x1080, depth 8


def render(ctx, s, frames, **kw):
    ctx.accum_clear()
    ctx.reset_counters()
    for f in frames:
        ctx.render_frame(s.push_constants(f, **kw))
    acc = ctx.read_accum()
    c = ctx.counters()
    return acc, int(c["extension_rays"]), int(c["shadow_rays"])


def test_configs1_full_size_invariances(terrain):
    from helios_b200 import api

    s = terrain
    assert s.num_triangles >= 1_000_000 and (s.width, s.height) == (1920, 1080)
    ctx = api.Context(s.width, s.height)
    ctx.load_scene(s)
    ref, ext, sh = render(ctx, s, range(FRAMES))
    assert ext >= FRAMES * s.width * s.height and sh > 0
    assert np.isfinite(ref).all() and ref[..., :3].min() >= 0.0 and ref[..., :3].max() <= 1.0 and np.all(ref[..., 3] == 1.0)
    primary = ctx.trace_primary_ids(s.push_constants(1))
    settings = [
        {abi.OPT_FRAMES_IN_FLIGHT: 1}, {abi.OPT_FRAMES_IN_FLIGHT: 8}, {abi.OPT_FRAMES_IN_FLIGHT: 4, abi.OPT_CUDA_GRAPH: 0},
        {abi.OPT_CUDA_GRAPH: 1, abi.OPT_TAIL_THRESHOLD: 0}, {abi.OPT_TAIL_THRESHOLD: 1 << 22, abi.OPT_TAIL_START: 1}, {abi.OPT_TAIL_THRESHOLD: 98304, abi.OPT_TAIL_START: 4, abi.OPT_PIPELINE: 0},
    ]
    for opt in settings:
        for k, v in opt.items():
            ctx.set_option(k, v)
        acc, e, h = render(ctx, s, range(FRAMES))
        assert np.array_equal(acc, ref), opt
        assert (e, h) == (ext, sh), opt
    ctx.set_option(abi.OPT_PIPELINE, 1)
    # 128 x 128 tiles (PathIntegrator::compute_tile_coords): 15 x 9 launches per sample, the bottom row clipped
    ctx.accum_clear()
    for f in range(2):
        for ty in range(0, s.height, 128):
            for tx in range(0, s.width, 128):
                ctx.render_frame(s.push_constants(f, tile=(tx, ty)), launch=(128, 128))
    tiled = ctx.read_accum()
    full2, _, _ = render(ctx, s, range(2))
    assert np.array_equal(tiled, full2)
    # spp sharding: 4 "ranks", frames round-robin, SUM buffers -> their sum equals the one-GPU SUM of the same frames
    ctx.set_accum_mode(abi.ACCUM_SUM)
    world, per_rank = 4, 2
    all_frames = sorted(f for r in range(world) for f in multi_gpu.frame_indices(r, world, per_rank))
    assert all_frames == list(range(1, world * per_rank + 1))
    total, _, _ = render(ctx, s, all_frames)
    parts = np.zeros_like(total)
    for r in range(world):
        part, _, _ = render(ctx, s, multi_gpu.frame_indices(r, world, per_rank))
        parts += part
    ctx.set_accum_mode(abi.ACCUM_RUNNING_MEAN)
    assert np.abs(parts[..., :3] - total[..., :3]).max() <= 1e-5 * world * per_rank
    ctx.close()
    # a different tree: plain LBVH (no SAH re-split) -> same hits, same image
    ctx2 = api.Context(s.width, s.height)
    ctx2.set_option(abi.OPT_SAH_CLUSTER, 0)
    ctx2.load_scene(s)
    for a, b in zip(ctx2.trace_primary_ids(s.push_constants(1)), primary):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    acc, e, h = render(ctx2, s, range(FRAMES))
    assert np.array_equal(acc, ref) and (e, h) == (ext, sh)
    ctx2.close()


@pytest.mark.parametrize("name", ["foliage", "city"])
def test_configs2_and_3_full_size_invariances(name):
    """configs[2] (5M alpha-tested triangles, 80 lights, 1080p) and configs[3] (19.8M instanced triangles, 3840x2160):
    scheduling and tree invariance at full size"""
    from helios_b200 import api

    s = scenes.foliage_scene() if name == "foliage" else scenes.city_scene()
    assert s.num_triangles >= (5_000_000 if name == "foliage" else 19_000_000)
    frames = range(3)
    ctx = api.Context(s.width, s.height)
    ctx.load_scene(s)
    ref, ext, sh = render(ctx, s, frames)
    assert np.isfinite(ref).all() and ref[..., :3].min() >= 0.0 and ref[..., :3].max() <= 1.0
    primary = ctx.trace_primary_ids(s.push_constants(1))
    for opt in ({abi.OPT_FRAMES_IN_FLIGHT: 1, abi.OPT_CUDA_GRAPH: 0}, {abi.OPT_FRAMES_IN_FLIGHT: 8, abi.OPT_CUDA_GRAPH: 1, abi.OPT_TAIL_THRESHOLD: 0}):
        for k, v in opt.items():
            ctx.set_option(k, v)
        acc, e, h = render(ctx, s, frames)
        assert np.array_equal(acc, ref) and (e, h) == (ext, sh), opt
    ctx.close()
    ctx2 = api.Context(s.width, s.height)
    ctx2.set_option(abi.OPT_SAH_CLUSTER, 0)  # plain LBVH for every mesh and for the instance tree
    ctx2.load_scene(s)
    for a, b in zip(ctx2.trace_primary_ids(s.push_constants(1)), primary):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    acc, e, h = render(ctx2, s, frames)
    assert np.array_equal(acc, ref) and (e, h) == (ext, sh)
    ctx2.close()


@pytest.mark.parametrize("name", ["terrain", "foliage", "city"])
def test_full_size_scene_tile_against_oracle(name, oracle_mod):
    """the full-size scenes of configs[1..3], one 480 x 270 launch rectangle in the middle of the frame (tile offsets as
    PathIntegrator uses them), 3 launches: primary hit ids equal the oracle's on >= 99.99 % of the tile's rays (north_star
    bar; thin-lens scenes differ by cosf / sinf ulps), radiance within 1e-4 on >= 99 % of the pixels, relMSE < 2e-3"""
    from helios_b200 import api
    from helios_b200.sky import sky_coefficients

    s = {"terrain": scenes.terrain_scene, "foliage": scenes.foliage_scene, "city": scenes.city_scene}[name]()
    cf = sky_coefficients(s.sun_direction) if s.sun_direction is not None else None
    ctx = api.Context(s.width, s.height)
    ctx.load_scene(s, sky_coeffs=cf)
    o = oracle_mod.OracleScene(s, sky_coeffs_override=cf)
    tw, th = 480, 270
    tx, ty = (s.width - tw) // 2, (s.height - th) // 2
    acc = np.zeros((s.height, s.width, 4), np.float32)
    acc[..., 3] = 1.0
    ctx.accum_clear()
    for f in range(3):
        pc = s.push_constants(f, tile=(tx, ty))
        ctx.render_frame(pc, launch=(tw, th))
        o.render_frame(pc, acc, launch=(tw, th))
    g = ctx.read_accum()[ty:ty + th, tx:tx + tw, :3]
    r = acc[ty:ty + th, tx:tx + tw, :3]
    d = np.abs(g - r).max(-1)
    assert (d > 1e-4).mean() < 0.01, f"{int((d > 1e-4).sum())} of {d.size} pixels differ by more than 1e-4"
    mse = float(np.mean((g.astype(np.float64) - r) ** 2) / max(float(np.mean(r.astype(np.float64) ** 2)), 1e-12))
    assert mse < 2e-3, mse
    # primary hits of the whole frame through both (the oracle's BVH is conservative: tests/test_emul_parity.py)
    pc = s.push_constants(1)
    a, b = ctx.trace_primary_ids(pc), o.trace_primary_ids(pc)
    same = (a[0] == b[0]) & (a[1] == b[1]) & (a[2] == b[2])
    assert same.mean() >= 1.0 - 1e-4, f"{int((~same).sum())} of {same.size} primary hits differ"
    ctx.close()


@pytest.mark.parametrize("aperture,bias", [(0.0, 0.0), (0.1, 1e-3)])
def test_configs0_cornell_512_64spp(aperture, bias, oracle_mod):
    """BASELINE configs[0] as specified: Cornell box, 512 x 512, 65 launches (num_frames 0..64 = 64 effective
    samples, frame 0 is overwritten: A.8-1), depth 8; pinhole / bias 0 and aperture 0.1 / bias 1e-3.  Against the oracle:
    primary hit ids bit-exact (pinhole), converged image within 1e-4 on >= 99 % of the pixels and relMSE < 2e-3"""
    from helios_b200 import api

    s = scenes.cornell_box(512, 512, aperture_radius=aperture, shadow_ray_bias=bias)
    assert s.max_ray_bounces == 8 and len(s.lights) == 1
    ctx = api.Context(s.width, s.height)
    ctx.load_scene(s)
    o = oracle_mod.OracleScene(s, brute_force=True)
    if aperture == 0.0:
        for f in (0, 1, 64):
            pc = s.push_constants(f)
            for a, b in zip(ctx.trace_primary_ids(pc), o.trace_primary_ids(pc)):
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    g, r = ctx.render(s, 65), o.render(65)
    d = np.abs(g - r)[..., :3].max(-1)
    assert (d > 1e-4).mean() < 0.01, f"{int((d > 1e-4).sum())} of {d.size} pixels differ by more than 1e-4"
    mse = float(np.mean((g[..., :3].astype(np.float64) - r[..., :3]) ** 2) / np.mean(r[..., :3].astype(np.float64) ** 2))
    assert mse < 2e-3, mse
    assert 0.005 < float(g[..., :3].mean()) < 0.5  # a lit box, not a black or saturated frame
    ctx.close()
