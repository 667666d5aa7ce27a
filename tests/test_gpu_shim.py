"""GPU: the C++ host layer (helios_b200/shim) drives the same CUDA path as the Python host; the reference's frame
loop (setup -> Scene::update -> Renderer::render) through helios_headless must reproduce the image of the
parity-tested Python path on the same scene."""
import json
import subprocess

import numpy as np
import pytest

from helios_b200 import scene_io, scenes
from helios_b200.build import build_shim

pytestmark = pytest.mark.gpu


def python_render(s, launches):
    from helios_b200 import api

    ctx = api.Context(s.width, s.height)
    ctx.load_scene(s)
    acc = ctx.render(s, launches)
    ctx.close()
    return acc


def headless_render(s, tmp_path, spp, extra=(), out="out.ppm"):
    exe = str(build_shim())
    f, a, img = tmp_path / "s.hlsc", tmp_path / "acc.f32", tmp_path / out
    scene_io.export_scene(s, f)
    r = subprocess.run([exe, "--scene", str(f), "--spp", str(spp), "--dump-accum", str(a), "--out", str(img), *extra], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    stats = json.loads(r.stdout.strip().splitlines()[-1])
    acc = np.fromfile(a, np.float32).reshape(s.height, s.width, 4)
    return acc, stats, img


@pytest.mark.parametrize("name", ["cornell", "foliage", "city"])
def test_headless_matches_python_host(name, tmp_path):
    s = {
        "cornell": lambda: scenes.cornell_box(128, 128),
        "foliage": lambda: scenes.foliage_scene(n_clusters=200, cards_per_cluster=10, width=128, height=72, ground_grid=8, tex_size=32),
        "city": lambda: scenes.city_scene(n_instances=30, n_meshes=3, width=128, height=72, floors=(2, 4), detail=(1, 3)),
    }[name]()
    spp = 6
    acc, stats, img = headless_render(s, tmp_path, spp)
    ref = python_render(s, spp)  # launches num_frames = 0..spp-1, as PathIntegrator::render counts them
    assert stats["launches"] == spp and stats["extension_rays"] >= spp * s.width * s.height
    # tables agree to float rounding (tests/test_shim_host.py); camera vectors differ in the last ulp, which moves a
    # few primary rays across triangle edges: compare the images statistically, and most pixels exactly
    d = np.abs(acc[..., :3] - ref[..., :3]).max(-1)
    assert (d > 1e-3).mean() < 0.02, (d > 1e-3).mean()
    assert abs(float(acc[..., :3].mean()) - float(ref[..., :3].mean())) < 2e-3 * max(1.0, float(ref[..., :3].mean()))
    head = img.read_bytes()[:20].split(b"\n")
    assert head[0] == b"P6" and head[1] == f"{s.width} {s.height}".encode()


def test_tiled_bake_equals_full_frame(tmp_path):
    """PathIntegrator::set_tiled(true): 128x128 tiles, max_samples per tile, tile after tile (path_integrator.cpp:48-84,
    :312-336); every pixel still receives launches num_frames = 0..spp-1, so the image is the full-frame one"""
    s = scenes.cornell_box(256, 192)
    full, _, _ = headless_render(s, tmp_path, 4)
    tiled, stats, _ = headless_render(s, tmp_path, 4, extra=("--tiled",))
    assert stats["launches"] == 4 * 2 * 2
    assert np.array_equal(full, tiled)


def test_save_image_to_disk_png(tmp_path):
    """Renderer::save_image_to_disk writes the tone-mapped image as 8-bit RGBA PNG (renderer.cpp:637-711, :651);
    decoded with the independent decoder of tests/test_png_writer.py it must equal the .ppm of the same render"""
    from tests.test_png_writer import decode_png

    s = scenes.cornell_box(160, 96)
    _, _, ppm = headless_render(s, tmp_path, 3)
    _, _, png = headless_render(s, tmp_path, 3, out="out.png")
    raw = ppm.read_bytes()
    rgb = np.frombuffer(raw[len(raw) - 160 * 96 * 3 :], np.uint8).reshape(96, 160, 3)
    img = decode_png(png.read_bytes())
    assert img.shape == (96, 160, 4) and np.array_equal(img[..., :3], rgb) and rgb.max() > 0


@pytest.mark.parametrize("name", ["cornell", "foliage"])
def test_asset_route_renders_like_the_direct_route(name, tmp_path):
    """the scene written as AssetCore files (scene JSON -> mesh .ast -> material JSON -> image .ast) and
    loaded through ResourceManager::load_scene renders like the scene built through the engine API directly"""
    from helios_b200 import ast_io

    s = {
        "cornell": lambda: scenes.cornell_box(128, 96),
        "foliage": lambda: scenes.foliage_scene(n_clusters=200, cards_per_cluster=10, width=128, height=72, ground_grid=8, tex_size=32),
    }[name]()
    direct, _, _ = headless_render(s, tmp_path, 5)
    rel = ast_io.export_assets(s, tmp_path, name)
    exe, a = str(build_shim()), tmp_path / "ast.f32"
    r = subprocess.run([exe, "--ast-scene", rel, "--asset-root", str(tmp_path), "--width", str(s.width), "--height", str(s.height), "--focal-length", str(s.camera.focal_length), "--aperture",
                        str(s.camera.aperture_radius), "--bounces", str(s.max_ray_bounces), "--spp", "5", "--dump-accum", str(a), "--out", str(tmp_path / "ast.png")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    acc = np.fromfile(a, np.float32).reshape(s.height, s.width, 4)
    if name == "cornell":
        assert np.array_equal(acc, direct)  # identical tables and push constants (tests/test_ast_loader.py) -> identical image
    else:
        d = np.abs(acc[..., :3] - direct[..., :3]).max(-1)
        assert (d > 1e-3).mean() < 0.03, (d > 1e-3).mean()
        assert abs(float(acc[..., :3].mean()) - float(direct[..., :3].mean())) < 3e-3 * max(1.0, float(direct[..., :3].mean()))


def test_output_buffer_views(tmp_path):
    """Renderer::set_current_output_buffer + read_output_buffer (include/gfx/renderer.h:25-33) — the
    albedo view of the Cornell box shows exactly the material table's albedos (and the clear colour off-surface)"""
    s = scenes.cornell_box(96, 96)
    ob = tmp_path / "albedo.f32"
    headless_render(s, tmp_path, 1, extra=("--output-buffer", "albedo", "--dump-output-buffer", str(ob)))
    img = np.fromfile(ob, np.float32).reshape(s.height, s.width, 4)
    assert np.all(img[..., 3] == 1.0)
    table = np.float32([m["albedo"][:3] for m in s.materials] + [[0.0, 0.0, 0.0]])
    seen = np.unique(img[..., :3].reshape(-1, 3), axis=0)
    assert 3 <= len(seen) <= len(table)
    assert all(np.abs(table - c).max(-1).min() < 1e-6 for c in seen)  # the two hosts' tables agree to float rounding


def test_ray_debug_view_through_renderer(tmp_path):
    """Renderer::add_ray_debug_view -> render() -> PathIntegrator::gather_debug_rays (renderer.cpp:229-250): segments of
    25 paths through the image centre, 7 bounces (the integrator's default), consecutive segments of a path chained"""
    s = scenes.cornell_box(96, 96)
    f = tmp_path / "rays.f32"
    headless_render(s, tmp_path, 1, extra=("--ray-debug", "48", "48", "25", "--dump-ray-debug", str(f)))
    v = np.fromfile(f, np.float32).reshape(-1, 2, 8)
    assert 0 < len(v) <= 1024
    assert np.array_equal(v[:, 0, 4:], v[:, 1, 4:]) and np.all(v[..., 3] == 1.0)
    colours = np.unique(v[:, 0, 4:7], axis=0)
    assert len(colours) <= 25 and np.all((colours >= 0.5) & (colours < 1.0))
    for c in colours[:5]:  # inside the closed box every secondary ray hits: each segment starts where another one ended
        seg = v[(v[:, 0, 4:7] == c).all(-1)]
        starts, ends = seg[:, 0, :3], seg[:, 1, :3]
        linked = [(np.abs(ends - st).max(-1) < 1e-4).any() for st in starts]
        assert sum(linked) >= len(seg) - 1


def test_headless_sharded_over_ranks_matches_one_rank(tmp_path):
    """SURVEY 8e through the C++ host: helios_headless --devices 0,0 runs MultiGpuRenderer — two Backend / Scene / Renderer
    sets (here on the same GPU, so the driver's one-GPU box runs it), renderer g takes frames g+1, g+3, ... into a SUM image,
    hl_multi_gpu_resolve combines them over peer memory.  sum / spp must equal the running mean one renderer builds over
    launches 0..spp (frame 0 is discarded by the blend, path_trace_rgen.glsl:219-247) up to fp32 summation order."""
    s = scenes.cornell_box(128, 96)
    spp = 8
    one, _, _ = headless_render(s, tmp_path, spp + 1)  # launches num_frames = 0..spp: the mean of frames 1..spp
    acc, stats, img = headless_render(s, tmp_path, spp, extra=("--devices", "0,0"), out="multi.ppm")
    assert stats["ranks"] == 2 and stats["launches"] == spp
    d = np.abs(acc[..., :3] / spp - one[..., :3])
    assert d.max() < 2e-6 * spp, d.max()
    raw = img.read_bytes()
    rgb = np.frombuffer(raw[len(raw) - s.width * s.height * 3 :], np.uint8)
    assert rgb.max() > 0


def test_moving_nodes_refit_the_instance_tree(tmp_path):
    """a transform-only change goes through hl_scene_update_instances (instance-tree refit) in Scene::update.
    Run A moves the mesh nodes before the first frame (everything is built for the moved scene); run B renders one frame,
    then moves the same nodes by the same offsets — the shim refits — and the restarted bake must give A's image bit for bit."""
    s = scenes.city_scene(n_instances=30, n_meshes=3, width=128, height=72, floors=(2, 4), detail=(1, 3))
    a, sa, _ = headless_render(s, tmp_path, 4, extra=("--jitter-nodes", "5", "--jitter-after", "0"))
    b, sb, _ = headless_render(s, tmp_path, 4, extra=("--jitter-nodes", "5", "--jitter-after", "1"))
    still, _, _ = headless_render(s, tmp_path, 4)
    assert np.array_equal(a, b)
    assert not np.array_equal(a, still)  # the nodes really moved
