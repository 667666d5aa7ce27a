"""The C-ABI library: loads without a GPU, exports every symbol include/helios_b200.h declares, struct layouts
match the reference's shader ABI; host-side table logic of the scene generators."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from helios_b200 import abi, scenes

ROOT = Path(__file__).resolve().parent.parent


def test_struct_sizes_match_reference_abi():
    assert abi.VERTEX.itemsize == 80  # include/resource/mesh.h:10-17
    assert abi.MATERIAL.itemsize == 80  # scene.cpp:25-32
    assert abi.LIGHT.itemsize == 64  # scene.cpp:36-42
    assert abi.INSTANCE.itemsize == 144  # scene.cpp:46-52
    assert abi.PUSH_CONSTANTS.itemsize == 192  # path_integrator.cpp:11-28
    assert abi.PUSH_CONSTANTS.fields["launch_id_size"][1] == 144 and abi.PUSH_CONSTANTS.fields["num_frames"][1] == 168


def test_library_exports_every_declared_symbol():
    from helios_b200 import _lib
    from helios_b200.build import build_library

    build_library()
    header = (ROOT / "include" / "helios_b200.h").read_text()
    declared = sorted(set(re.findall(r"HL_API\s+[\w\s\*]+?\b(hl_\w+)\s*\(", header)))
    assert len(declared) >= 28
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(_lib.SYMBOLS) == declared
    assert b"sm_100a" in lib.hl_version()


def test_no_cpu_fallback_without_device():
    """hl_context_create must fail loudly when there is no CUDA device (never fall back to the CPU)"""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from helios_b200 import api
    from helios_b200._lib import HeliosError

    with pytest.raises(HeliosError) as e:
        api.Context(16, 16)
    assert e.value.status == 2 and "no CUDA device" in str(e.value)


def test_product_package_never_imports_the_oracle():
    for f in (ROOT / "helios_b200").rglob("*"):
        if f.suffix in (".py", ".cu", ".h", ".cpp"):
            txt = f.read_text()
            assert "oracle" not in txt.replace("the oracle", "").replace("CPU oracle", "").replace("oracle's", "").lower() or f.name in ("scenes.py",), f


def test_light_table_order_and_area_lights():
    """Scene::create_gpu_resources (scene.cpp:915-1311): area lights first, then env, directional, point, spot"""
    s = scenes.foliage_scene(n_clusters=4, cards_per_cluster=2, width=16, height=9, ground_grid=2, tex_size=8)
    t = s.lights["light_data0"][:, 0].astype(int).tolist()
    assert t == [abi.LIGHT_AREA] * 64 + [abi.LIGHT_POINT] * 8 + [abi.LIGHT_SPOT] * 8
    l0 = s.lights[0]
    assert l0["light_data0"][1] == 2 and l0["light_data1"][0] == 2 and l0["light_data1"][2] == 0  # count in .x, shader reads .z (A.8-2)
    assert [int(l["light_data0"][3]) for l in s.lights[:3]] == [0, 2, 4]  # base_index / 3
    t2 = scenes.terrain_scene(grid=4, n_spheres=2, sphere_level=0, width=16, height=9)
    assert t2.lights["light_data0"][:, 0].astype(int).tolist() == [abi.LIGHT_ENVIRONMENT_MAP, abi.LIGHT_DIRECTIONAL]
    assert np.allclose(t2.lights[1]["light_data1"][:3], -t2.sun_direction, atol=1e-6)  # forward() = -sun


def test_vertex_w_is_submesh_index_and_albedo_is_linearised():
    s = scenes.cornell_box(8, 8)
    m = s.meshes[0]
    for g, sub in enumerate(m.submeshes):
        idx = m.indices[sub["base_index"] : sub["base_index"] + sub["index_count"]]
        assert np.all(m.vertices["position"][idx, 3] == g)  # resource_manager.cpp:467-473
    assert np.allclose(s.materials[1]["albedo"][:3], np.power(np.float32([0.65, 0.05, 0.05]), np.float32(2.2)))  # scene.cpp:1015-1017
    # the emissive submesh's triangle 0 faces down
    sub = m.submeshes[3]
    tri = m.vertices["position"][m.indices[sub["base_index"] : sub["base_index"] + 3], :3]
    assert np.cross(tri[1] - tri[0], tri[2] - tri[0])[1] < 0


def test_push_constants_follow_launch_rays():
    s = scenes.cornell_box(64, 32)
    pc = s.push_constants(5, tile=(16, 8))
    assert pc["launch_id_size"].tolist() == [16, 8, 64, 32] and pc["num_frames"] == 5 and pc["num_lights"] == 1
    f = -s.camera.forward
    fp = pc["focal_plane"]
    assert np.allclose(fp[:3], -f) and abs(fp[3] + np.dot(-f, s.camera.position + f * s.camera.focal_length)) < 1e-5  # path_integrator.cpp:140-142
    vpi = pc["view_proj_inverse"].reshape(4, 4).T
    centre = vpi @ np.array([0, 0, 0, 1.0])
    centre = centre[:3] / centre[3]
    d = centre - s.camera.position
    assert np.allclose(d / np.linalg.norm(d), f, atol=1e-5)


def test_scene_triangle_counts():
    assert scenes.cornell_box().num_triangles == 36
    assert 2 * 700 * 700 + 16 * 1280 == 1_000_480  # terrain_scene defaults (configs[1])
    assert 2 * 64 * 64 + 50_000 * 50 * 2 + 128 == 5_008_320  # foliage_scene defaults (configs[2])
