"""C++ host layer (helios_b200/shim: the reference's Scene / Mesh / Material / Node / PathIntegrator surface over the
C ABI), host-only part: the scene graph rebuilt through the engine API must produce the same Material /
Instance / Light tables, per-instance (primitive offset, material) pairs, push constants and sky coefficients
as the Python restatement of src/engine/resource/scene.cpp:915-1311 that the parity tests drive.  No GPU:
helios_headless --no-device builds the tables and stops (rendering without a device throws)."""
import struct
import subprocess

import numpy as np
import pytest

from helios_b200 import abi, scene_io, scenes
from helios_b200.build import build_library, build_shim
from helios_b200.sky import sky_coefficients

SCENES = {
    "cornell": lambda: scenes.cornell_box(64, 64),
    "terrain": lambda: scenes.terrain_scene(grid=40, n_spheres=4, sphere_level=1, width=64, height=36),
    "foliage": lambda: scenes.foliage_scene(n_clusters=20, cards_per_cluster=5, width=64, height=36, ground_grid=4, tex_size=16),
    "city": lambda: scenes.city_scene(n_instances=12, n_meshes=3, width=64, height=36, floors=(2, 4), detail=(1, 3)),
}


@pytest.fixture(scope="module")
def headless():
    build_library()
    return str(build_shim())


def read_tables(path):
    b = open(path, "rb").read()
    pos = 0

    def vec(dt, per=1):
        nonlocal pos
        n = struct.unpack_from("<I", b, pos)[0]
        pos += 4
        a = np.frombuffer(b, dt, n * per, pos)
        pos += n * per * np.dtype(dt).itemsize
        return a

    mats, insts, lights = vec(abi.MATERIAL), vec(abi.INSTANCE), vec(abi.LIGHT)
    ni = struct.unpack_from("<I", b, pos)[0]
    pos += 4
    infos = [vec("<u4", 2).reshape(-1, 2) for _ in range(ni)]
    pc = np.frombuffer(b, abi.PUSH_CONSTANTS, 1, pos)[0]
    pos += 192
    sky = np.frombuffer(b, "<f4", 40, pos)
    return mats, insts, lights, infos, pc, sky


@pytest.mark.parametrize("name", sorted(SCENES))
def test_engine_api_builds_the_same_tables(name, headless, tmp_path):
    s = SCENES[name]()
    f, t = tmp_path / "s.hlsc", tmp_path / "s.tab"
    scene_io.export_scene(s, f)
    r = subprocess.run([headless, "--scene", str(f), "--no-device", "--dump-tables", str(t)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    mats, insts, lights, infos, pc, sky = read_tables(t)
    assert len(insts) == len(s.instances) and len(lights) == len(s.lights) and len(infos) == len(s.instances)
    assert len(mats) <= len(s.materials)  # the engine only emits materials that a submesh uses (first-use order)
    mesh_map = {}  # the engine numbers meshes by first use (scene.cpp:946-950): a relabelling of the Python order
    for i, inst in enumerate(insts):
        assert mesh_map.setdefault(int(inst["mesh_index"]), int(s.instances[i]["mesh_index"])) == int(s.instances[i]["mesh_index"])
        np.testing.assert_allclose(inst["model_matrix"], s.instances[i]["model_matrix"], atol=1e-6)
        np.testing.assert_allclose(inst["normal_matrix"], s.instances[i]["normal_matrix"], atol=1e-6)
        assert np.array_equal(infos[i][:, 0], s.submesh_info[i][:, 0])
        for k in range(len(infos[i])):
            a, c = mats[infos[i][k, 1]], s.materials[s.submesh_info[i][k, 1]]
            assert np.array_equal(a["texture_indices0"], c["texture_indices0"]) and np.array_equal(a["texture_indices1"], c["texture_indices1"])
            for fld in ("albedo", "emissive", "roughness_metallic"):
                np.testing.assert_allclose(a[fld], c[fld], atol=1e-6)
    for a, c in zip(lights, s.lights):  # same order: area..., environment, directional..., point..., spot...
        kind = int(a["light_data0"][0])
        assert kind == int(c["light_data0"][0])
        if kind == abi.LIGHT_AREA:  # (type, mesh node, material row, first triangle), (triangle count)
            assert a["light_data0"][1] == c["light_data0"][1] and a["light_data0"][3] == c["light_data0"][3] and a["light_data1"][0] == c["light_data1"][0]
            ma, mc = mats[int(a["light_data0"][2])], s.materials[int(c["light_data0"][2])]
            np.testing.assert_allclose(ma["emissive"], mc["emissive"], atol=1e-6)
        else:
            for fld in ("light_data0", "light_data1", "light_data2", "light_data3"):
                np.testing.assert_allclose(a[fld], c[fld], atol=2e-6)
    assert len(set(mesh_map.values())) == len(mesh_map)
    ref = s.push_constants(0)
    for fld in abi.PUSH_CONSTANTS.names:
        np.testing.assert_allclose(np.asarray(pc[fld], np.float64), np.asarray(ref[fld], np.float64), rtol=2e-5, atol=2e-5, err_msg=fld)
    if s.sun_direction is not None:
        np.testing.assert_allclose(sky, sky_coefficients(s.sun_direction), rtol=1e-5, atol=1e-6)


def test_rendering_without_a_device_fails_loudly(headless, tmp_path):
    """there is no CPU path: asking the headless driver to render on a machine without CUDA is an error"""
    import ctypes

    try:
        ctypes.CDLL("libcuda.so.1")
        pytest.skip("a CUDA driver is present")
    except OSError:
        pass
    s = SCENES["cornell"]()
    f = tmp_path / "s.hlsc"
    scene_io.export_scene(s, f)
    r = subprocess.run([headless, "--scene", str(f), "--spp", "1"], capture_output=True, text=True)
    assert r.returncode != 0 and "CUDA" in r.stderr
