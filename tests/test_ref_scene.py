"""SURVEY.md §8 row a17 PINNED: the host table build against the reference's own code.

oracle/_ref/ref_scene_tool is the reference's src/engine/resource/scene.cpp (node hierarchy, transforms, light gathering,
Scene::update, Scene::create_gpu_resources :915-1311) and material.cpp compiled where they lie (one lexical rule, two sites:
oracle/ref_scene/gen.py) against a stand-in device layer (oracle/ref_scene/stub/gfx/vk.h) and driven through the engine's
public API.  What it writes into its Material / Light / Instance storage buffers, the per-node (primitive offset, material)
buffers and the texture descriptor array must equal, BYTE FOR BYTE, what the C++ host layer (helios_b200/shim/src/scene.cpp)
installs through hl_scene_set_tables — on all five BASELINE scene types, a material-override scene and the texture de-dup quirk
(a second material using an already-claimed texture keeps -1, scene.cpp:958-1066).  The Python host (helios_b200/scenes.py)
is held to the same tables (integers exact, matrices and light directions to float rounding: it computes in float64).

The tool only exists where /root/reference is mounted (or as a prebuilt file on the GPU box); its outputs are committed under
tests/golden/ref_scene/ (tools/make_ref_scene_golden.py), so the shim is checked against the reference's bytes everywhere.
glm_pin: every GLM function on the path's host side, stand-in (helios_b200/shim/include/glm.hpp) vs the GLM the reference
vendors, bit for bit on seeded random inputs (golden: tests/golden/glm_pin.bin)."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from helios_b200 import abi, scene_io, scenes
from helios_b200.build import build_library, build_shim
from tests.test_shim_host import read_tables

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden" / "ref_scene"


def shared_texture_scene():
    """two materials that use the SAME albedo texture (+ one with its own): the quirk of scene.cpp:958-1066 — the texture is
    claimed by the first material that uses it, the second one's row keeps -1"""
    s = scenes.terrain_scene(grid=12, n_spheres=2, sphere_level=1, width=32, height=18, textured=True)
    m = s.materials.copy()
    first = [k for k in range(len(m)) if m["texture_indices0"][k][0] >= 0]
    assert len(first) >= 2
    m["texture_indices0"][first[1]][0] = m["texture_indices0"][first[0]][0]  # the file now names the same texture twice
    s.materials = m
    return s


REF_SCENES = {
    "cornell": lambda: scenes.cornell_box(64, 64),
    "soup": lambda: scenes.triangle_soup(200, 64, 36),
    "terrain": lambda: scenes.terrain_scene(grid=40, n_spheres=4, sphere_level=1, width=64, height=36),
    "terrain_textured": lambda: scenes.terrain_scene(grid=20, n_spheres=4, sphere_level=1, width=64, height=36, textured=True),
    "foliage": lambda: scenes.foliage_scene(n_clusters=20, cards_per_cluster=5, width=64, height=36, ground_grid=4, tex_size=16),
    "city": lambda: scenes.city_scene(n_instances=40, n_meshes=5, width=64, height=36, floors=(2, 4), detail=(1, 3)),
    "shared_texture": shared_texture_scene,
}


def read_tables_and_textures(path):
    mats, insts, lights, infos, pc, sky = read_tables(path)
    b = open(path, "rb").read()
    pos = 4 + mats.nbytes + 4 + insts.nbytes + 4 + lights.nbytes + 4 + sum(4 + i.nbytes for i in infos) + 192 + 160
    nt = int(np.frombuffer(b, "<u4", 1, pos)[0])
    tex = np.frombuffer(b, "<u4", nt, pos + 4)
    return mats, insts, lights, infos, tex


@pytest.fixture(scope="module")
def headless():
    build_library()
    return str(build_shim())


@pytest.fixture(scope="module")
def ref_tools():
    from oracle import oracle

    return oracle.build_ref_scene()


def shim_tables(headless, s, tmp_path):
    f, t = tmp_path / "s.hlsc", tmp_path / "shim.tab"
    scene_io.export_scene(s, f)
    r = subprocess.run([headless, "--scene", str(f), "--no-device", "--dump-tables", str(t)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return read_tables_and_textures(t), f


def assert_same_tables(a, b, what):
    for name, x, y in zip(("materials", "instances", "lights"), a[:3], b[:3]):
        assert len(x) == len(y), f"{what}: {name} count {len(x)} vs {len(y)}"
        assert x.tobytes() == y.tobytes(), f"{what}: {name} table differs"
    assert len(a[3]) == len(b[3]) and all(np.array_equal(x, y) for x, y in zip(a[3], b[3])), f"{what}: submesh pairs differ"
    assert np.array_equal(a[4], b[4]), f"{what}: texture array order differs"


@pytest.mark.parametrize("name", sorted(REF_SCENES))
def test_shim_tables_equal_the_reference_bytes(name, headless, ref_tools, tmp_path):
    s = REF_SCENES[name]()
    shim, f = shim_tables(headless, s, tmp_path)
    gold = read_tables_and_textures(GOLD / f"{name}.tab")  # written by the reference's code (tools/make_ref_scene_golden.py)
    assert_same_tables(shim, gold, "shim vs committed reference output")
    if ref_tools[0] is not None:  # the reference's code itself, here and now
        t = tmp_path / "ref.tab"
        r = subprocess.run([str(ref_tools[0]), str(f), str(t)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert_same_tables(shim, read_tables_and_textures(t), "shim vs reference scene.cpp")
        assert open(t, "rb").read() == open(GOLD / f"{name}.tab", "rb").read(), "committed golden is stale"


def assert_same_material(a, c, what):
    """texture slots and channels exactly; colours to float rounding (the scene file carries the sRGB albedo, which the
    reference converts with powf(x, 2.2), scene.cpp:1015-1017 — the Python host did the same in numpy)"""
    assert np.array_equal(a["texture_indices0"], c["texture_indices0"]) and np.array_equal(a["texture_indices1"], c["texture_indices1"]), what
    for f in ("albedo", "emissive", "roughness_metallic"):
        np.testing.assert_allclose(a[f], c[f], rtol=2e-6, atol=1e-7, err_msg=str(what))


@pytest.mark.parametrize("name", sorted(set(REF_SCENES) - {"shared_texture"}))
def test_python_host_tables_equal_the_reference(name):
    """helios_b200/scenes.py (what the parity tests and the bench upload) against the reference's tables: light rows and
    submesh pairs exactly, material rows exactly after the first-use relabelling, matrices to float rounding"""
    s = REF_SCENES[name]()
    mats, insts, lights, infos, tex = read_tables_and_textures(GOLD / f"{name}.tab")
    assert len(insts) == len(s.instances) and len(infos) == len(s.instances)
    assert len(lights) == len(s.lights)
    area = lights["light_data0"][:, 0] == abi.LIGHT_AREA
    # punctual / environment rows: type and order exactly; directions to float rounding (the scene file carries the node's
    # quaternion, and the reference derives forward() = q * (0, 0, 1) from it in fp32, scene.cpp:206-209)
    for f in ("light_data0", "light_data1", "light_data2", "light_data3"):
        np.testing.assert_allclose(lights[f][~area], s.lights[f][~area], rtol=0, atol=2e-6 * max(1.0, float(np.abs(s.lights[f][~area]).max()) if (~area).any() else 1.0))
    assert np.array_equal(lights["light_data0"][:, 0], s.lights["light_data0"][:, 0])
    for i in range(len(insts)):
        np.testing.assert_allclose(insts["model_matrix"][i], s.instances["model_matrix"][i], rtol=0, atol=1e-6 * max(1.0, float(np.abs(s.instances["model_matrix"][i]).max())))
        np.testing.assert_allclose(insts["normal_matrix"][i], s.instances["normal_matrix"][i], atol=1e-6)
        assert np.array_equal(infos[i][:, 0], s.submesh_info[i][:, 0])
        for k in range(len(infos[i])):
            a, c = mats[infos[i][k, 1]], s.materials[s.submesh_info[i][k, 1]]
            assert_same_material(a, c, (name, i, k))
    for k in np.flatnonzero(area):  # area rows: instance, first primitive and count exactly; the material index through the relabelling
        a, c = lights[k], s.lights[k]
        assert a["light_data0"][1] == c["light_data0"][1] and a["light_data0"][3] == c["light_data0"][3] and np.array_equal(a["light_data1"], c["light_data1"])
        assert_same_material(mats[int(a["light_data0"][2])], s.materials[int(c["light_data0"][2])], (name, "area light", int(k)))


def test_shared_texture_quirk_is_in_the_reference_tables():
    mats, _, _, _, tex = read_tables_and_textures(GOLD / "shared_texture.tab")
    albedo = mats["texture_indices0"][:, 0]
    assert (albedo >= 0).sum() >= 1 and len(tex) == len(set(tex.tolist()))  # every texture sits in the array once
    s = shared_texture_scene()
    named = [int(m["texture_indices0"][0]) for m in s.materials if int(m["texture_indices0"][0]) >= 0]
    assert len(named) == len(set(named)) + 1  # the file names one texture twice ...
    assert (albedo >= 0).sum() == len(set(named))  # ... and the second user's row keeps -1 (scene.cpp:996-1012)


def test_glm_stand_in_is_bit_identical_to_the_vendored_glm(ref_tools, tmp_path):
    exe = tmp_path / "glm_pin_shim"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-w", "-ffp-contract=off", f"-I{ROOT / 'helios_b200' / 'shim' / 'include'}", "-o", str(exe), str(ROOT / "oracle" / "ref_scene" / "glm_pin.cpp")])
    mine = subprocess.run([str(exe), "120"], capture_output=True).stdout
    gold = (ROOT / "tests" / "golden" / "glm_pin.bin").read_bytes()
    assert len(mine) == len(gold) > 50_000 and mine == gold
    if ref_tools[1] is not None:
        assert subprocess.run([str(ref_tools[1]), "120"], capture_output=True).stdout == gold, "committed golden is stale"
        assert subprocess.run([str(ref_tools[1]), "1500"], capture_output=True).stdout == subprocess.run([str(exe), "1500"], capture_output=True).stdout
