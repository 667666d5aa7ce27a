"""SURVEY.md 8 row f1 — the AssetCore asset route (image / mesh .ast, material / scene JSON -> ResourceManager).
The loader in helios_b200/shim is checked against the REFERENCE'S OWN loader (external/AssetCore/src/loader/loader.cpp,
compiled as oracle/_ref/ref_ast_tool where /root/reference is mounted) on fixture files that the reference's own
exporters wrote (tests/golden/ast/, tools/make_ast_golden.py); the reference loader's dumps are committed next to
the fixtures, so the comparison also runs where neither the checkout nor the prebuilt tool exists.
Host only: no GPU needed."""
import math
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

from helios_b200 import abi, ast_io, scene_io, scenes
from helios_b200.build import ROOT as PKG, build_library, build_shim

REPO = Path(__file__).resolve().parent.parent
GOLD = REPO / "tests" / "golden" / "ast"
KINDS = {"texture": "image", "mesh": "mesh", "material": "material", "scene": "scene"}
FIXTURES = sorted(str(f.relative_to(GOLD)) for sub in KINDS for f in (GOLD / sub).iterdir() if f.suffix != ".dump")


@pytest.fixture(scope="module")
def my_dump():
    build_library()
    build_shim()
    src, exe = REPO / "tests" / "ast" / "my_ast_dump.cpp", REPO / "tests" / "ast" / "_build" / "my_ast_dump"
    deps = [src, REPO / "oracle" / "ref_ast" / "dump_format.h", PKG / "libhelios_engine.so"]
    if not exe.exists() or any(d.stat().st_mtime > exe.stat().st_mtime for d in deps):
        exe.parent.mkdir(exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", f"-I{PKG / 'shim' / 'include'}", f"-I{REPO / 'include'}", str(src), "-o", str(exe), f"-L{PKG}", "-lhelios_engine", "-lhelios_b200", f"-Wl,-rpath,{PKG}"])
    return str(exe)


def run(*cmd):
    return subprocess.run([str(c) for c in cmd], capture_output=True, text=True, check=True).stdout


@pytest.mark.parametrize("rel", FIXTURES)
def test_loader_matches_reference_loader_golden(rel, my_dump):
    """byte-identical canonical dump: every header field, every float bit pattern, payload hashes, resolved paths"""
    f = GOLD / rel
    want = Path(str(f) + ".dump").read_text()
    got = run(my_dump, KINDS[rel.split("/")[0]], f, str(GOLD) + "/")
    assert got == want
    assert want.strip() != ""


def test_fixture_set_exercises_the_quirks():
    d = (GOLD / "material" / "fixture_glow.json.dump").read_text()
    assert "properties=3" in d  # the reference's exporter writes emissive with 3 values; its loader drops them (4 required)
    s = (GOLD / "material" / "sparse.json.dump").read_text()
    assert 'name="untitled"' in s and "srgb=1" in s  # defaults: name, srgb = true


@pytest.mark.parametrize("rel", FIXTURES)
def test_loader_matches_reference_loader_live(rel, my_dump, tmp_path):
    """same comparison against the reference loader run here (skipped where oracle/_ref/ref_ast_tool cannot exist)"""
    from oracle import oracle as O

    tool = O.build_ref_ast()
    if tool is None:
        pytest.skip("no /root/reference and no prebuilt oracle/_ref/ref_ast_tool")
    kind = KINDS[rel.split("/")[0]]
    assert run(my_dump, kind, GOLD / rel, str(GOLD) + "/") == run(tool, "dump", kind, GOLD / rel, str(GOLD) + "/")


def test_missing_and_corrupt_files_fail_like_the_reference(my_dump, tmp_path):
    for kind in KINDS.values():
        assert run(my_dump, kind, tmp_path / "nope").strip() == "load failed"
    # a truncated image: the reference's stream reads fail silently and it returns stale / uninitialised fields;
    # here the load fails
    (tmp_path / "cut.ast").write_bytes((GOLD / "texture" / "checker.ast").read_bytes()[:100])
    assert run(my_dump, "image", tmp_path / "cut.ast").strip() == "load failed"
    (tmp_path / "bad.json").write_text('{"name": ')
    assert run(my_dump, "material", tmp_path / "bad.json").strip() == "load failed"
    # mesh whose material file is missing: the whole mesh fails (loader.cpp:155-159)
    m = tmp_path / "mesh"
    m.mkdir()
    ast_io.write_mesh(m / "m.ast", "m", np.zeros(3, ast_io.AST_VERTEX), [0, 1, 2], np.zeros(1, ast_io.AST_SUBMESH), ["../material/none.json"])
    assert "load failed" in run(my_dump, "mesh", m / "m.ast")
    # counts larger than the file: rejected, not over-read
    raw = bytearray((GOLD / "mesh" / "fixture_mesh.ast").read_bytes())
    struct.pack_into("<I", raw, 8 + 8, 1 << 30)  # vertex_count
    (tmp_path / "huge.ast").write_bytes(raw)
    assert run(my_dump, "mesh", tmp_path / "huge.ast").strip() == "load failed"


# ---- texel conversion (core/resource_manager.cpp:14-70 format tables) and block decompression ---------------------
def _rgb565(c):
    r, g, b = (c >> 11) & 31, (c >> 5) & 63, c & 31
    return np.array([(r << 3) | (r >> 2), (g << 2) | (g >> 4), (b << 3) | (b >> 2)], np.int64)


def _color_block(b, punch):
    c0, c1 = b[0] | (b[1] << 8), b[2] | (b[3] << 8)
    p = np.zeros((4, 4), np.int64)
    p[0, :3], p[1, :3], p[:, 3] = _rgb565(c0), _rgb565(c1), 255
    if c0 > c1 or not punch:
        p[2, :3], p[3, :3] = (2 * p[0, :3] + p[1, :3]) // 3, (p[0, :3] + 2 * p[1, :3]) // 3
    else:
        p[2, :3], p[3] = (p[0, :3] + p[1, :3]) // 2, 0
    idx = int.from_bytes(bytes(b[4:8]), "little")
    return np.stack([p[(idx >> (2 * i)) & 3] for i in range(16)])


def _alpha_block(b):
    a = [int(b[0]), int(b[1])]
    if a[0] > a[1]:
        a += [((7 - i) * a[0] + i * a[1]) // 7 for i in range(1, 7)]
    else:
        a += [((5 - i) * a[0] + i * a[1]) // 5 for i in range(1, 5)] + [0, 255]
    bits = int.from_bytes(bytes(b[2:8]), "little")
    return np.array([a[(bits >> (3 * i)) & 7] for i in range(16)])


def _decode_bc(comp, data, w, h):
    bb = 8 if comp in (1, 2, 6) else 16
    bw, bh = (w + 3) // 4, (h + 3) // 4
    out = np.zeros((h, w, 4), np.uint8)
    for by in range(bh):
        for bx in range(bw):
            b = [int(x) for x in data[(by * bw + bx) * bb : (by * bw + bx + 1) * bb]]
            if comp in (1, 2):
                px = _color_block(b, True)
            elif comp == 3:
                px = _color_block(b[8:], False)
                px[:, 3] = [((b[i // 2] >> (4 * (i & 1))) & 15) * 17 for i in range(16)]
            elif comp in (4, 5):
                px = _color_block(b[8:], False)
                px[:, 3] = _alpha_block(b)
            else:
                px = np.zeros((16, 4), np.int64)
                px[:, 0], px[:, 3] = _alpha_block(b), 255
                if comp == 7:
                    px[:, 1] = _alpha_block(b[8:])
            for i in range(16):
                x, y = bx * 4 + (i & 3), by * 4 + (i >> 2)
                if x < w and y < h:
                    out[y, x] = px[i]
    return out


def _level0(path):
    raw = path.read_bytes()
    (nlen,) = struct.unpack_from("<H", raw, 8)
    pos = 10 + nlen + 8
    w, h, size = struct.unpack_from("<HHi", raw, pos)
    return w, h, np.frombuffer(raw, np.uint8, size, pos + 8)


@pytest.mark.parametrize("name,comp", [("bc1", 1), ("bc1a", 2), ("bc2", 3), ("bc3", 4), ("bc4", 6), ("bc5", 7)])
def test_block_compressed_textures_decode(name, comp, my_dump, tmp_path):
    w, h, data = _level0(GOLD / "texture" / f"{name}.ast")
    want = _decode_bc(comp, data, w, h)
    for srgb in (0, 1):
        head = run(my_dump, "texels", GOLD / "texture" / f"{name}.ast", srgb, 0, tmp_path / "t.bin").split()
        if srgb and comp in (6, 7):
            assert head == ["no", "format"]  # BC4 / BC5 have no sRGB VkFormat in the reference's table
            continue
        assert head == [str(abi.TEX_RGBA8_SRGB if srgb else abi.TEX_RGBA8_UNORM), str(w), str(h)]
        got = np.fromfile(tmp_path / "t.bin", np.uint8).reshape(h, w, 4)
        assert np.array_equal(got, want)


def test_uncompressed_format_table(my_dump, tmp_path):
    t = tmp_path / "t.bin"
    # 4 x 8-bit: sRGB flag -> R8G8B8A8_SRGB, otherwise the reference's SNORM quirk
    assert run(my_dump, "texels", GOLD / "texture" / "checker.ast", 1, 0, t).split() == [str(abi.TEX_RGBA8_SRGB), "8", "8"]
    assert run(my_dump, "texels", GOLD / "texture" / "checker.ast", 0, 0, t).split() == [str(abi.TEX_RGBA8_SNORM), "8", "8"]
    # 2 x 8-bit: no sRGB format; SNORM with the missing channels read as 0, 0, 1
    assert run(my_dump, "texels", GOLD / "texture" / "gray2.ast", 1, 0, t).split() == ["no", "format"]
    assert run(my_dump, "texels", GOLD / "texture" / "gray2.ast", 0, 0, t).split() == [str(abi.TEX_RGBA8_SNORM), "5", "3"]
    w, h, data = _level0(GOLD / "texture" / "gray2.ast")
    got = np.fromfile(t, np.uint8).reshape(3, 5, 4)
    assert np.array_equal(got[..., :2], data.reshape(3, 5, 2)) and (got[..., 2] == 0).all() and (got[..., 3] == 127).all()
    # 3 x fp16 -> RGBA32F, alpha 1
    assert run(my_dump, "texels", GOLD / "texture" / "rgb_f16.ast", 0, 0, t).split() == [str(abi.TEX_RGBA32F), "4", "4"]
    w, h, data = _level0(GOLD / "texture" / "rgb_f16.ast")
    got = np.fromfile(t, np.float32).reshape(4, 4, 4)
    assert np.array_equal(got[..., :3], data.view(np.float16).reshape(4, 4, 3).astype(np.float32)) and (got[..., 3] == 1).all()
    # cube map face 5
    assert run(my_dump, "texels", GOLD / "texture" / "env.ast", 0, 5, t).split() == [str(abi.TEX_RGBA32F), "4", "4"]


def test_local_matrix_is_imguizmo_recompose(my_dump):
    """populate_transform_node: ImGuizmo::RecomposeMatrixFromComponents (core/resource_manager.cpp:666-672)"""
    rng = np.random.default_rng(5)
    cases = [([0, 0, 0], [0, 0, 0], [1, 1, 1]), ([1, 2, 3], [90, 0, 0], [1, 1, 1]), ([0, 0, 0], [10, 20, 30], [1, 2, 0.5]), ([0, 1, 0], [0, -45, 0], [1, 1, 0])]
    cases += [(rng.uniform(-5, 5, 3), rng.uniform(-180, 180, 3), rng.uniform(0.1, 3, 3)) for _ in range(20)]
    for pos, rot, scl in cases:
        args = [repr(float(np.float32(v))) for v in (*pos, *rot, *scl)]
        got = np.array([int(x, 16) for x in run(my_dump, "matrix", *args).split()], np.uint32).view(np.float32)
        want = ast_io.recompose([np.float32(v) for v in pos], [np.float32(v) for v in rot], [np.float32(v) for v in scl])
        assert np.allclose(got, want, rtol=0, atol=2e-6 * max(1.0, float(np.abs(want).max())))
    ident = np.array([int(x, 16) for x in run(my_dump, "matrix", *["0"] * 6, "1", "1", "1").split()], np.uint32).view(np.float32)
    assert np.array_equal(ident, np.eye(4, dtype=np.float32).reshape(16))


# ---- the whole route: ResourceManager::load_scene builds the same GPU tables as the direct engine-API route --------
SCENES = {
    "cornell": lambda: scenes.cornell_box(64, 64),
    "terrain_textured": lambda: scenes.terrain_scene(grid=24, n_spheres=3, sphere_level=1, width=64, height=36, textured=True) if "textured" in scenes.terrain_scene.__code__.co_varnames else scenes.terrain_scene(grid=24, n_spheres=3, sphere_level=1, width=64, height=36),
    "foliage": lambda: scenes.foliage_scene(n_clusters=20, cards_per_cluster=5, width=64, height=36, ground_grid=4, tex_size=16),
    "city": lambda: scenes.city_scene(n_instances=12, n_meshes=3, width=64, height=36, floors=(2, 4), detail=(1, 3)),
}


@pytest.mark.parametrize("name", sorted(SCENES))
def test_resource_manager_route_builds_the_same_tables(name, tmp_path):
    from tests.test_shim_host import read_tables

    build_library()
    exe = str(build_shim())
    s = SCENES[name]()
    scene_io.export_scene(s, tmp_path / "s.hlsc")
    rel = ast_io.export_assets(s, tmp_path, name)
    run(exe, "--scene", tmp_path / "s.hlsc", "--no-device", "--dump-tables", tmp_path / "a.tab")
    run(exe, "--ast-scene", rel, "--asset-root", tmp_path, "--width", s.width, "--height", s.height, "--focal-length", s.camera.focal_length, "--aperture", s.camera.aperture_radius, "--no-device",
        "--dump-tables", tmp_path / "b.tab")
    A, B = read_tables(tmp_path / "a.tab"), read_tables(tmp_path / "b.tab")
    assert A[0].tobytes() == B[0].tobytes()  # materials: bit-identical (JSON numbers round-trip float32 exactly)
    assert len(A[1]) == len(B[1]) and len(A[2]) == len(B[2])
    for f in A[1].dtype.names:  # instances: matrices go through Euler angles (degrees) and back
        assert np.allclose(A[1][f], B[1][f], rtol=0, atol=1e-5 * max(1.0, float(np.abs(np.asarray(A[1][f], np.float64)).max())))
    spot = np.array([int(l["light_data0"][0]) == abi.LIGHT_SPOT for l in A[2]])
    for f in A[2].dtype.names:
        a, b = np.asarray(A[2][f], np.float64), np.asarray(B[2][f], np.float64)
        if f == "light_data3" and spot.any():
            # ResourceManager::create_spot_light_node passes the INNER cone angle as the outer one too (:591)
            assert np.allclose(b[spot, 1], a[spot, 0], atol=1e-6)
            a, b = a.copy(), b.copy()
            a[spot, 1] = b[spot, 1] = 0
        assert np.allclose(a, b, rtol=0, atol=1e-5 * max(1.0, float(np.abs(a).max())))
    assert all(np.array_equal(x, y) for x, y in zip(A[3], B[3]))
    for f in A[4].dtype.names:
        assert np.allclose(A[4][f], B[4][f], rtol=0, atol=1e-5 * max(1.0, float(np.abs(np.asarray(A[4][f], np.float64)).max())))
    assert np.allclose(A[5], B[5], atol=1e-5)
    if name == "cornell":  # identity transforms survive exactly: the whole table set is bit-identical
        assert all(A[k].tobytes() == B[k].tobytes() for k in (1, 2, 4))
