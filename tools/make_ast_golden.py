"""Writes tests/golden/ast/: AssetCore fixture files and the REFERENCE loader's canonical dumps of them.
  * mesh / material / scene files are written by the reference's own exporters (oracle/_ref/ref_ast_tool
    write-fixtures, built from /root/reference/external/AssetCore/src/exporter/*);
  * image files (which the reference writes with nvtt / cmft, not buildable here) are written by
    helios_b200/ast_io.py — and read back by the reference's loader like everything else;
  * <file>.dump = `ref_ast_tool dump <kind> <file>`: what external/AssetCore/src/loader/loader.cpp loads.
Run where /root/reference is mounted:  python tools/make_ast_golden.py"""
import shutil
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from helios_b200 import ast_io  # noqa: E402
from oracle import oracle as O  # noqa: E402


def bc_blocks(rng, n_blocks, block_bytes):
    return rng.integers(0, 256, n_blocks * block_bytes, dtype=np.uint8).tobytes()


def write_images(d: Path):
    rng = np.random.default_rng(11)
    (d / "texture").mkdir(parents=True, exist_ok=True)
    (d / "source").mkdir(exist_ok=True)
    chk = np.zeros((8, 8, 4), np.uint8)
    chk[..., :3] = np.where(((np.arange(8)[:, None] // 2 + np.arange(8)[None, :] // 2) % 2)[..., None] == 0, 220, 40)
    chk[..., 3] = 255
    mip1 = chk.reshape(4, 2, 4, 2, 4).mean((1, 3)).astype(np.uint8)
    ast_io.write_image(d / "texture" / "checker.ast", "checker", [[(8, 8, chk.tobytes()), (4, 4, mip1.tobytes())]], 4)
    ast_io.write_image(d / "texture" / "gray2.ast", "gray2", [[(5, 3, rng.integers(0, 256, 5 * 3 * 2, dtype=np.uint8).tobytes())]], 2)
    ast_io.write_image(d / "texture" / "rgb_f16.ast", "rgb_f16", [[(4, 4, rng.random((4, 4, 3)).astype(np.float16).tobytes())]], 3, ast_io.PIXEL_FLOAT16)
    env = rng.random((6, 4, 4, 4)).astype(np.float32)
    ast_io.write_image(d / "texture" / "env.ast", "env", [[(4, 4, env[f].tobytes()), (2, 2, env[f, ::2, ::2].tobytes())] for f in range(6)], 4, ast_io.PIXEL_FLOAT32)
    for name, comp, bb, comps in (("bc1", 1, 8, 3), ("bc1a", 2, 8, 4), ("bc2", 3, 16, 4), ("bc3", 4, 16, 4), ("bc4", 6, 8, 1), ("bc5", 7, 16, 2)):
        ast_io.write_image(d / "texture" / f"{name}.ast", name, [[(12, 8, bc_blocks(rng, 3 * 2, bb)), (6, 4, bc_blocks(rng, 2 * 1, bb))]], comps, ast_io.PIXEL_UNORM8, comp)


def main():
    tool = O.build_ref_ast()
    if tool is None:
        raise SystemExit("oracle/_ref/ref_ast_tool is not available (no /root/reference)")
    d = ROOT / "tests" / "golden" / "ast"
    if d.exists():
        shutil.rmtree(d)
    d.mkdir(parents=True)
    write_images(d)
    (d / "scene").mkdir()
    subprocess.check_call([str(tool), "write-fixtures", str(d)])
    # two hand-written documents for the loader's defaults and rejections (a PROPERTY_ROUGHNESS without a value is left
    # out: the reference keeps the property with an uninitialised float, loader.cpp:339-345)
    (d / "material" / "sparse.json").write_text('{"textures": [{"type": "TEXTURE_NORMAL"}, {"path": "../texture/gray2.ast", "srgb": false, "type": "TEXTURE_METALLIC", "channel_index": 1}],'
                                                 ' "properties": [{"type": "PROPERTY_ALBEDO", "value": [1, 2, 3]}, {"type": "PROPERTY_METALLIC"}, {"type": "PROPERTY_ROUGHNESS", "value": 0.75},'
                                                 ' {"type": "PROPERTY_EMISSIVE", "value": [0.5, 0.25, 0.125, 1.0]}, {"type": "PROPERTY_UNKNOWN", "value": 3}, {"value": 1}]}')
    (d / "scene" / "minimal.json").write_text('{"scene_graph": {"type": "SCENE_NODE_ROOT", "name": "r\\u00e9 \\"q\\"", "position": [1, 2, 3], "rotation": [0, 0, 0], "scale": [1, 1, 1], "children": ['
                                               '{"type": "SCENE_NODE_IBL", "name": "ibl"}, {"type": "SCENE_NODE_CAMERA", "name": "c", "position": [0, 0, 0, 1], "rotation": [1, 2, 3], "scale": [1, 1, 1], "near_plane": 0.5, "far_plane": 10, "fov": 45}]}}')
    shutil.rmtree(d / "source")
    kinds = {"texture": "image", "mesh": "mesh", "material": "material", "scene": "scene"}
    for sub, kind in kinds.items():
        for f in sorted((d / sub).iterdir()):
            if f.suffix == ".dump":
                continue
            out = subprocess.run([str(tool), "dump", kind, str(f), str(d) + "/"], capture_output=True, text=True, check=True).stdout
            Path(str(f) + ".dump").write_text(out)
            print(f.relative_to(d), "->", len(out.splitlines()), "lines")


if __name__ == "__main__":
    main()
