"""Writer for the reference's own on-disk asset formats (AssetCore: binary image / mesh `.ast`, material JSON,
scene JSON — SURVEY.md Appendix D; layouts from external/AssetCore/include/common/{header,image,mesh}.h, keys from
external/AssetCore/src/exporter/{material,scene}_exporter.cpp and what src/loader/loader.cpp reads).

`export_assets(scene, root)` lays a procedural SceneData out the way the reference's release zip is laid out,

    <root>/assets/scene/<name>.json      <root>/assets/mesh/mesh<k>.ast
    <root>/assets/material/mat<k>.json   <root>/assets/texture/tex<k>.ast   <root>/assets/texture/env.ast

so that `helios_headless --ast-scene scene/<name>.json --asset-root <root>` (ResourceManager::load_scene, the
reference's asset route) renders the same scene as the direct route.  Test and bench tooling: the product reads
these files, it never writes them."""
from __future__ import annotations

import json
import math
import struct
from pathlib import Path

import numpy as np

from . import abi

AST_VERTEX = np.dtype([("position", "<f4", 3), ("tex_coord", "<f4", 2), ("normal", "<f4", 3), ("tangent", "<f4", 3), ("bitangent", "<f4", 3)])
AST_SUBMESH = np.dtype(
    [("material_index", "<u4"), ("index_count", "<u4"), ("vertex_count", "<u4"), ("base_vertex", "<u4"), ("base_index", "<u4"), ("max_extents", "<f4", 3), ("min_extents", "<f4", 3), ("name", "S150"), ("pad", "V2")]
)
assert AST_VERTEX.itemsize == 56 and AST_SUBMESH.itemsize == 196
COMPRESSION_NONE, PIXEL_UNORM8, PIXEL_FLOAT16, PIXEL_FLOAT32 = 0, 1, 2, 4
TYPE_IMAGE, TYPE_MESH = 0, 1


def _file_header(asset_type: int) -> bytes:
    # BINFileHeader: uint32 magic ('a','s','t', unspecified), uint8 version = 1, uint8 type, 2 B padding
    return b"ast\0" + struct.pack("<BBH", 1, asset_type, 0)


def write_image(path, name: str, levels, components: int, pixel_type: int = PIXEL_UNORM8, compression: int = COMPRESSION_NONE):
    """levels[array_slice][mip] = (width, height, bytes)"""
    out = bytearray(_file_header(TYPE_IMAGE))
    nb = name.encode()
    out += struct.pack("<H", len(nb)) + nb
    # BINImageHeader: u8 compression, u8 channel_size, u8 num_channels, (pad), u16 num_array_slices, u8 num_mip_slices, (pad)
    out += struct.pack("<BBBxHBx", compression, pixel_type, components, len(levels), len(levels[0]))
    for mips in levels:
        for w, h, data in mips:
            data = bytes(data)
            out += struct.pack("<HHi", w, h, len(data)) + data
    Path(path).write_bytes(out)


def write_mesh(path, name: str, vertices, indices, submeshes, material_paths):
    """vertices: AST_VERTEX array; submeshes: AST_SUBMESH array; material_paths: JSON paths relative to the mesh file"""
    v = np.ascontiguousarray(vertices, AST_VERTEX)
    idx = np.ascontiguousarray(indices, np.uint32)
    sm = np.ascontiguousarray(submeshes, AST_SUBMESH)
    pos = v["position"] if len(v) else np.zeros((1, 3), np.float32)
    out = bytearray(_file_header(TYPE_MESH))
    # BINMeshFileHeader (196 B)
    out += struct.pack("<IIIII", len(sm), len(material_paths), len(v), 0, len(idx))
    out += pos.max(0).astype("<f4").tobytes() + pos.min(0).astype("<f4").tobytes()
    out += name.encode()[:149].ljust(150, b"\0") + b"\0\0"
    out += v.tobytes() + idx.tobytes() + sm.tobytes()
    for p in material_paths:
        out += p.encode()[:149].ljust(150, b"\0")
    Path(path).write_bytes(out)


def euler_degrees(right, up, forward):
    """Euler angles (degrees) that ImGuizmo::RecomposeMatrixFromComponents turns back into the rotation whose basis
    vectors are right / up / forward (the inverse: ImGuizmo::DecomposeMatrixToComponents, ImGuizmo.cpp:2045-2067)"""
    m = np.stack([np.asarray(right, np.float64), np.asarray(up, np.float64), np.asarray(forward, np.float64)])  # rows
    m = m / np.linalg.norm(m, axis=1, keepdims=True)
    cy = math.hypot(m[1, 2], m[2, 2])
    ry = math.degrees(math.atan2(-m[0, 2], cy))
    if cy < 1e-6:
        # gimbal lock (ry = +-90 degrees): rx and rz turn about the same axis; put everything into rx
        return [math.degrees(math.atan2(-m[2, 1], m[1, 1])), ry, 0.0]
    return [math.degrees(math.atan2(m[1, 2], m[2, 2])), ry, math.degrees(math.atan2(m[0, 1], m[0, 0]))]


def recompose(position, rotation_deg, scale) -> np.ndarray:
    """ImGuizmo::RecomposeMatrixFromComponents in float64 -> 16 floats (column-major glm::mat4)"""
    def axis_rot(i, deg):
        a = math.radians(deg)
        s, c = math.sin(a), math.cos(a)
        m = np.eye(4)
        j, k = (i + 1) % 3, (i + 2) % 3
        m[j, j], m[j, k], m[k, j], m[k, k] = c, s, -s, c
        return m

    m = axis_rot(0, rotation_deg[0]) @ axis_rot(1, rotation_deg[1]) @ axis_rot(2, rotation_deg[2])
    for i in range(3):
        m[i, :] *= 0.001 if abs(scale[i]) < 1.1920929e-07 else scale[i]
    m[3, :3] = position
    return m.reshape(16).astype(np.float32)


def _basis_for_forward(fwd):
    f = np.asarray(fwd, np.float64)
    f = f / np.linalg.norm(f)
    ref = np.array([0.0, 1.0, 0.0]) if abs(f[1]) < 0.99 else np.array([1.0, 0.0, 0.0])
    r = np.cross(ref, f)
    r /= np.linalg.norm(r)
    return r, np.cross(f, r), f


def _f(x):
    """a float32 value as a JSON number that parses back to the same float32"""
    return float(np.float32(x))


def _transform_from_matrix(model16):
    m = np.asarray(model16, np.float64).reshape(4, 4)  # rows = glm columns = basis vectors
    scale = np.linalg.norm(m[:3, :3], axis=1)
    rot = euler_degrees(*(m[:3, :3] / scale[:, None]))
    return {"position": [_f(v) for v in m[3, :3]], "rotation": [_f(v) for v in rot], "scale": [_f(v) for v in scale]}


def export_assets(scene, root, name: str = "scene"):
    """SceneData -> AssetCore files under <root>/assets/.  Returns the scene path relative to assets/."""
    assets = Path(root) / "assets"
    for d in ("scene", "mesh", "material", "texture"):
        (assets / d).mkdir(parents=True, exist_ok=True)
    # textures: 8-bit RGBA; sRGB ones load as R8G8B8A8_SRGB, the others as R8G8B8A8_SNORM (the reference's
    # format table has no 8-bit UNORM entry, core/resource_manager.cpp:32-36) — so UNORM textures cannot be expressed
    srgb_flag = []
    for k, (fmt, w, h, data) in enumerate(scene.textures):
        if fmt == abi.TEX_RGBA32F:
            write_image(assets / "texture" / f"tex{k}.ast", f"tex{k}", [[(w, h, np.ascontiguousarray(data, np.float32).tobytes())]], 4, PIXEL_FLOAT32)
            srgb_flag.append(False)
        elif fmt in (abi.TEX_RGBA8_SRGB, abi.TEX_RGBA8_SNORM):
            write_image(assets / "texture" / f"tex{k}.ast", f"tex{k}", [[(w, h, np.ascontiguousarray(data, np.uint8).tobytes())]], 4, PIXEL_UNORM8)
            srgb_flag.append(fmt == abi.TEX_RGBA8_SRGB)
        else:
            raise ValueError("8-bit UNORM textures have no AssetCore/engine format (they would load as SNORM)")
    alpha = np.zeros(len(scene.materials), bool)
    for m in scene.meshes:
        for i, s in enumerate(m.submeshes):
            if not int(s["opaque"]):
                alpha[m.materials[i]] = True
    for k, m in enumerate(scene.materials):
        a = m["albedo"].astype(np.float64)
        src = [*np.power(np.maximum(a[:3], 0.0), 1.0 / 2.2), a[3]]  # undo scene.cpp's pow(rgb, 2.2) on constant albedo
        t0, t1 = m["texture_indices0"], m["texture_indices1"]
        slots = [("TEXTURE_ALBEDO", int(t0[0]), 0), ("TEXTURE_NORMAL", int(t0[1]), 0), ("TEXTURE_ROUGHNESS", int(t0[2]), int(t1[2])), ("TEXTURE_METALLIC", int(t0[3]), int(t1[3])), ("TEXTURE_EMISSIVE", int(t1[0]), 0)]
        doc = {
            "name": f"mat{k}",
            "double_sided": False,
            "alpha_mask": bool(alpha[k]),
            "material_type": "MATERIAL_OPAQUE",
            "shading_model": "SHADING_MODEL_STANDARD",
            "textures": [{"path": f"../texture/tex{t}.ast", "srgb": bool(srgb_flag[t]), "type": ty, "channel_index": max(ch, 0)} for ty, t, ch in slots if t >= 0],
            "properties": [
                {"type": "PROPERTY_ALBEDO", "value": [_f(v) for v in src]},
                {"type": "PROPERTY_EMISSIVE", "value": [_f(v) for v in m["emissive"]]},
                {"type": "PROPERTY_METALLIC", "value": _f(m["roughness_metallic"][1])},
                {"type": "PROPERTY_ROUGHNESS", "value": _f(m["roughness_metallic"][0])},
            ],
        }
        (assets / "material" / f"mat{k}.json").write_text(json.dumps(doc, indent=1))
    for k, m in enumerate(scene.meshes):
        v = np.zeros(len(m.vertices), AST_VERTEX)
        v["position"], v["tex_coord"] = m.vertices["position"][:, :3], m.vertices["tex_coord"][:, :2]
        v["normal"], v["tangent"], v["bitangent"] = m.vertices["normal"][:, :3], m.vertices["tangent"][:, :3], m.vertices["bitangent"][:, :3]
        sm = np.zeros(len(m.submeshes), AST_SUBMESH)
        for i, s in enumerate(m.submeshes):
            sm[i]["material_index"], sm[i]["index_count"], sm[i]["vertex_count"], sm[i]["base_index"] = i, s["index_count"], s["vertex_count"], s["base_index"]
            sm[i]["name"] = f"submesh{i}".encode()
        write_mesh(assets / "mesh" / f"mesh{k}.ast", f"mesh{k}", v, m.indices, sm, [f"../material/mat{g}.json" for g in m.materials])
    children = []
    for i, inst in enumerate(scene.instances):
        children.append({"type": "SCENE_NODE_MESH", "name": f"mesh_node{i}", "mesh": f"mesh/mesh{int(inst['mesh_index'])}.ast", "material_override": "", "casts_shadow": True, "children": [], **_transform_from_matrix(inst["model_matrix"])})
    c = scene.camera
    children.append({"type": "SCENE_NODE_CAMERA", "name": "camera", "near_plane": _f(c.near), "far_plane": _f(c.far), "fov": _f(c.fov), "children": [], "position": [_f(v) for v in c.position], "rotation": [_f(v) for v in euler_degrees(c.right, c.up, c.forward)], "scale": [1.0, 1.0, 1.0]})
    for i, l in enumerate(scene.lights):
        t = int(l["light_data0"][0])
        base = {"name": f"light{i}", "children": [], "color": [_f(v) for v in l["light_data0"][1:4]], "intensity": _f(l["light_data1"][3]), "radius": _f(l["light_data2"][3]), "casts_shadows": True, "scale": [1.0, 1.0, 1.0]}
        if t == abi.LIGHT_DIRECTIONAL:
            children.append({"type": "SCENE_NODE_DIRECTIONAL_LIGHT", **base, "position": [0.0, 0.0, 0.0], "rotation": [_f(v) for v in euler_degrees(*_basis_for_forward(l["light_data1"][:3]))]})
        elif t == abi.LIGHT_POINT:
            children.append({"type": "SCENE_NODE_POINT_LIGHT", **base, "position": [_f(v) for v in l["light_data2"][:3]], "rotation": [0.0, 0.0, 0.0]})
        elif t == abi.LIGHT_SPOT:
            inner = math.degrees(math.acos(min(1.0, float(l["light_data3"][0]))))
            outer = math.degrees(math.acos(min(1.0, float(l["light_data3"][1]))))
            children.append({"type": "SCENE_NODE_SPOT_LIGHT", **base, "position": [_f(v) for v in l["light_data2"][:3]], "rotation": [_f(v) for v in euler_degrees(*_basis_for_forward(l["light_data1"][:3]))], "inner_cone_angle": _f(inner), "outer_cone_angle": _f(outer)})
        # area and environment lights are derived by Scene::update from emissive submeshes / the IBL node
    if scene.env_cube is not None:
        size, faces = scene.env_cube
        faces = np.ascontiguousarray(faces, np.float32).reshape(6, size, size, 4)
        write_image(assets / "texture" / "env.ast", "env", [[(size, size, faces[f].tobytes())] for f in range(6)], 4, PIXEL_FLOAT32)
        children.append({"type": "SCENE_NODE_IBL", "name": "ibl", "image": "texture/env.ast", "children": []})
    doc = {"name": name, "scene_graph": {"type": "SCENE_NODE_ROOT", "name": "root", "position": [0.0, 0.0, 0.0], "rotation": [0.0, 0.0, 0.0], "scale": [1.0, 1.0, 1.0], "children": children}}
    (assets / "scene" / f"{name}.json").write_text(json.dumps(doc, indent=1))
    return f"scene/{name}.json"
