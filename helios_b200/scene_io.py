"""Scene description files for the C++ host layer (helios_b200/shim): `export_scene` writes a SceneData as the
*authoring-level* objects the reference's engine API takes — textures, Material parameters, meshes with
SubMesh ranges, mesh nodes with model matrices, light nodes, camera node — so that `helios_headless` can
rebuild the scene through Material::create / Mesh::create / MeshNode / ...LightNode / CameraNode and let
Scene::update derive the GPU tables itself.  (The reference's own on-disk format, AssetCore .ast/JSON, is
SURVEY.md §8 row f1.)

Layout (little endian):  "HLSC0001", u32 width, height, max_ray_bounces, f32 shadow_ray_bias,
  textures   u32 n { i32 format, u32 w, u32 h, bytes }
  materials  u32 n { u32 type, u32 alpha_test, f32 albedo[4] (sRGB, as authored), f32 emissive[4], f32 metallic, f32 roughness,
                     i32 albedo_tex, normal_tex, metallic_tex, roughness_tex, emissive_tex, i32 roughness_channel, metallic_channel }
  meshes     u32 n { u32 nv, ni, nsub, Vertex[nv], u32[ni], { u32 mat_idx, index_count, vertex_count, base_vertex, base_index }[nsub],
                     u32 n_mat, u32 material_id[n_mat] }
  mesh nodes u32 n { u32 mesh, f32 model[16] column-major }
  camera     f32 position[3], quat (w,x,y,z), fov, near, far, focal_length, aperture_radius
  dir lights u32 n { quat, f32 color[3], intensity, radius }
  point      u32 n { f32 position[3], color[3], intensity, radius }
  spot       u32 n { f32 position[3], quat, color[3], intensity, radius, inner_deg, outer_deg }
  ibl        u32 size (0 = none), f32 faces[6*size*size*4]
"""
from __future__ import annotations

import math
import struct

import numpy as np

from . import abi

MAGIC = b"HLSC0001"


def _quat_from_basis(right, up, forward):
    """rotation matrix with columns (right, up, forward) -> unit quaternion (w, x, y, z)"""
    m = np.stack([right, up, forward], 1).astype(np.float64)
    tr = m[0, 0] + m[1, 1] + m[2, 2]
    if tr > 0:
        s = math.sqrt(tr + 1.0) * 2
        w, x, y, z = 0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = math.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
        w, x, y, z = (m[2, 1] - m[1, 2]) / s, 0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s
    elif m[1, 1] > m[2, 2]:
        s = math.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
        w, x, y, z = (m[0, 2] - m[2, 0]) / s, (m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s
    else:
        s = math.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
        w, x, y, z = (m[1, 0] - m[0, 1]) / s, (m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s
    q = np.array([w, x, y, z])
    return q / np.linalg.norm(q)


def _quat_for_forward(fwd):
    """a rotation whose +Z axis is `fwd` (TransformNode::forward() = q * (0,0,1))"""
    f = np.asarray(fwd, np.float64)
    f = f / np.linalg.norm(f)
    ref = np.array([0.0, 1.0, 0.0]) if abs(f[1]) < 0.99 else np.array([1.0, 0.0, 0.0])
    r = np.cross(ref, f)
    r /= np.linalg.norm(r)
    u = np.cross(f, r)
    return _quat_from_basis(r, u, f)


def export_scene(scene, path):
    """SceneData -> scene description file.  Material indices in the file are the rows of scene.materials; the
    C++ side re-derives table order by first use, exactly as the reference does."""
    out = bytearray()
    w = out.extend
    w(MAGIC)
    w(struct.pack("<IIIf", scene.width, scene.height, int(scene.max_ray_bounces), float(scene.shadow_ray_bias)))
    w(struct.pack("<I", len(scene.textures)))
    for fmt, tw, th, data in scene.textures:
        w(struct.pack("<iII", int(fmt), int(tw), int(th)))
        w(np.ascontiguousarray(data).tobytes())
    # which materials are used by a non-opaque geometry -> alpha tested
    alpha = np.zeros(len(scene.materials), bool)
    for m in scene.meshes:
        for i, s in enumerate(m.submeshes):
            if not int(s["opaque"]):
                alpha[m.materials[i]] = True
    w(struct.pack("<I", len(scene.materials)))
    for k, m in enumerate(scene.materials):
        a = m["albedo"].astype(np.float64)
        src = np.array([*np.power(np.maximum(a[:3], 0.0), 1.0 / 2.2), a[3]], np.float32)  # undo scene.cpp's pow(rgb, 2.2)
        t0, t1 = m["texture_indices0"], m["texture_indices1"]
        w(struct.pack("<II", 0, int(alpha[k])))
        w(src.tobytes())
        w(m["emissive"].astype(np.float32).tobytes())
        w(struct.pack("<ff", float(m["roughness_metallic"][1]), float(m["roughness_metallic"][0])))
        w(struct.pack("<iiiiiii", int(t0[0]), int(t0[1]), int(t0[3]), int(t0[2]), int(t1[0]), int(t1[2]), int(t1[3])))
    w(struct.pack("<I", len(scene.meshes)))
    for m in scene.meshes:
        v, idx = np.ascontiguousarray(m.vertices), np.ascontiguousarray(m.indices, np.uint32)
        w(struct.pack("<III", len(v), len(idx), len(m.submeshes)))
        w(v.tobytes())
        w(idx.tobytes())
        for i, s in enumerate(m.submeshes):
            w(struct.pack("<IIIII", i, int(s["index_count"]), int(s["vertex_count"]), 0, int(s["base_index"])))
        w(struct.pack("<I", len(m.materials)))
        w(np.asarray(m.materials, np.uint32).tobytes())
    w(struct.pack("<I", len(scene.instances)))
    for inst in scene.instances:
        w(struct.pack("<I", int(inst["mesh_index"])))
        w(inst["model_matrix"].astype(np.float32).tobytes())
    c = scene.camera
    w(np.asarray(c.position, np.float32).tobytes())
    w(_quat_from_basis(c.right, c.up, c.forward).astype(np.float32).tobytes())
    w(struct.pack("<fffff", c.fov, c.near, c.far, c.focal_length, c.aperture_radius))
    rows = {t: [l for l in scene.lights if int(l["light_data0"][0]) == t] for t in (abi.LIGHT_DIRECTIONAL, abi.LIGHT_POINT, abi.LIGHT_SPOT)}
    w(struct.pack("<I", len(rows[abi.LIGHT_DIRECTIONAL])))
    for l in rows[abi.LIGHT_DIRECTIONAL]:
        w(_quat_for_forward(l["light_data1"][:3]).astype(np.float32).tobytes())
        w(l["light_data0"][1:4].astype(np.float32).tobytes())
        w(struct.pack("<ff", float(l["light_data1"][3]), float(l["light_data2"][3])))
    w(struct.pack("<I", len(rows[abi.LIGHT_POINT])))
    for l in rows[abi.LIGHT_POINT]:
        w(l["light_data2"][:3].astype(np.float32).tobytes())
        w(l["light_data0"][1:4].astype(np.float32).tobytes())
        w(struct.pack("<ff", float(l["light_data1"][3]), float(l["light_data2"][3])))
    w(struct.pack("<I", len(rows[abi.LIGHT_SPOT])))
    for l in rows[abi.LIGHT_SPOT]:
        w(l["light_data2"][:3].astype(np.float32).tobytes())
        w(_quat_for_forward(l["light_data1"][:3]).astype(np.float32).tobytes())
        w(l["light_data0"][1:4].astype(np.float32).tobytes())
        inner = math.degrees(math.acos(min(1.0, float(l["light_data3"][0]))))
        outer = math.degrees(math.acos(min(1.0, float(l["light_data3"][1]))))
        w(struct.pack("<ffff", float(l["light_data1"][3]), float(l["light_data2"][3]), inner, outer))
    if scene.env_cube is not None:
        size, faces = scene.env_cube
        w(struct.pack("<I", int(size)))
        w(np.ascontiguousarray(faces, np.float32).tobytes())
    else:
        w(struct.pack("<I", 0))
    with open(path, "wb") as f:
        f.write(out)
    return path
