"""Procedural, seeded scene generators for the BASELINE.json configurations.

The reference ships no assets (README.md:32-35) and there is no network, so every scene is generated.
A generator returns a SceneData holding exactly what the reference uploads for the path-trace pass:
per-mesh Vertex/index/submesh arrays (include/resource/mesh.h:10-29), the Material / Instance / Light
tables and the per-instance submesh table in the order Scene::create_gpu_resources produces them
(src/engine/resource/scene.cpp:915-1311: area lights first, then env, directional, point, spot), the
environment (cube map or Hosek-Wilkie sun direction) and a camera.  The same arrays feed the CUDA path
(through the C ABI) and the CPU oracle (tests only).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import abi


# ----------------------------------------------------------------------------------------------
# containers
# ----------------------------------------------------------------------------------------------
@dataclass
class MeshData:
    vertices: np.ndarray  # abi.VERTEX
    indices: np.ndarray  # uint32
    submeshes: np.ndarray  # abi.SUBMESH
    materials: list  # scene-global material index per submesh


@dataclass
class Camera:
    position: np.ndarray
    right: np.ndarray  # q * (1,0,0)   (TransformNode::left(), scene.cpp:220-223)
    up: np.ndarray  # q * (0,1,0)
    forward: np.ndarray  # q * (0,0,1); the camera looks along -forward (CameraNode::camera_forward)
    fov: float = 60.0  # include/resource/scene.h:256-260 defaults
    near: float = 1.0
    far: float = 1000.0
    focal_length: float = 8.0
    aperture_radius: float = 0.1

    @staticmethod
    def look_at(position, target, up=(0, 1, 0), **kw) -> "Camera":
        position = np.asarray(position, np.float64)
        view = np.asarray(target, np.float64) - position
        view /= np.linalg.norm(view)
        fwd = -view
        r = np.cross(np.asarray(up, np.float64), fwd)
        r /= np.linalg.norm(r)
        u = np.cross(fwd, r)
        return Camera(position.astype(np.float32), r.astype(np.float32), u.astype(np.float32), fwd.astype(np.float32), **kw)


@dataclass
class SceneData:
    name: str
    width: int
    height: int
    meshes: list
    materials: np.ndarray
    instances: np.ndarray
    submesh_info: list  # per instance: uint32 [n_submeshes, 2] = (base_index/3, material index)
    lights: np.ndarray
    camera: Camera
    textures: list = field(default_factory=list)  # (format, w, h, ndarray)
    env_cube: tuple | None = None  # (size, float32[6,size,size,4])
    sun_direction: np.ndarray | None = None  # towards the sun (= -DirectionalLightNode.forward()); Hosek sky
    max_ray_bounces: int = 8
    shadow_ray_bias: float = 0.0

    @property
    def num_triangles(self) -> int:
        return int(sum(int(self.meshes[i["mesh_index"]].indices.size) // 3 for i in self.instances))

    def push_constants(self, num_frames: int, tile=(0, 0), max_ray_bounces=None, shadow_ray_bias=None, pixel_coord=(0, 0)) -> np.ndarray:
        """PathIntegrator::launch_rays, src/engine/gfx/path_integrator.cpp:136-161."""
        cam = self.camera
        W, H = self.width, self.height
        f = 1.0 / math.tan(math.radians(cam.fov) / 2.0)
        aspect = W / H
        P = np.zeros((4, 4), np.float64)  # glm::perspective (GL clip space), row-major math here
        P[0, 0] = f / aspect
        P[1, 1] = f
        P[2, 2] = -(cam.far + cam.near) / (cam.far - cam.near)
        P[3, 2] = -1.0
        P[2, 3] = -(2.0 * cam.far * cam.near) / (cam.far - cam.near)
        TR = np.eye(4)
        TR[:3, 0], TR[:3, 1], TR[:3, 2], TR[:3, 3] = cam.right, cam.up, cam.forward, cam.position
        V = np.linalg.inv(TR)  # CameraNode::update, scene.cpp:638-639
        vpi = np.linalg.inv(P @ V)
        pc = np.zeros((), abi.PUSH_CONSTANTS)
        pc["view_proj_inverse"] = vpi.T.reshape(16).astype(np.float32)  # column-major
        view_dir = -cam.forward.astype(np.float32)  # `forward` in launch_rays
        pc["camera_pos"] = [*cam.position, 0.0]
        pc["up_direction"] = [*cam.up, 0.0]
        pc["right_direction"] = [*cam.right, 0.0]
        fpp = cam.position + view_dir * np.float32(cam.focal_length)
        fp = -view_dir
        pc["focal_plane"] = [*fp, -(fp[0] * fpp[0] + fp[1] * fpp[1] + fp[2] * fpp[2])]
        pc["ray_debug_pixel_coord"] = [pixel_coord[0], H - pixel_coord[1], W, H]  # path_integrator.cpp:146 (gather_debug_rays passes the clicked pixel)
        pc["launch_id_size"] = [tile[0], tile[1], W, H]
        pc["accumulation"] = num_frames / (num_frames + 1.0)
        pc["num_lights"] = len(self.lights)
        pc["num_frames"] = num_frames
        pc["max_ray_bounces"] = self.max_ray_bounces if max_ray_bounces is None else max_ray_bounces
        pc["shadow_ray_bias"] = self.shadow_ray_bias if shadow_ray_bias is None else shadow_ray_bias
        pc["focal_length"] = cam.focal_length
        pc["aperture_radius"] = cam.aperture_radius
        return pc


# ----------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------
def make_material(albedo=(0.8, 0.8, 0.8, 1.0), emissive=(0, 0, 0), roughness=1.0, metallic=0.0, srgb_to_linear=True, albedo_tex=-1, normal_tex=-1) -> np.ndarray:
    """One MaterialData row as Scene::create_gpu_resources fills it (scene.cpp:985-1108): constant albedo
    is converted pow(rgb, 2.2) on the host (scene.cpp:1015-1017)."""
    m = np.zeros((), abi.MATERIAL)
    m["texture_indices0"] = [albedo_tex, normal_tex, -1, -1]
    m["texture_indices1"] = [-1, -1, -1, -1]
    a = np.asarray(albedo, np.float32)
    if albedo_tex == -1:
        rgb = np.power(a[:3], np.float32(2.2)) if srgb_to_linear else a[:3]
        m["albedo"] = [*rgb, a[3]]
    m["emissive"] = [*emissive, 0.0]
    m["roughness_metallic"] = [roughness, metallic, 0, 0]
    return m


def make_instance(model=None, mesh_index=0) -> np.ndarray:
    """InstanceData (scene.cpp:1260-1265); normal_matrix = global transform without scale (scene.cpp:277-280)."""
    inst = np.zeros((), abi.INSTANCE)
    M = np.eye(4) if model is None else np.asarray(model, np.float64)
    inst["model_matrix"] = M.T.reshape(16).astype(np.float32)
    # strip scale from the upper 3x3 columns
    N = M.copy()
    for c in range(3):
        n = np.linalg.norm(N[:3, c])
        if n > 0:
            N[:3, c] /= n
    inst["normal_matrix"] = N.T.reshape(16).astype(np.float32)
    inst["mesh_index"] = mesh_index
    return inst


def trs(translate=(0, 0, 0), scale=(1, 1, 1), rot_y_deg=0.0) -> np.ndarray:
    c, s = math.cos(math.radians(rot_y_deg)), math.sin(math.radians(rot_y_deg))
    R = np.array([[c, 0, s, 0], [0, 1, 0, 0], [-s, 0, c, 0], [0, 0, 0, 1]], np.float64)
    S = np.diag([scale[0], scale[1], scale[2], 1.0])
    T = np.eye(4)
    T[:3, 3] = translate
    return T @ R @ S


def make_vertices(pos, uv=None, normal=None, tangent=None, bitangent=None, submesh_index=None) -> np.ndarray:
    """Vertex widening as ResourceManager::load_mesh_internal does it: w = 0 except position.w = submesh index
    (core/resource_manager.cpp:446-473)."""
    n = len(pos)
    v = np.zeros(n, abi.VERTEX)
    v["position"][:, :3] = pos
    if submesh_index is not None:
        v["position"][:, 3] = submesh_index
    if uv is not None:
        v["tex_coord"][:, :2] = uv
    if normal is not None:
        v["normal"][:, :3] = normal
        if tangent is None:
            nn = np.asarray(normal, np.float64)
            ref = np.where(np.abs(nn[:, 1:2]) > 0.99, np.array([[1.0, 0, 0]]), np.array([[0, 1.0, 0]]))
            t = np.cross(ref, nn)
            t /= np.maximum(np.linalg.norm(t, axis=1, keepdims=True), 1e-20)
            tangent = t
            bitangent = np.cross(nn, t)
    if tangent is not None:
        v["tangent"][:, :3] = tangent
    if bitangent is not None:
        v["bitangent"][:, :3] = bitangent
    return v


def flat_normals(pos, idx):
    p = np.asarray(pos, np.float64)
    tri = np.asarray(idx).reshape(-1, 3)
    n = np.cross(p[tri[:, 1]] - p[tri[:, 0]], p[tri[:, 2]] - p[tri[:, 0]])
    n /= np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-30)
    return n


def _quad(p0, p1, p2, p3):
    """two triangles (p0,p1,p2), (p0,p2,p3), unshared vertices; returns positions[4], indices[6], normal"""
    pos = np.array([p0, p1, p2, p3], np.float64)
    n = np.cross(pos[1] - pos[0], pos[2] - pos[0])
    n /= np.linalg.norm(n)
    return pos, np.array([0, 1, 2, 0, 2, 3], np.uint32), n


class _MeshBuilder:
    """accumulates submeshes (contiguous index ranges, one material each) into one mesh"""

    def __init__(self):
        self.pos, self.uv, self.nrm, self.sub_of_vertex = [], [], [], []
        self.idx = []
        self.subs = []  # (base_index, index_count, vertex_count, opaque, material)
        self._nv = 0
        self._ni = 0

    def begin_submesh(self, material, opaque=True):
        self._cur = dict(material=material, opaque=opaque, base=self._ni, v0=self._nv)

    def add(self, pos, idx, nrm, uv=None):
        pos = np.asarray(pos, np.float64)
        self.pos.append(pos)
        self.nrm.append(np.broadcast_to(np.asarray(nrm, np.float64), pos.shape))
        self.uv.append(np.zeros((len(pos), 2)) if uv is None else np.asarray(uv, np.float64))
        self.sub_of_vertex.append(np.full(len(pos), len(self.subs), np.float32))
        self.idx.append(np.asarray(idx, np.uint32) + np.uint32(self._nv))
        self._nv += len(pos)
        self._ni += len(idx)

    def end_submesh(self):
        c = self._cur
        self.subs.append((c["base"], self._ni - c["base"], self._nv - c["v0"], 1 if c["opaque"] else 0, c["material"]))

    def build(self) -> MeshData:
        pos = np.concatenate(self.pos)
        v = make_vertices(pos, np.concatenate(self.uv), np.concatenate(self.nrm), submesh_index=np.concatenate(self.sub_of_vertex))
        subs = np.zeros(len(self.subs), abi.SUBMESH)
        for i, s in enumerate(self.subs):
            subs[i] = (s[0], s[1], s[2], s[3])
        return MeshData(v, np.concatenate(self.idx).astype(np.uint32), subs, [s[4] for s in self.subs])


def submesh_table(mesh: MeshData, material_override=None) -> np.ndarray:
    """(base_index / 3, global material index) per submesh, scene.cpp:1131-1146"""
    t = np.zeros((len(mesh.submeshes), 2), np.uint32)
    for i, s in enumerate(mesh.submeshes):
        t[i, 0] = s["base_index"] // 3
        t[i, 1] = mesh.materials[i] if material_override is None else material_override
    return t


def area_lights_for(mesh: MeshData, instance_index: int, materials: np.ndarray) -> list:
    """AREA LightData rows for every emissive submesh (scene.cpp:1109-1119): light_data0 = (4, mesh-node index,
    material index, base_index/3), light_data1 = (triangle count, 0, 0, 0)."""
    out = []
    for i, s in enumerate(mesh.submeshes):
        m = materials[mesh.materials[i]]
        if np.any(m["emissive"][:3] != 0):  # Material::is_emissive
            l = np.zeros((), abi.LIGHT)
            l["light_data0"] = [abi.LIGHT_AREA, instance_index, mesh.materials[i], s["base_index"] // 3]
            l["light_data1"] = [s["index_count"] // 3, 0, 0, 0]
            out.append(l)
    return out


def env_light():
    l = np.zeros((), abi.LIGHT)
    l["light_data0"] = [abi.LIGHT_ENVIRONMENT_MAP, 0, 0, 0]
    return l


def directional_light(forward, color=(1, 1, 1), intensity=1.0, radius=0.0):
    """scene.cpp:1275-1284; `forward` = DirectionalLightNode::forward(), the shader negates it."""
    l = np.zeros((), abi.LIGHT)
    l["light_data0"] = [abi.LIGHT_DIRECTIONAL, *color]
    l["light_data1"] = [*forward, intensity]
    l["light_data2"] = [0, 0, 0, radius]
    return l


def point_light(position, color=(1, 1, 1), intensity=1.0, radius=0.0):
    l = np.zeros((), abi.LIGHT)  # scene.cpp:1286-1296
    l["light_data0"] = [abi.LIGHT_POINT, *color]
    l["light_data1"] = [0, 0, 0, intensity]
    l["light_data2"] = [*position, radius]
    return l


def spot_light(position, forward, color=(1, 1, 1), intensity=1.0, radius=0.0, inner_deg=30.0, outer_deg=45.0):
    l = np.zeros((), abi.LIGHT)  # scene.cpp:1298-1308
    l["light_data0"] = [abi.LIGHT_SPOT, *color]
    l["light_data1"] = [*forward, intensity]
    l["light_data2"] = [*position, radius]
    l["light_data3"] = [math.cos(math.radians(inner_deg)), math.cos(math.radians(outer_deg)), 0, 0]
    return l


def _stack(rows, dtype):
    a = np.zeros(len(rows), dtype)
    for i, r in enumerate(rows):
        a[i] = r
    return a


# ----------------------------------------------------------------------------------------------
# config 1: Cornell box
# ----------------------------------------------------------------------------------------------
def cornell_box(width=512, height=512, aperture_radius=0.0, max_ray_bounces=8, shadow_ray_bias=0.0) -> SceneData:
    """5 walls + 2 boxes + 1 quad light = 36 triangles, one mesh at identity with one submesh per material
    (white / red / green / light); the light quad's triangle 0 faces down (only it is sampled)."""
    WHITE, RED, GREEN, LIGHT = 0, 1, 2, 3
    mats = _stack(
        [
            make_material((0.73, 0.73, 0.73, 1.0)),
            make_material((0.65, 0.05, 0.05, 1.0)),
            make_material((0.12, 0.45, 0.15, 1.0)),
            make_material((0.0, 0.0, 0.0, 1.0), emissive=(15.0, 15.0, 15.0)),
        ],
        abi.MATERIAL,
    )
    b = _MeshBuilder()
    S = 1.0  # box spans [-1,1]^3

    def add_quad(p0, p1, p2, p3):
        pos, idx, n = _quad(p0, p1, p2, p3)
        b.add(pos, idx, n, uv=[[0, 0], [1, 0], [1, 1], [0, 1]])

    def add_box(cx, cz, hx, hy, hz, rot_deg):
        c, s = math.cos(math.radians(rot_deg)), math.sin(math.radians(rot_deg))

        def P(x, y, z):
            return (cx + c * x + s * z, y, cz - s * x + c * z)

        y0, y1 = -S, -S + 2 * hy
        c000, c100, c110, c010 = P(-hx, y0, -hz), P(hx, y0, -hz), P(hx, y1, -hz), P(-hx, y1, -hz)
        c001, c101, c111, c011 = P(-hx, y0, hz), P(hx, y0, hz), P(hx, y1, hz), P(-hx, y1, hz)
        add_quad(c001, c101, c111, c011)  # +z
        add_quad(c100, c000, c010, c110)  # -z
        add_quad(c101, c100, c110, c111)  # +x
        add_quad(c000, c001, c011, c010)  # -x
        add_quad(c011, c111, c110, c010)  # top
        add_quad(c000, c100, c101, c001)  # bottom

    b.begin_submesh(WHITE)
    add_quad((-S, -S, S), (S, -S, S), (S, -S, -S), (-S, -S, -S))  # floor (normal +y)
    add_quad((-S, S, -S), (S, S, -S), (S, S, S), (-S, S, S))  # ceiling (normal -y)
    add_quad((-S, -S, -S), (S, -S, -S), (S, S, -S), (-S, S, -S))  # back wall (normal +z)
    add_box(0.33, 0.35, 0.3, 0.3, 0.3, -18.0)
    add_box(-0.35, -0.3, 0.3, 0.6, 0.3, 20.0)
    b.end_submesh()
    b.begin_submesh(RED)
    add_quad((-S, -S, S), (-S, -S, -S), (-S, S, -S), (-S, S, S))  # left wall (normal +x)
    b.end_submesh()
    b.begin_submesh(GREEN)
    add_quad((S, -S, -S), (S, -S, S), (S, S, S), (S, S, -S))  # right wall (normal -x)
    b.end_submesh()
    b.begin_submesh(LIGHT)
    L, y = 0.25, S - 0.005
    add_quad((-L, y, -L), (L, y, -L), (L, y, L), (-L, y, L))  # normal -y (faces down)
    b.end_submesh()
    mesh = b.build()
    instances = _stack([make_instance()], abi.INSTANCE)
    lights = _stack(area_lights_for(mesh, 0, mats), abi.LIGHT)
    cam = Camera.look_at((0.0, 0.0, 3.4), (0.0, 0.0, 0.0), fov=60.0, near=0.1, far=100.0, focal_length=3.4, aperture_radius=aperture_radius)
    return SceneData("cornell", width, height, [mesh], mats, instances, [submesh_table(mesh)], lights, cam, max_ray_bounces=max_ray_bounces, shadow_ray_bias=shadow_ray_bias)


# ----------------------------------------------------------------------------------------------
# config 5: uniform triangle soup
# ----------------------------------------------------------------------------------------------
def triangle_soup(n_triangles: int, width=1920, height=1080, seed=None) -> SceneData:
    """uniform random triangles in the unit cube, edge length ~ N^(-1/3), seed = N; pinhole camera."""
    rng = np.random.default_rng(n_triangles if seed is None else seed)
    edge = 1.5 * n_triangles ** (-1.0 / 3.0)
    c = rng.random((n_triangles, 1, 3), dtype=np.float32)
    off = (rng.random((n_triangles, 3, 3), dtype=np.float32) - 0.5) * np.float32(edge)
    pos = (c + off).reshape(-1, 3)
    idx = np.arange(3 * n_triangles, dtype=np.uint32)
    nrm = np.repeat(flat_normals(pos, idx), 3, axis=0).astype(np.float32)
    v = np.zeros(3 * n_triangles, abi.VERTEX)
    v["position"][:, :3] = pos
    v["normal"][:, :3] = nrm
    v["tangent"][:, 0] = 1.0
    v["bitangent"][:, 2] = 1.0
    subs = np.zeros(1, abi.SUBMESH)
    subs[0] = (0, 3 * n_triangles, 3 * n_triangles, 1)
    mesh = MeshData(v, idx, subs, [0])
    mats = _stack([make_material((0.7, 0.7, 0.7, 1.0))], abi.MATERIAL)
    instances = _stack([make_instance()], abi.INSTANCE)
    lights = _stack([point_light((0.5, 3.0, 0.5), intensity=20.0)], abi.LIGHT)
    cam = Camera.look_at((0.5, 0.5, 2.2), (0.5, 0.5, 0.5), fov=60.0, near=0.1, far=100.0, focal_length=1.7, aperture_radius=0.0)
    return SceneData(f"soup{n_triangles}", width, height, [mesh], mats, instances, [submesh_table(mesh)], lights, cam)


# ----------------------------------------------------------------------------------------------
# config 2: procedural terrain + icospheres, Hosek-Wilkie sky
# ----------------------------------------------------------------------------------------------
def _value_noise(x, y, seed):
    xi, yi = np.floor(x).astype(np.int64), np.floor(y).astype(np.int64)
    fx, fy = x - xi, y - yi

    def h(ix, iy):
        n = (ix * 374761393 + iy * 668265263 + seed * 2147483647) & 0xFFFFFFFF
        n = ((n ^ (n >> 13)) * 1274126177) & 0xFFFFFFFF
        n = n ^ (n >> 16)
        return (n & 0xFFFFFF) / float(0x1000000)

    sx, sy = fx * fx * (3 - 2 * fx), fy * fy * (3 - 2 * fy)
    a, b_, c, d = h(xi, yi), h(xi + 1, yi), h(xi, yi + 1), h(xi + 1, yi + 1)
    return (a * (1 - sx) + b_ * sx) * (1 - sy) + (c * (1 - sx) + d * sx) * sy


def _fbm(x, y, seed, octaves=5):
    amp, freq, out = 0.5, 1.0, np.zeros_like(x)
    for o in range(octaves):
        out += amp * _value_noise(x * freq, y * freq, seed + o)
        amp *= 0.5
        freq *= 2.0
    return out


def _icosphere(level):
    t = (1.0 + math.sqrt(5.0)) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    v = [np.array(p, np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(level):
        cache, nf = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                cache[key] = len(v) - 1
            return cache[key]

        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return np.array(v), np.array(f, np.uint32)


def terrain_scene(grid=700, n_spheres=16, sphere_level=3, width=1920, height=1080, seed=1, sun_elevation_deg=45.0, textured=False) -> SceneData:
    """config 2: grid x grid displaced quads (fBm) in 4 material strips + n_spheres icospheres in 4 material
    groups = 8 constant materials, one mesh, identity instance; one directional light at the given elevation ->
    Hosek-Wilkie sky cube map + environment light + directional light (scene.cpp:891-899, 1269-1284).
    Default sizes give 2*700^2 + 16*1280 = 1,000,480 triangles."""
    palette = [(0.35, 0.5, 0.2), (0.55, 0.45, 0.3), (0.5, 0.5, 0.5), (0.85, 0.85, 0.9), (0.8, 0.2, 0.2), (0.2, 0.3, 0.8), (0.9, 0.8, 0.3), (0.95, 0.95, 0.95)]
    rough = [0.9, 0.8, 0.6, 0.4, 0.3, 0.5, 0.2, 0.15]
    metal = [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 1.0]
    textures = []
    mats = []
    for i in range(8):
        tex = -1
        if textured and i < 4:
            textures.append(checker_texture(1024, seed + i, palette[i]))
            tex = len(textures) - 1
        mats.append(make_material((*palette[i], 1.0), roughness=rough[i], metallic=metal[i], albedo_tex=tex))
    mats = _stack(mats, abi.MATERIAL)

    extent = 40.0
    n = grid + 1
    gx, gz = np.meshgrid(np.linspace(-extent / 2, extent / 2, n), np.linspace(-extent / 2, extent / 2, n), indexing="xy")
    hscale = 4.0

    def height_at(x, z):
        return hscale * (_fbm(x * 0.15 + 100.0, z * 0.15 + 100.0, seed) - 0.5)

    gy = height_at(gx, gz)
    pos = np.stack([gx, gy, gz], -1).reshape(-1, 3)
    # analytic-ish normals by central differences on the grid
    dydx = np.gradient(gy, axis=1) / (extent / grid)
    dydz = np.gradient(gy, axis=0) / (extent / grid)
    nrm = np.stack([-dydx, np.ones_like(gy), -dydz], -1).reshape(-1, 3)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    uv = np.stack([gx / 4.0, gz / 4.0], -1).reshape(-1, 2)
    # quads: strips along z, 4 material strips
    qi, qj = np.meshgrid(np.arange(grid), np.arange(grid), indexing="xy")  # qi: x index, qj: z index
    v00 = (qj * n + qi).reshape(-1)
    v10, v01, v11 = v00 + 1, v00 + n, v00 + n + 1
    tris = np.stack([v00, v01, v11, v00, v11, v10], -1).astype(np.uint32)  # CCW seen from +y
    strip_of_quad = (qj.reshape(-1) * 4) // grid
    vert_sub = np.minimum((np.arange(n)[:, None] * 4) // grid, 3).repeat(n, 1).reshape(-1).astype(np.float32)

    all_pos, all_nrm, all_uv, all_sub, all_idx, subs, sub_mats = [pos], [nrm], [uv], [vert_sub], [], [], []
    ni = 0
    for s in range(4):
        t = tris[strip_of_quad == s].reshape(-1)
        all_idx.append(t)
        subs.append((ni, t.size, pos.shape[0], 1))
        sub_mats.append(s)
        ni += t.size
    nv = pos.shape[0]
    sv, sf = _icosphere(sphere_level)
    rng = np.random.default_rng(seed)
    groups = [[] for _ in range(4)]
    for k in range(n_spheres):
        groups[k % 4].append(k)
    centers = (rng.random((n_spheres, 2)) - 0.5) * extent * 0.7
    radii = 0.6 + rng.random(n_spheres) * 1.2
    for g in range(4):
        base = ni
        v0 = nv
        for k in groups[g]:
            cx, cz = centers[k]
            cy = height_at(np.array([cx]), np.array([cz]))[0] + radii[k] * 0.8
            p = sv * radii[k] + np.array([cx, cy, cz])
            all_pos.append(p)
            all_nrm.append(sv)
            all_uv.append(np.stack([np.arctan2(sv[:, 2], sv[:, 0]) / (2 * math.pi) + 0.5, np.arccos(np.clip(sv[:, 1], -1, 1)) / math.pi], -1))
            all_sub.append(np.full(len(sv), 4 + g, np.float32))
            all_idx.append((sf.reshape(-1) + np.uint32(nv)).astype(np.uint32))
            nv += len(sv)
            ni += sf.size
        if ni > base:
            subs.append((base, ni - base, nv - v0, 1))
            sub_mats.append(4 + g)
    v = make_vertices(np.concatenate(all_pos), np.concatenate(all_uv), np.concatenate(all_nrm), submesh_index=np.concatenate(all_sub))
    sm = np.zeros(len(subs), abi.SUBMESH)
    for i, s in enumerate(subs):
        sm[i] = s
    mesh = MeshData(v, np.concatenate(all_idx).astype(np.uint32), sm, sub_mats)
    instances = _stack([make_instance()], abi.INSTANCE)
    el = math.radians(sun_elevation_deg)
    sun = np.array([math.cos(el) * 0.6, math.sin(el), math.cos(el) * 0.8], np.float64)
    sun /= np.linalg.norm(sun)
    lights = _stack([env_light(), directional_light(-sun, color=(1.0, 0.95, 0.85), intensity=3.0, radius=0.02)], abi.LIGHT)
    cam = Camera.look_at((0.0, 6.0, 19.0), (0.0, 0.0, 0.0), fov=60.0, near=1.0, far=1000.0, focal_length=18.0, aperture_radius=0.05)
    return SceneData(f"terrain{grid}", width, height, [mesh], mats, instances, [submesh_table(mesh)], lights, cam, textures=textures, sun_direction=sun.astype(np.float32))


def checker_texture(size, seed, base_rgb, alpha_disc=False):
    """procedural RGBA8 sRGB texture: noisy checker in the base colour; alpha_disc -> alpha in {0,255} disc mask"""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:size, 0:size]
    cell = max(size // 16, 1)
    chk = (((x // cell) + (y // cell)) & 1).astype(np.float32)
    noise = rng.random((size, size), dtype=np.float32) * 0.15
    lum = 0.6 + 0.4 * chk - noise
    rgb = np.clip(lum[..., None] * np.asarray(base_rgb, np.float32)[None, None, :] * 255.0, 0, 255).astype(np.uint8)
    a = np.full((size, size, 1), 255, np.uint8)
    if alpha_disc:
        r2 = (x - size / 2 + 0.5) ** 2 + (y - size / 2 + 0.5) ** 2
        a[..., 0] = np.where(r2 < (0.45 * size) ** 2, 255, 0)
    return (abi.TEX_RGBA8_SRGB, size, size, np.ascontiguousarray(np.concatenate([rgb, a], -1)))


# ----------------------------------------------------------------------------------------------
# config 3: alpha-tested foliage, 64 area + 16 punctual lights
# ----------------------------------------------------------------------------------------------
def foliage_scene(n_clusters=50_000, cards_per_cluster=50, width=1920, height=1080, seed=2, ground_grid=64, tex_size=256) -> SceneData:
    """config 3: a stand of trees made of leaf-card clusters (alpha-masked albedo texture, geometry NOT opaque -> any-hit)
    over a ground mesh; 64 area lights = ONE untranslated mesh node with 64 emissive 2-triangle submeshes (lights are
    registered per emissive submesh; only triangle 0 of each quad is sampled, A.8-2) + 8 point + 8 spot lights;
    no directional light and no IBL -> no environment light: num_lights = 80.
    Defaults give 2*64^2 + 50,000*50*2 + 128 = 5,008,320 triangles.
    Layout (round 2: the round-1 layout was one uniform 3 m slab of cards with the camera inside it — 1 % of the pixels were
    lit): clusters of cards (sigma 0.25) are grouped into tree crowns (ellipsoids on a jittered grid, ~125 clusters each)
    with gaps between them, the area lights hang above the crowns, and the camera looks down on the stand from outside, so
    most pixels see lit foliage or lit ground and nearly every shading point draws a non-black light sample (NEE stress)."""
    rng = np.random.default_rng(seed)
    extent = 60.0
    ground_extent = 3.0 * extent
    textures = [checker_texture(tex_size, seed, (0.25, 0.6, 0.2), alpha_disc=True)]
    GROUND, LEAF, EMIT = 0, 1, 2
    mats = _stack(
        [
            make_material((0.45, 0.38, 0.3, 1.0), roughness=0.9),
            make_material((0.0, 0.0, 0.0, 0.0), roughness=0.6, albedo_tex=0),
            make_material((0.0, 0.0, 0.0, 1.0), emissive=(40.0, 36.0, 30.0)),
        ],
        abi.MATERIAL,
    )
    # ground
    n = ground_grid + 1
    gx, gz = np.meshgrid(np.linspace(-ground_extent / 2, ground_extent / 2, n), np.linspace(-ground_extent / 2, ground_extent / 2, n), indexing="xy")
    gy = 1.5 * (_fbm(gx * 0.1 + 50.0, gz * 0.1 + 50.0, seed) - 0.5)
    pos = np.stack([gx, gy, gz], -1).reshape(-1, 3)
    qi, qj = np.meshgrid(np.arange(ground_grid), np.arange(ground_grid), indexing="xy")
    v00 = (qj * n + qi).reshape(-1)
    tris = np.stack([v00, v00 + n, v00 + n + 1, v00, v00 + n + 1, v00 + 1], -1).astype(np.uint32).reshape(-1)
    nrm = np.zeros_like(pos)
    nrm[:, 1] = 1.0
    gv = make_vertices(pos, np.stack([gx / 4, gz / 4], -1).reshape(-1, 2), nrm, submesh_index=np.zeros(len(pos), np.float32))
    gs = np.zeros(1, abi.SUBMESH)
    gs[0] = (0, tris.size, len(pos), 1)
    ground = MeshData(gv, tris, gs, [GROUND])
    # foliage: tree crowns = ellipsoids of clusters, clusters = randomly oriented quads ("leaf cards")
    ncards = n_clusters * cards_per_cluster
    n_trees = max(4, int(round(n_clusters / 125.0)))
    side = int(math.ceil(math.sqrt(n_trees)))
    cell = extent * 0.9 / side
    tk = np.arange(n_trees)
    tree = np.stack([((tk % side) + 0.5) * cell - extent * 0.45, np.zeros(n_trees), ((tk // side) + 0.5) * cell - extent * 0.45], -1)
    tree[:, [0, 2]] += (rng.random((n_trees, 2)) - 0.5) * 0.5 * cell
    tree[:, 1] = 3.0 + rng.random(n_trees) * 1.5
    crown = np.stack([0.42 * cell * (0.8 + 0.4 * rng.random(n_trees)), 1.1 + 0.6 * rng.random(n_trees), 0.42 * cell * (0.8 + 0.4 * rng.random(n_trees))], -1)
    d = rng.normal(size=(n_clusters, 3))
    d *= (rng.random((n_clusters, 1)) ** (1.0 / 3.0)) / np.linalg.norm(d, axis=1, keepdims=True)  # uniform in the unit ball
    owner = np.arange(n_clusters) % n_trees
    cc = tree[owner] + d * crown[owner]
    centre = np.repeat(cc, cards_per_cluster, axis=0) + rng.normal(scale=0.25, size=(ncards, 3))
    a = rng.normal(size=(ncards, 3))
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    b = np.cross(a, rng.normal(size=(ncards, 3)))
    b /= np.linalg.norm(b, axis=1, keepdims=True)
    size = 0.06 + rng.random((ncards, 1)) * 0.05
    a *= size
    b *= size
    quad = np.stack([centre - a - b, centre + a - b, centre + a + b, centre - a + b], 1).reshape(-1, 3)
    fn = np.repeat(np.cross(a, b) / np.linalg.norm(np.cross(a, b), axis=1, keepdims=True), 4, axis=0)
    fuv = np.tile(np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float64), (ncards, 1))
    base = (np.arange(ncards, dtype=np.uint32) * 4)[:, None]
    fidx = (base + np.array([0, 1, 2, 0, 2, 3], np.uint32)[None, :]).reshape(-1).astype(np.uint32)
    fv = make_vertices(quad, fuv, fn, submesh_index=np.zeros(len(quad), np.float32))
    fs = np.zeros(1, abi.SUBMESH)
    fs[0] = (0, fidx.size, len(quad), 0)  # alpha tested -> no VK_GEOMETRY_OPAQUE_BIT (mesh.cpp:73-76)
    foliage = MeshData(fv, fidx, fs, [LEAF])
    # 64 emissive quads (8 x 8 grid above the crowns), one submesh each, facing down
    lb = _MeshBuilder()
    for k in range(64):
        lx, lz = ((k % 8) - 3.5) * extent / 6.5, ((k // 8) - 3.5) * extent / 6.5
        h, y = 1.0, 13.0
        lb.begin_submesh(EMIT)
        p, i, nn = _quad((lx - h, y, lz - h), (lx + h, y, lz - h), (lx + h, y, lz + h), (lx - h, y, lz + h))
        lb.add(p, i, nn, uv=[[0, 0], [1, 0], [1, 1], [0, 1]])
        lb.end_submesh()
    lights_mesh = lb.build()
    meshes = [ground, foliage, lights_mesh]
    instances = _stack([make_instance(mesh_index=k) for k in range(3)], abi.INSTANCE)
    light_rows = area_lights_for(lights_mesh, 2, mats)
    for k in range(8):
        ang = 2 * math.pi * k / 8
        light_rows.append(point_light((20 * math.cos(ang), 8.0, 20 * math.sin(ang)), color=(1.0, 0.9, 0.8), intensity=90.0, radius=0.2))
    for k in range(8):
        ang = 2 * math.pi * (k + 0.5) / 8
        light_rows.append(spot_light((10 * math.cos(ang), 9.0, 10 * math.sin(ang)), forward=(0.0, 1.0, 0.0), color=(0.8, 0.9, 1.0), intensity=160.0, radius=0.1, inner_deg=25.0, outer_deg=25.0))
    cam = Camera.look_at((0.0, 40.0, 36.0), (0.0, 0.0, 3.0), fov=60.0, near=1.0, far=1000.0, focal_length=50.0, aperture_radius=0.0)
    return SceneData(f"foliage{n_clusters}x{cards_per_cluster}", width, height, meshes, mats, instances, [submesh_table(m) for m in meshes], _stack(light_rows, abi.LIGHT), cam, textures=textures)


# ----------------------------------------------------------------------------------------------
# config 4: instanced city
# ----------------------------------------------------------------------------------------------
def _building_mesh(rng, floors, detail, materials):
    """a box tower with `detail` x `detail` window-ledge quads per floor and side: 12 + floors*4*detail*2*... triangles"""
    b = _MeshBuilder()
    w, d, fh = 1.0, 1.0, 0.35
    h = floors * fh

    def add_quad(p0, p1, p2, p3):
        p, i, nn = _quad(p0, p1, p2, p3)
        b.add(p, i, nn, uv=[[0, 0], [1, 0], [1, 1], [0, 1]])

    b.begin_submesh(materials[0])
    # walls as a grid of quads (floors x detail per side) so triangle counts scale; vectorised, quad order
    # side -> floor -> column, vertices (u0,y0) (u1,y0) (u1,y1) (u0,y1)
    kk = np.arange(detail, dtype=np.float64)
    u0, u1 = -w + 2 * w * kk / detail, -w + 2 * w * (kk + 1) / detail
    ff = np.arange(floors, dtype=np.float64)
    y0, y1 = ff * fh, (ff + 1) * fh
    U0, Y0 = np.meshgrid(u0, y0, indexing="xy")  # [floors, detail]
    U1, Y1 = np.meshgrid(u1, y1, indexing="xy")
    ua = np.stack([U0, U1, U1, U0], -1).reshape(-1)  # per-vertex u, 4 per quad
    ya = np.stack([Y0, Y0, Y1, Y1], -1).reshape(-1)
    nq = floors * detail
    qidx = (np.arange(nq, dtype=np.uint32)[:, None] * 4 + np.array([0, 1, 2, 0, 2, 3], np.uint32)[None, :]).reshape(-1)
    quv = np.tile(np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float64), (nq, 1))
    ub = np.stack([U1, U0, U0, U1], -1).reshape(-1)  # side 1 runs the other way
    one = np.ones_like(ua)
    for side in range(4):
        if side == 0:
            pos, nrm = np.stack([ua, ya, d * one], -1), (0.0, 0.0, 1.0)
        elif side == 1:
            pos, nrm = np.stack([ub, ya, -d * one], -1), (0.0, 0.0, -1.0)
        elif side == 2:
            pos, nrm = np.stack([w * one, ya, -ua], -1), (1.0, 0.0, 0.0)
        else:
            pos, nrm = np.stack([-w * one, ya, ua], -1), (-1.0, 0.0, 0.0)
        b.add(pos, qidx, nrm, uv=quv)
    b.end_submesh()
    b.begin_submesh(materials[1])
    add_quad((-w, h, d), (w, h, d), (w, h, -d), (-w, h, -d))  # roof
    # ledges: small protruding boxes' top faces per floor
    for f in range(floors):
        y = f * fh + 0.02
        add_quad((-w * 1.05, y, d * 1.05), (w * 1.05, y, d * 1.05), (w * 1.05, y, d), (-w * 1.05, y, d))
    b.end_submesh()
    return b.build()


def city_scene(n_instances=1023, n_meshes=32, width=3840, height=2160, seed=4, floors=(8, 40), detail=(40, 125)) -> SceneData:
    """config 4: <= 1024 instances (reference limit include/resource/scene.h:14) of <= 32 unique building meshes on a
    grid + a ground instance; directional light -> Hosek-Wilkie sky + environment light.  Defaults: ~20M instanced triangles."""
    rng = np.random.default_rng(seed)
    palette = [(0.6, 0.6, 0.62), (0.5, 0.45, 0.4), (0.7, 0.68, 0.6), (0.35, 0.4, 0.5), (0.8, 0.8, 0.85), (0.3, 0.3, 0.32)]
    mats = [make_material((*c, 1.0), roughness=0.4 + 0.1 * i, metallic=1.0 if i == 4 else 0.0) for i, c in enumerate(palette)]
    mats.append(make_material((0.25, 0.25, 0.25, 1.0), roughness=0.95))  # ground
    mats = _stack(mats, abi.MATERIAL)
    GROUND = len(palette)
    meshes = []
    for m in range(n_meshes):
        fl = int(rng.integers(floors[0], floors[1] + 1))
        dt = int(rng.integers(detail[0], detail[1] + 1))
        meshes.append(_building_mesh(rng, fl, dt, [int(rng.integers(0, 4)), int(4 + rng.integers(0, 2))]))
    side = int(math.ceil(math.sqrt(n_instances)))
    spacing = 3.2
    half = side * spacing / 2
    gb = _MeshBuilder()
    gb.begin_submesh(GROUND)
    p, i, nn = _quad((-half - 5, 0.0, half + 5), (half + 5, 0.0, half + 5), (half + 5, 0.0, -half - 5), (-half - 5, 0.0, -half - 5))
    gb.add(p, i, nn, uv=[[0, 0], [8, 0], [8, 8], [0, 8]])
    gb.end_submesh()
    meshes.append(gb.build())
    inst_rows, tables = [], []
    for k in range(n_instances):
        mi = int(rng.integers(0, n_meshes))
        gx, gz = k % side, k // side
        sc = 0.8 + 0.6 * rng.random()
        M = trs(((gx + 0.5) * spacing - half, 0.0, (gz + 0.5) * spacing - half), (sc, 0.7 + 0.8 * rng.random(), sc), rot_y_deg=float(rng.integers(0, 4)) * 90.0 + rng.normal() * 3.0)
        inst_rows.append(make_instance(M, mesh_index=mi))
        tables.append(submesh_table(meshes[mi]))
    inst_rows.append(make_instance(mesh_index=n_meshes))
    tables.append(submesh_table(meshes[n_meshes]))
    el = math.radians(35.0)
    sun = np.array([math.cos(el) * 0.5, math.sin(el), math.cos(el) * 0.866], np.float64)
    sun /= np.linalg.norm(sun)
    lights = _stack([env_light(), directional_light(-sun, color=(1.0, 0.95, 0.9), intensity=3.0, radius=0.02)], abi.LIGHT)
    cam = Camera.look_at((half * 0.9, half * 0.55, half * 1.1), (0.0, 2.0, 0.0), fov=55.0, near=1.0, far=2000.0, focal_length=half, aperture_radius=0.0)
    return SceneData(f"city{n_instances}", width, height, meshes, mats, _stack(inst_rows, abi.INSTANCE), tables, lights, cam, sun_direction=sun.astype(np.float32))
