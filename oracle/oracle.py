"""ORACLE — TEST INFRASTRUCTURE ONLY (ctypes wrapper around oracle/_build/libhelios_oracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The product package (helios_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libhelios_oracle.so"
_lib = None


def build(force: bool = False) -> Path:
    src = [_HERE / "helios_oracle.cpp", _HERE / "glsl_math.h"]
    if force or not _LIB_PATH.exists() or any(s.stat().st_mtime > _LIB_PATH.stat().st_mtime for s in src):
        subprocess.check_call(["make", "-C", str(_HERE), "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def set_threads(n: int):
    """OpenMP thread count of the oracle libraries (torchrun exports OMP_NUM_THREADS=1 to its workers; libgomp reads the
    variable once, so it is set through the runtime call)"""
    try:
        C.CDLL("libgomp.so.1").omp_set_num_threads(C.c_int(int(n)))
    except OSError:
        pass


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.or_scene_new.restype = C.c_void_p
        _lib.or_scene_add_mesh.restype = C.c_int
        _lib.or_scene_add_texture.restype = C.c_int
        _lib.or_rng_hash.restype = C.c_uint32
        _lib.or_tri_test.restype = C.c_int
    return _lib


# ---- oracle/_ref: the reference's own GLSL compiled as C++ (oracle/ref_glsl/) -----------------------------------
_REF_PATH = _HERE / "_ref" / "libhelios_glsl_ref.so"
_REF_SHADERS = Path("/root/reference/src/engine/shader")
_ref_lib = None


def build_ref(force: bool = False):
    """(Re)build oracle/_ref/libhelios_glsl_ref.so where the reference checkout is mounted; elsewhere (the GPU box)
    the prebuilt file that travelled with the snapshot is used.  Returns its path, or None if there is neither."""
    if _REF_SHADERS.is_dir():
        subprocess.check_call(["make", "-C", str(_HERE), "-s", "ref"] + (["-B"] if force else []))
    return _REF_PATH if _REF_PATH.exists() else None


_REF_AST_TOOL = _HERE / "_ref" / "ref_ast_tool"


def build_ref_ast(force: bool = False):
    """(Re)build oracle/_ref/ref_ast_tool (the reference's AssetCore loader and exporters behind a dump / fixture
    command line, oracle/ref_ast/) where the reference checkout is mounted; elsewhere the prebuilt file is used.
    Returns its path, or None if there is neither."""
    if Path("/root/reference/external/AssetCore/src/loader/loader.cpp").is_file():
        subprocess.check_call(["make", "-C", str(_HERE), "-s", "ref_ast"] + (["-B"] if force else []))
    return _REF_AST_TOOL if _REF_AST_TOOL.exists() else None


_REF_SCENE_TOOL = _HERE / "_ref" / "ref_scene_tool"
_REF_GLM_PIN = _HERE / "_ref" / "glm_pin_ref"


def build_ref_scene(force: bool = False):
    """(Re)build oracle/_ref/ref_scene_tool (the reference's own scene.cpp + material.cpp behind a stand-in device layer,
    oracle/ref_scene/) and oracle/_ref/glm_pin_ref (oracle/ref_scene/glm_pin.cpp against the GLM the reference vendors) where
    the reference checkout is mounted; elsewhere the prebuilt files are used.  Returns (tool, glm_pin) paths, None where absent."""
    if Path("/root/reference/src/engine/resource/scene.cpp").is_file():
        subprocess.check_call(["make", "-C", str(_HERE), "-s", "ref_scene"] + (["-B"] if force else []))
    return (_REF_SCENE_TOOL if _REF_SCENE_TOOL.exists() else None, _REF_GLM_PIN if _REF_GLM_PIN.exists() else None)


_REF_BC_TOOL = _HERE / "_ref" / "ref_bc_tool"


def build_ref_bc(force: bool = False):
    """(Re)build oracle/_ref/ref_bc_tool (BC7 / BC6H decoding by the nvidia-texture-tools sources AssetCore vendors, compiled
    where they lie) where the reference checkout is mounted; elsewhere the prebuilt file is used.  Path, or None."""
    if Path("/root/reference/external/AssetCore/external/nvidia-texture-tools/src/bc7/avpcl.cpp").is_file():
        subprocess.check_call(["make", "-C", str(_HERE), "-s", "ref_bc"] + (["-B"] if force else []))
    return _REF_BC_TOOL if _REF_BC_TOOL.exists() else None


def ref_lib():
    """ctypes handle of the reference-GLSL library (contains the restatement's or_* entry points as well), or None"""
    global _ref_lib
    if _ref_lib is None:
        path = build_ref()
        if path is None:
            return None
        _ref_lib = C.CDLL(str(path))
        _ref_lib.or_scene_new.restype = C.c_void_p
        _ref_lib.or_scene_add_mesh.restype = C.c_int
        _ref_lib.or_scene_add_texture.restype = C.c_int
        _ref_lib.or_rng_hash.restype = C.c_uint32
        _ref_lib.ref_rng_hash.restype = C.c_uint32
        _ref_lib.or_tri_test.restype = C.c_int
    return _ref_lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def hosek_dataset():
    d = np.fromfile(_HERE.parent / "helios_b200" / "data" / "hosek_rgb_v1_4a.f64", dtype="<f8")
    assert d.size == 3600
    return np.ascontiguousarray(d[:3240]), np.ascontiguousarray(d[3240:])


def sky_coeffs(sun_direction, turbidity=4.0, albedo=0.1, normalized_sun_y=1.15) -> np.ndarray:
    rgb, rad = hosek_dataset()
    out = np.zeros(40, np.float32)
    d = np.asarray(sun_direction, np.float32)
    lib().or_sky_coeffs(_p(rgb), _p(rad), _p(d), C.c_float(turbidity), C.c_float(albedo), C.c_float(normalized_sun_y), _p(out))
    return out


def sky_bake(coeffs, sun_direction, size=512) -> np.ndarray:
    out = np.zeros((6, size, size, 4), np.float32)
    cf = np.ascontiguousarray(coeffs, np.float32)
    d = np.ascontiguousarray(sun_direction, np.float32)
    lib().or_sky_bake(_p(cf), _p(d), C.c_uint32(size), _p(out))
    return out


class OracleScene:
    """CPU scene built from a helios_b200.scenes.SceneData."""

    def __init__(self, scene, brute_force: bool = False, sky_size: int = 512, sky_coeffs_override=None, library=None):
        L = self.L = library if library is not None else lib()
        self.scene = scene
        self.h = C.c_void_p(L.or_scene_new())
        self._keep = []
        for m in scene.meshes:
            v, i, s = np.ascontiguousarray(m.vertices), np.ascontiguousarray(m.indices), np.ascontiguousarray(m.submeshes)
            L.or_scene_add_mesh(self.h, _p(v), C.c_uint32(len(v)), _p(i), C.c_uint32(len(i)), _p(s), C.c_uint32(len(s)))
        for fmt, w, h, data in scene.textures:
            d = np.ascontiguousarray(data)
            L.or_scene_add_texture(self.h, C.c_int(fmt), C.c_uint32(w), C.c_uint32(h), _p(d))
        if scene.env_cube is not None:
            size, faces = scene.env_cube
            f = np.ascontiguousarray(faces, np.float32)
            L.or_scene_set_envmap(self.h, C.c_uint32(size), _p(f))
        elif scene.sun_direction is not None:
            cf = sky_coeffs(scene.sun_direction) if sky_coeffs_override is None else sky_coeffs_override
            faces = sky_bake(cf, scene.sun_direction, sky_size)
            L.or_scene_set_envmap(self.h, C.c_uint32(sky_size), _p(faces))
        mats = np.ascontiguousarray(scene.materials)
        inst = np.ascontiguousarray(scene.instances)
        lights = np.ascontiguousarray(scene.lights)
        tabs = [np.ascontiguousarray(t, np.uint32) for t in scene.submesh_info]
        ptrs = (C.c_void_p * len(tabs))(*[t.ctypes.data for t in tabs])
        self._keep += tabs
        L.or_scene_set_tables(self.h, _p(mats), C.c_uint32(len(mats)), _p(inst), ptrs, C.c_uint32(len(inst)), _p(lights), C.c_uint32(len(lights)))
        L.or_scene_set_brute_force(self.h, C.c_int(1 if brute_force else 0))
        self.counters = np.zeros(2, np.uint64)

    def __del__(self):
        try:
            self.L.or_scene_free(self.h)
        except Exception:
            pass

    def set_brute_force(self, on: bool):
        self.L.or_scene_set_brute_force(self.h, C.c_int(1 if on else 0))

    def render_frame(self, pc, accum: np.ndarray, launch=(0, 0), raw_L: np.ndarray | None = None):
        """one launch; accum (H,W,4 float32) is updated in place (prev == cur, pixels are independent)"""
        pcb = np.ascontiguousarray(pc)
        self.L.or_render_frame(self.h, _p(pcb), C.c_uint32(launch[0]), C.c_uint32(launch[1]), _p(accum), _p(accum), _p(self.counters), _p(raw_L))

    def render(self, n_launches: int, **kw) -> np.ndarray:
        """Renderer::render loop: clear, then launches num_frames = 0 .. n_launches-1 (frame 0 is discarded
        by the reference blend)."""
        s = self.scene
        accum = np.zeros((s.height, s.width, 4), np.float32)
        accum[..., 3] = 1.0
        for f in range(n_launches):
            self.render_frame(s.push_constants(f, **kw), accum)
        return accum

    def trace_primary_ids(self, pc):
        s = self.scene
        n = s.width * s.height
        inst, geom, prim = (np.zeros(n, np.uint32) for _ in range(3))
        t, u, v = (np.zeros(n, np.float32) for _ in range(3))
        pcb = np.ascontiguousarray(pc)
        self.L.or_trace_primary_ids(self.h, _p(pcb), _p(inst), _p(geom), _p(prim), _p(t), _p(u), _p(v))
        return inst, geom, prim, t, u, v

    def output_buffer(self, pc, which: int) -> np.ndarray:
        s = self.scene
        out = np.zeros((s.height, s.width, 4), np.float32)
        pcb = np.ascontiguousarray(pc)
        self.L.or_output_buffer(self.h, _p(pcb), C.c_int(which), _p(out))
        return out

    def gather_debug_rays(self, pc, num_debug_rays: int, max_vertices: int = 2048):
        out = np.zeros((max_vertices, 8), np.float32)
        pcb = np.ascontiguousarray(pc)
        self.L.or_gather_debug_rays.restype = C.c_uint32
        n = self.L.or_gather_debug_rays(self.h, _p(pcb), C.c_uint32(num_debug_rays), _p(out), C.c_uint32(max_vertices))
        return out[: min(n, max_vertices)], n

    def trace_rays(self, rays: np.ndarray, flags: int = 0):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        hits = np.zeros((len(rays), 6), np.float32)
        self.L.or_trace_rays(self.h, _p(rays), C.c_uint32(len(rays)), C.c_uint32(flags), _p(hits))
        return hits


class GlslRefScene(OracleScene):
    """The same scene rendered by the REFERENCE'S OWN SHADERS (oracle/_ref/libhelios_glsl_ref.so): render_frame runs
    path_trace_rgen/rchit/rahit/rmiss.glsl + path_trace_shadow.* as compiled from /root/reference; `restated_frame`
    runs the restatement inside the same library on the same scene object, for side-by-side checks."""

    def __init__(self, scene, **kw):
        L = ref_lib()
        if L is None:
            raise RuntimeError("oracle/_ref/libhelios_glsl_ref.so is not available (no /root/reference and no prebuilt copy)")
        super().__init__(scene, library=L, **kw)
        self.restated_counters = np.zeros(2, np.uint64)

    def restated_frame(self, pc, accum, launch=(0, 0)):
        pcb = np.ascontiguousarray(pc)
        self.L.or_render_frame(self.h, _p(pcb), C.c_uint32(launch[0]), C.c_uint32(launch[1]), _p(accum), _p(accum), _p(self.restated_counters), None)

    def render_frame(self, pc, accum: np.ndarray, launch=(0, 0), raw_L=None):
        pcb = np.ascontiguousarray(pc)
        self.L.ref_render_frame(self.h, _p(pcb), C.c_uint32(launch[0]), C.c_uint32(launch[1]), _p(accum), _p(accum), _p(self.counters))


_REF_DEBUG_PATH = _HERE / "_ref" / "libhelios_glsl_ref_raydebug.so"
_ref_debug_lib = None


def ref_debug_lib():
    """ctypes handle of the RAY_DEBUG_VIEW build of the reference's shaders (oracle/_ref/libhelios_glsl_ref_raydebug.so:
    the same sources compiled with -DRAY_DEBUG_VIEW, as path_integrator.cpp:259-307 builds its second pipeline), or None"""
    global _ref_debug_lib
    if _ref_debug_lib is None:
        build_ref()
        if not _REF_DEBUG_PATH.exists():
            return None
        _ref_debug_lib = C.CDLL(str(_REF_DEBUG_PATH))
        _ref_debug_lib.or_scene_new.restype = C.c_void_p
        _ref_debug_lib.or_scene_add_mesh.restype = C.c_int
        _ref_debug_lib.or_scene_add_texture.restype = C.c_int
        _ref_debug_lib.ref_gather_debug_rays.restype = C.c_uint32
        _ref_debug_lib.or_gather_debug_rays.restype = C.c_uint32
    return _ref_debug_lib


class GlslRefDebugScene(OracleScene):
    """The scene under the reference's RAY_DEBUG_VIEW pipeline: gather_debug_rays runs the reference's own shader files
    (compiled with that define); `restated_debug_rays` runs the restatement inside the same library."""

    def __init__(self, scene, **kw):
        L = ref_debug_lib()
        if L is None:
            raise RuntimeError("oracle/_ref/libhelios_glsl_ref_raydebug.so is not available (no /root/reference and no prebuilt copy)")
        super().__init__(scene, library=L, **kw)

    def gather_debug_rays(self, pc, num_debug_rays: int, max_vertices: int = 2048):
        out = np.zeros((max_vertices, 8), np.float32)
        pcb = np.ascontiguousarray(pc)
        n = self.L.ref_gather_debug_rays(self.h, _p(pcb), C.c_uint32(num_debug_rays), _p(out), C.c_uint32(max_vertices))
        return out[: min(n, max_vertices)], n

    def restated_debug_rays(self, pc, num_debug_rays: int, max_vertices: int = 2048):
        return OracleScene.gather_debug_rays(self, pc, num_debug_rays, max_vertices)


def ref_tonemap(accum: np.ndarray, exposure=1.0, op=0) -> np.ndarray:
    H, W = accum.shape[:2]
    out = np.zeros((H, W, 4), np.uint8)
    a = np.ascontiguousarray(accum, np.float32)
    ref_lib().ref_tonemap(_p(a), C.c_uint32(W), C.c_uint32(H), C.c_float(exposure), C.c_int(op), _p(out))
    return out


def ref_sky_bake(coeffs, sun_direction, size=64) -> np.ndarray:
    out = np.zeros((6, size, size, 4), np.float32)
    cf = np.ascontiguousarray(coeffs, np.float32)
    d = np.ascontiguousarray(sun_direction, np.float32)
    ref_lib().ref_sky_bake(_p(cf), _p(d), C.c_uint32(size), _p(out))
    return out


def tonemap(accum: np.ndarray, exposure=1.0, op=0, sample_scale=1.0) -> np.ndarray:
    H, W = accum.shape[:2]
    out = np.zeros((H, W, 4), np.uint8)
    a = np.ascontiguousarray(accum, np.float32)
    lib().or_tonemap(_p(a), C.c_uint32(W), C.c_uint32(H), C.c_float(exposure), C.c_int(op), C.c_float(sample_scale), _p(out))
    return out
