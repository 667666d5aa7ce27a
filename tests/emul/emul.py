"""ctypes wrapper for the kernel-logic emulator (tests/emul/emul.cpp) — a debug harness for the GPU-less
build container; never imported by the helios_b200 package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_EXTRA = os.environ.get("HL_EMUL_CXXFLAGS", "").split()  # tuning experiments (tools/bvh_stats.py)
_SO = _HERE / "_build" / ("libhl_emul.so" if not _EXTRA else "libhl_emul_" + "_".join(x.strip("-").replace("=", "") for x in _EXTRA) + ".so")
_lib = None


def build(force=False):
    srcs = [_HERE / "emul.cpp"] + list((_HERE.parent.parent / "helios_b200" / "csrc").glob("*.h")) + [_HERE.parent.parent / "include" / "helios_b200.h"]
    if force or not _SO.exists() or any(s.stat().st_mtime > _SO.stat().st_mtime for s in srcs):
        _SO.parent.mkdir(exist_ok=True)
        cxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
        subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-fvisibility=hidden",
                               "-Wno-unknown-pragmas", *_EXTRA, "-o", str(_SO), str(_HERE / "emul.cpp")])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_SO))
        _lib.em_scene_new.restype = C.c_void_p
        _lib.em_scene_add_mesh.restype = C.c_int
        _lib.em_scene_add_texture.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class EmulScene:
    def __init__(self, scene, sky_faces=None, sky_size=512, force_two_level=False):
        L = lib()
        self.scene = scene
        self.h = C.c_void_p(L.em_scene_new())
        self._keep = []
        for m in scene.meshes:
            v, i, s = np.ascontiguousarray(m.vertices), np.ascontiguousarray(m.indices), np.ascontiguousarray(m.submeshes)
            L.em_scene_add_mesh(self.h, _p(v), C.c_uint32(len(v)), _p(i), C.c_uint32(len(i)), _p(s), C.c_uint32(len(s)))
        for fmt, w, h, data in scene.textures:
            d = np.ascontiguousarray(data)
            L.em_scene_add_texture(self.h, C.c_int(fmt), C.c_uint32(w), C.c_uint32(h), _p(d))
        if scene.env_cube is not None:
            size, faces = scene.env_cube
            f = np.ascontiguousarray(faces, np.float32)
            L.em_scene_set_envmap(self.h, C.c_uint32(size), _p(f))
        elif sky_faces is not None:
            f = np.ascontiguousarray(sky_faces, np.float32)
            L.em_scene_set_envmap(self.h, C.c_uint32(f.shape[1]), _p(f))
        mats = np.ascontiguousarray(scene.materials)
        inst = np.ascontiguousarray(scene.instances)
        lights = np.ascontiguousarray(scene.lights)
        tabs = [np.ascontiguousarray(t, np.uint32) for t in scene.submesh_info]
        ptrs = (C.c_void_p * len(tabs))(*[t.ctypes.data for t in tabs])
        self._keep += tabs
        L.em_scene_set_tables(self.h, _p(mats), C.c_uint32(len(mats)), _p(inst), ptrs, C.c_uint32(len(inst)), _p(lights), C.c_uint32(len(lights)))
        if force_two_level:
            L.em_scene_force_two_level(self.h)
        self.counters = np.zeros(2, np.uint64)

    def __del__(self):
        try:
            lib().em_scene_free(self.h)
        except Exception:
            pass

    def update_instances(self, instances):
        """hl_scene_update_instances on the emulator: new transforms, instance tree refitted (not rebuilt)"""
        inst = np.ascontiguousarray(instances)
        assert lib().em_scene_update_instances(self.h, _p(inst), C.c_uint32(len(inst))) == 0

    def render_frame(self, pc, accum, launch=(0, 0), accum_mode=0):
        pcb = np.ascontiguousarray(pc)
        lib().em_render_frame(self.h, _p(pcb), C.c_uint32(launch[0]), C.c_uint32(launch[1]), _p(accum), _p(self.counters), C.c_int(accum_mode))

    def render(self, n_launches, **kw):
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
        lib().em_trace_primary_ids(self.h, _p(pcb), _p(inst), _p(geom), _p(prim), _p(t), _p(u), _p(v))
        return inst, geom, prim, t, u, v

    def output_buffer(self, pc, which):
        s = self.scene
        out = np.zeros((s.height, s.width, 4), np.float32)
        pcb = np.ascontiguousarray(pc)
        lib().em_output_buffer(self.h, _p(pcb), C.c_int(which), _p(out))
        return out

    def gather_debug_rays(self, pc, num_debug_rays, max_vertices=2048):
        out = np.zeros((max_vertices, 8), np.float32)
        pcb = np.ascontiguousarray(pc)
        lib().em_gather_debug_rays.restype = C.c_uint32
        n = lib().em_gather_debug_rays(self.h, _p(pcb), C.c_uint32(num_debug_rays), _p(out), C.c_uint32(max_vertices))
        return out[: min(n, max_vertices)], n

    def trace_rays(self, rays, flags=0):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        hits = np.zeros((len(rays), 6), np.float32)
        lib().em_trace_rays(self.h, _p(rays), C.c_uint32(len(rays)), C.c_uint32(flags), _p(hits))
        return hits

    def traversal_stats(self, rays, flags=0):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        out = np.zeros(3, np.uint64)
        lib().em_traversal_stats(self.h, _p(rays), C.c_uint32(len(rays)), C.c_uint32(flags), _p(out))
        return {"nodes_per_ray": out[0] / len(rays), "leaf_tests_per_ray": out[1] / len(rays), "max_nodes": int(out[2])}

    def mesh_stats(self, mesh=0):
        out = np.zeros(3, np.uint32)
        lib().em_mesh_stats(self.h, C.c_int(mesh), _p(out))
        return out


def sky_bake(coeffs, sun, size=512):
    out = np.zeros((6, size, size, 4), np.float32)
    cf = np.ascontiguousarray(coeffs, np.float32)
    d = np.ascontiguousarray(sun, np.float32)
    lib().em_sky_bake(_p(cf), _p(d), C.c_uint32(size), _p(out))
    return out
