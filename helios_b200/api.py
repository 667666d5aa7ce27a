"""Thin Python driver over the C ABI — test / bench orchestration only (the host layer proper is the C++
shim in helios_b200/host, mirroring the reference's Scene / Renderer / PathIntegrator classes)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi
from ._lib import HeliosError, load


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Context:
    """hl_context: one GPU, one stream, one W x H accumulation image."""

    def __init__(self, width: int, height: int, device: int = 0):
        self.lib = load()
        self.h = C.c_void_p()
        st = self.lib.hl_context_create(C.c_int(device), C.c_uint32(width), C.c_uint32(height), C.byref(self.h))
        if st != 0:
            raise HeliosError(st, self.lib.hl_last_error(None).decode())
        self.width, self.height = width, height
        self.meshes = []
        self._keep = []

    def resize(self, width: int, height: int):
        """Renderer::on_window_resize: new extent, accumulation cleared; meshes, textures and tables stay"""
        self._chk(self.lib.hl_context_resize(self.h, C.c_uint32(width), C.c_uint32(height)))
        self.width, self.height = width, height

    def close(self):
        if self.h:
            self.lib.hl_context_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, st):
        if st != 0:
            raise HeliosError(st, self.lib.hl_last_error(self.h).decode())

    # ---- resources
    def create_mesh(self, vertices, indices, submeshes):
        v = np.ascontiguousarray(vertices, abi.VERTEX)
        i = np.ascontiguousarray(indices, np.uint32)
        s = np.ascontiguousarray(submeshes, abi.SUBMESH)
        m = C.c_void_p()
        self._chk(self.lib.hl_mesh_create(self.h, _p(v), C.c_uint32(len(v)), _p(i), C.c_uint32(len(i)), _p(s), C.c_uint32(len(s)), C.byref(m)))
        self.meshes.append(m)
        return m

    def destroy_mesh(self, mesh):
        self._chk(self.lib.hl_mesh_destroy(self.h, mesh))
        self.meshes = [m for m in self.meshes if m.value != mesh.value]

    def mesh_build_stats(self, mesh) -> np.ndarray:
        out = np.zeros((), abi.BUILD_STATS)
        self._chk(self.lib.hl_mesh_build_stats(self.h, mesh, _p(out)))
        return out

    def create_texture(self, fmt, w, h, data) -> int:
        d = np.ascontiguousarray(data)
        idx = C.c_int32()
        self._chk(self.lib.hl_texture2d_create(self.h, C.c_int(fmt), C.c_uint32(w), C.c_uint32(h), _p(d), C.byref(idx)))
        return idx.value

    def set_envmap(self, size, faces):
        f = np.ascontiguousarray(faces, np.float32) if size else None
        self._chk(self.lib.hl_envmap_set(self.h, C.c_uint32(size), _p(f)))

    def sky_update(self, coeffs40, sun_direction):
        cf = np.ascontiguousarray(coeffs40, np.float32)
        d = np.ascontiguousarray(sun_direction, np.float32)
        self._chk(self.lib.hl_sky_update(self.h, _p(cf), _p(d)))

    def read_envmap(self):
        size = C.c_uint32()
        self._chk(self.lib.hl_envmap_read(self.h, None, C.byref(size)))
        out = np.zeros((6, size.value, size.value, 4), np.float32)
        if size.value:
            self._chk(self.lib.hl_envmap_read(self.h, _p(out), C.byref(size)))
        return out

    def set_tables(self, materials, instances, meshes, submesh_info, lights):
        mats = np.ascontiguousarray(materials, abi.MATERIAL)
        inst = np.ascontiguousarray(instances, abi.INSTANCE)
        lts = np.ascontiguousarray(lights, abi.LIGHT)
        tabs = [np.ascontiguousarray(t, np.uint32) for t in submesh_info]
        ptrs = (C.c_void_p * max(len(tabs), 1))(*[t.ctypes.data for t in tabs])
        mh = (C.c_void_p * max(len(meshes), 1))(*[m.value for m in meshes])
        self._chk(self.lib.hl_scene_set_tables(self.h, _p(mats), C.c_uint32(len(mats)), _p(inst), mh, ptrs, C.c_uint32(len(inst)), _p(lts), C.c_uint32(len(lts))))

    def update_instances(self, instances):
        """new transforms for the installed instances: instance-tree REFIT instead of a rebuild (hl_scene_update_instances)"""
        inst = np.ascontiguousarray(instances, abi.INSTANCE)
        self._chk(self.lib.hl_scene_update_instances(self.h, _p(inst), C.c_uint32(len(inst))))

    def load_scene(self, scene, sky_coeffs=None):
        """uploads a helios_b200.scenes.SceneData: meshes (+BLAS build), textures, environment, tables (+TLAS)"""
        handles = [self.create_mesh(m.vertices, m.indices, m.submeshes) for m in scene.meshes]
        for fmt, w, h, data in scene.textures:
            self.create_texture(fmt, w, h, data)
        if scene.env_cube is not None:
            self.set_envmap(scene.env_cube[0], scene.env_cube[1])
        elif scene.sun_direction is not None:
            if sky_coeffs is None:
                from .sky import sky_coefficients

                sky_coeffs = sky_coefficients(scene.sun_direction)
            self.sky_update(sky_coeffs, scene.sun_direction)
        self.set_tables(scene.materials, scene.instances, [handles[int(i["mesh_index"])] for i in scene.instances], scene.submesh_info, scene.lights)
        return handles

    # ---- hot path
    def render_frame(self, pc, launch=(0, 0)):
        pcb = np.ascontiguousarray(pc, abi.PUSH_CONSTANTS)
        self._chk(self.lib.hl_render_frame(self.h, _p(pcb), C.c_uint32(launch[0]), C.c_uint32(launch[1])))

    def render_frame_tonemapped(self, pc, exposure=1.0, op=abi.TONE_MAP_ACES, launch=(0, 0)):
        """frame + fused accumulate / tone-map resolve pass (Renderer::render in one call)"""
        pcb = np.ascontiguousarray(pc, abi.PUSH_CONSTANTS)
        self._chk(self.lib.hl_render_frame_tonemapped(self.h, _p(pcb), C.c_uint32(launch[0]), C.c_uint32(launch[1]), C.c_float(exposure), C.c_int(op)))

    def render_frame_readback(self, pc, host_rgba8, exposure=1.0, op=abi.TONE_MAP_ACES, launch=(0, 0)):
        """frame + fused resolve + asynchronous read-back of the RGBA8 image into `host_rgba8` (pinned uint8 [H, W, 4]);
        complete after synchronize()"""
        pcb = np.ascontiguousarray(pc, abi.PUSH_CONSTANTS)
        assert host_rgba8.dtype == np.uint8 and host_rgba8.size == self.width * self.height * 4 and host_rgba8.flags["C_CONTIGUOUS"]
        self._chk(self.lib.hl_render_frame_readback(self.h, _p(pcb), C.c_uint32(launch[0]), C.c_uint32(launch[1]), C.c_float(exposure), C.c_int(op), _p(host_rgba8)))

    def read_rgba8(self, out=None):
        out = np.zeros((self.height, self.width, 4), np.uint8) if out is None else out
        self._chk(self.lib.hl_read_rgba8(self.h, _p(out)))
        return out

    def accum_clear(self):
        self._chk(self.lib.hl_accum_clear(self.h))

    def set_accum_mode(self, mode):
        self._chk(self.lib.hl_set_accum_mode(self.h, C.c_int(mode)))

    def trace_primary_ids(self, pc):
        n = self.width * self.height
        inst, geom, prim = (np.zeros(n, np.uint32) for _ in range(3))
        t, u, v = (np.zeros(n, np.float32) for _ in range(3))
        pcb = np.ascontiguousarray(pc, abi.PUSH_CONSTANTS)
        self._chk(self.lib.hl_trace_primary_ids(self.h, _p(pcb), _p(inst), _p(geom), _p(prim), _p(t), _p(u), _p(v)))
        return inst, geom, prim, t, u, v

    def trace_primary_device_only(self, pc):
        """primary-ray closest hits without the read-back (timing: bracket with event_record)"""
        pcb = np.ascontiguousarray(pc, abi.PUSH_CONSTANTS)
        self._chk(self.lib.hl_trace_primary_ids(self.h, _p(pcb), None, None, None, None, None, None))

    def render_output_buffer(self, pc, output_buffer: int) -> np.ndarray:
        """debug output buffer (abi.OUTPUT_BUFFER_*) of the surfaces seen by the primary rays of `pc`: [H, W, 4] float32"""
        pcb = np.ascontiguousarray(pc, abi.PUSH_CONSTANTS)
        out = np.zeros((self.height, self.width, 4), np.float32)
        self._chk(self.lib.hl_render_output_buffer(self.h, _p(pcb), C.c_int(output_buffer), _p(out)))
        return out

    def gather_debug_rays(self, pc, num_debug_rays: int, max_vertices: int = 2 * abi.MAX_DEBUG_RAY_DRAW_COUNT):
        """ray debug view (PathIntegrator::gather_debug_rays): (vertices [min(count, max_vertices)] of abi.DEBUG_RAY_VERTEX,
        count) — two vertices per secondary-ray segment of num_debug_rays paths through pc.ray_debug_pixel_coord"""
        pcb = np.ascontiguousarray(pc, abi.PUSH_CONSTANTS)
        out = np.zeros(max_vertices, abi.DEBUG_RAY_VERTEX)
        n = C.c_uint32(0)
        self._chk(self.lib.hl_gather_debug_rays(self.h, _p(pcb), C.c_uint32(num_debug_rays), _p(out), C.c_uint32(max_vertices), C.byref(n)))
        return out[: min(n.value, max_vertices)], n.value

    def trace_rays(self, rays, flags=0):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        hits = np.zeros((len(rays), 6), np.float32)
        self._chk(self.lib.hl_trace_rays(self.h, _p(rays), C.c_uint32(len(rays)), C.c_uint32(flags), _p(hits)))
        return hits

    def tonemap(self, exposure=1.0, op=abi.TONE_MAP_ACES, sample_scale=1.0, download=True, out=None):
        if out is None:
            out = np.zeros((self.height, self.width, 4), np.uint8) if download else None
        else:
            assert out.dtype == np.uint8 and out.size == self.width * self.height * 4 and out.flags["C_CONTIGUOUS"]
        self._chk(self.lib.hl_tonemap(self.h, C.c_float(exposure), C.c_int(op), C.c_float(sample_scale), _p(out)))
        return out

    def read_accum(self):
        out = np.zeros((self.height, self.width, 4), np.float32)
        self._chk(self.lib.hl_read_accum(self.h, _p(out)))
        return out

    def write_accum(self, a):
        a = np.ascontiguousarray(a, np.float32)
        assert a.shape == (self.height, self.width, 4)
        self._chk(self.lib.hl_write_accum(self.h, _p(a)))

    def accum_device_ptr(self) -> int:
        p = C.c_void_p()
        self._chk(self.lib.hl_accum_device_ptr(self.h, C.byref(p)))
        return p.value

    def synchronize(self):
        self._chk(self.lib.hl_synchronize(self.h))

    def counters(self) -> np.ndarray:
        out = np.zeros((), abi.COUNTERS)
        self._chk(self.lib.hl_get_counters(self.h, _p(out)))
        return out

    def bounce_profile(self):
        """per-bounce breakdown of the last frame rendered under set_profiling(True): (array of abi.BOUNCE_PROFILE, rays traced by
        the tail kernel: extension, shadow)"""
        out = np.zeros(64, abi.BOUNCE_PROFILE)
        n, te, ts = C.c_uint32(), C.c_uint64(), C.c_uint64()
        self._chk(self.lib.hl_get_bounce_profile(self.h, _p(out), C.c_uint32(64), C.byref(n), C.byref(te), C.byref(ts)))
        return out[: n.value], te.value, ts.value

    def reset_counters(self):
        self._chk(self.lib.hl_reset_counters(self.h))

    def set_profiling(self, on: bool):
        self._chk(self.lib.hl_set_profiling(self.h, C.c_int(1 if on else 0)))

    def kernel_launches(self) -> int:
        n = C.c_uint64()
        self._chk(self.lib.hl_kernel_launches(self.h, C.byref(n)))
        return n.value

    def set_option(self, option: int, value: int):
        self._chk(self.lib.hl_set_option(self.h, C.c_int(option), C.c_int64(value)))

    def event_record(self, slot: int):
        self._chk(self.lib.hl_event_record(self.h, C.c_int(slot)))

    def event_elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_float()
        self._chk(self.lib.hl_event_elapsed_ms(self.h, C.c_int(a), C.c_int(b), C.byref(ms)))
        return ms.value

    # ---- multi-GPU, one process per GPU (hl_comm_init_rank + NCCL on the context's stream)
    def comm_init_rank(self, unique_id: bytes, n_ranks: int, rank: int):
        assert len(unique_id) == abi.COMM_ID_BYTES
        buf = (C.c_uint8 * abi.COMM_ID_BYTES).from_buffer_copy(unique_id)
        self._chk(self.lib.hl_comm_init_rank(self.h, buf, C.c_int(n_ranks), C.c_int(rank)))

    def comm_destroy(self):
        self._chk(self.lib.hl_comm_destroy(self.h))

    def accum_all_reduce(self):
        """accumulation image <- sum over ranks (asynchronous, behind the frames in flight)"""
        self._chk(self.lib.hl_accum_all_reduce(self.h))

    def accum_reduce(self, root: int = 0):
        self._chk(self.lib.hl_accum_reduce(self.h, C.c_int(root)))

    def render(self, scene, n_launches, **kw):
        """Renderer::render loop: clear, then launches with num_frames = 0 .. n_launches-1"""
        self.accum_clear()
        for f in range(n_launches):
            self.render_frame(scene.push_constants(f, **kw))
        return self.read_accum()


def comm_unique_id() -> bytes:
    """rank 0: the NCCL unique id every rank passes to Context.comm_init_rank"""
    lib = load()
    buf = (C.c_uint8 * abi.COMM_ID_BYTES)()
    st = lib.hl_comm_unique_id(buf)
    if st != 0:
        raise HeliosError(st, lib.hl_comm_last_error().decode())
    return bytes(buf)


class Group:
    """all GPUs in one process (hl_comm_init_all): contexts[i] is rank i"""

    def __init__(self, contexts):
        self.ctxs = list(contexts)
        self.lib = load()
        self._arr = (C.c_void_p * len(self.ctxs))(*[c.h.value for c in self.ctxs])
        self._chk(self.lib.hl_comm_init_all(self._arr, C.c_int(len(self.ctxs))))

    def _chk(self, st):
        if st != 0:
            raise HeliosError(st, self.lib.hl_comm_last_error().decode())

    def reduce(self, root: int = 0):
        self._chk(self.lib.hl_multi_gpu_reduce(self._arr, C.c_int(len(self.ctxs)), C.c_int(root)))

    def resolve(self, root=0, exposure=1.0, op=abi.TONE_MAP_ACES, sample_scale=1.0, download=True):
        c = self.ctxs[root]
        out = np.zeros((c.height, c.width, 4), np.uint8) if download else None
        self._chk(self.lib.hl_multi_gpu_resolve(self._arr, C.c_int(len(self.ctxs)), C.c_int(root), C.c_float(exposure), C.c_int(op), C.c_float(sample_scale), _p(out)))
        return out
