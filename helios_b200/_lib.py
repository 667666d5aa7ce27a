"""ctypes binding of libhelios_b200.so (the C ABI in include/helios_b200.h).

Fails loudly: if the library is missing or cannot be loaded there is NO fallback — importing
helios_b200.api raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

# HELIOS_B200_LIB selects another build of the same library (tools/tune_trace.py compiles tuning variants)
LIB_PATH = Path(os.environ.get("HELIOS_B200_LIB") or Path(__file__).resolve().parent / "libhelios_b200.so")

# every symbol include/helios_b200.h declares
SYMBOLS = [
    "hl_context_create", "hl_context_destroy", "hl_context_resize", "hl_last_error", "hl_version",
    "hl_mesh_create", "hl_mesh_destroy", "hl_mesh_build_stats", "hl_texture2d_create", "hl_textures_clear",
    "hl_envmap_set", "hl_sky_update", "hl_envmap_read", "hl_scene_set_tables", "hl_scene_update_instances", "hl_render_frame", "hl_render_frame_tonemapped", "hl_render_frame_readback", "hl_read_rgba8", "hl_accum_clear",
    "hl_set_accum_mode", "hl_trace_primary_ids", "hl_render_output_buffer", "hl_gather_debug_rays", "hl_trace_rays", "hl_tonemap", "hl_read_accum", "hl_write_accum",
    "hl_accum_device_ptr", "hl_synchronize", "hl_get_counters", "hl_reset_counters", "hl_set_profiling", "hl_kernel_launches", "hl_event_record", "hl_event_elapsed_ms", "hl_set_option",
    "hl_get_bounce_profile", "hl_comm_unique_id", "hl_comm_init_rank", "hl_comm_init_all", "hl_comm_destroy", "hl_comm_last_error", "hl_accum_all_reduce", "hl_accum_reduce", "hl_multi_gpu_reduce", "hl_multi_gpu_resolve",
]

_lib = None


class HeliosError(RuntimeError):
    """raised for any non-zero hl_status (the reference throws std::runtime_error after HELIOS_LOG_FATAL)"""

    def __init__(self, status: int, message: str):
        super().__init__(f"helios_b200 error {status}: {message}")
        self.status = status


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m helios_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    for s in SYMBOLS:
        getattr(lib, s)  # AttributeError if the library does not export what the header declares
    lib.hl_last_error.restype = C.c_char_p
    lib.hl_last_error.argtypes = [C.c_void_p]
    lib.hl_version.restype = C.c_char_p
    lib.hl_comm_last_error.restype = C.c_char_p
    _lib = lib
    return lib
