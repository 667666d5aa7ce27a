"""numpy views of the shader-ABI structs declared in include/helios_b200.h.

Byte layouts follow the reference (paths relative to the reference checkout):
  Vertex        include/resource/mesh.h:10-17          80 B
  MaterialData  src/engine/resource/scene.cpp:25-32    80 B
  LightData     src/engine/resource/scene.cpp:36-42    64 B
  InstanceData  src/engine/resource/scene.cpp:46-52   144 B
  PushConstants src/engine/gfx/path_integrator.cpp:11-28  192 B
"""
import numpy as np

VERTEX = np.dtype(
    [("position", "<f4", 4), ("tex_coord", "<f4", 4), ("normal", "<f4", 4), ("tangent", "<f4", 4), ("bitangent", "<f4", 4)]
)
MATERIAL = np.dtype(
    [
        ("texture_indices0", "<i4", 4),
        ("texture_indices1", "<i4", 4),
        ("albedo", "<f4", 4),
        ("emissive", "<f4", 4),
        ("roughness_metallic", "<f4", 4),
    ]
)
LIGHT = np.dtype([("light_data0", "<f4", 4), ("light_data1", "<f4", 4), ("light_data2", "<f4", 4), ("light_data3", "<f4", 4)])
INSTANCE = np.dtype([("model_matrix", "<f4", 16), ("normal_matrix", "<f4", 16), ("mesh_index", "<u4"), ("padding", "<f4", 3)])
PUSH_CONSTANTS = np.dtype(
    [
        ("view_proj_inverse", "<f4", 16),
        ("camera_pos", "<f4", 4),
        ("up_direction", "<f4", 4),
        ("right_direction", "<f4", 4),
        ("focal_plane", "<f4", 4),
        ("ray_debug_pixel_coord", "<i4", 4),
        ("launch_id_size", "<u4", 4),
        ("accumulation", "<f4"),
        ("num_lights", "<u4"),
        ("num_frames", "<u4"),
        ("debug_vis", "<u4"),
        ("max_ray_bounces", "<u4"),
        ("shadow_ray_bias", "<f4"),
        ("focal_length", "<f4"),
        ("aperture_radius", "<f4"),
    ]
)
SUBMESH = np.dtype([("base_index", "<u4"), ("index_count", "<u4"), ("vertex_count", "<u4"), ("opaque", "<u4")])
COUNTERS = np.dtype(
    [
        ("extension_rays", "<u8"),
        ("shadow_rays", "<u8"),
        ("frames", "<u8"),
        ("ms_generate", "<f4"),
        ("ms_extend", "<f4"),
        ("ms_shade", "<f4"),
        ("ms_connect", "<f4"),
        ("ms_resolve", "<f4"),
        ("ms_frame", "<f4"),
    ],
    align=True,
)
BUILD_STATS = np.dtype(
    [
        ("triangles", "<u4"),
        ("wide_nodes", "<u4"),
        ("binary_nodes", "<u4"),
        ("ms_build", "<f4"),
        ("sah_cost", "<f4"),
        ("bytes_nodes", "<u8"),
        ("bytes_triangles", "<u8"),
    ],
    align=True,
)

assert VERTEX.itemsize == 80 and MATERIAL.itemsize == 80 and LIGHT.itemsize == 64
assert INSTANCE.itemsize == 144 and PUSH_CONSTANTS.itemsize == 192 and SUBMESH.itemsize == 16

LIGHT_DIRECTIONAL, LIGHT_SPOT, LIGHT_POINT, LIGHT_ENVIRONMENT_MAP, LIGHT_AREA = 0, 1, 2, 3, 4
TONE_MAP_ACES, TONE_MAP_REINHARD = 0, 1
TEX_RGBA8_UNORM, TEX_RGBA8_SRGB, TEX_RGBA8_SNORM, TEX_RGBA32F = 0, 1, 2, 3
OUTPUT_BUFFER_ALBEDO, OUTPUT_BUFFER_NORMALS, OUTPUT_BUFFER_ROUGHNESS, OUTPUT_BUFFER_METALLIC, OUTPUT_BUFFER_EMISSIVE = 0, 1, 2, 3, 4  # include/gfx/renderer.h:25-33
ACCUM_RUNNING_MEAN, ACCUM_SUM = 0, 1
OPT_TAIL_THRESHOLD, OPT_TAIL_START, OPT_PIPELINE, OPT_SAH_CLUSTER, OPT_FRAMES_IN_FLIGHT, OPT_CUDA_GRAPH = 1, 2, 3, 4, 5, 6  # hl_set_option
DEBUG_RAY_VERTEX = np.dtype([("position", "<f4", 4), ("color", "<f4", 4)])  # common.glsl:54-58
MAX_DEBUG_RAY_DRAW_COUNT = 1024  # include/gfx/renderer.h:9
MISS_ID = 0xFFFFFFFF
COMM_ID_BYTES = 128  # HL_COMM_ID_BYTES
BOUNCE_PROFILE = np.dtype([("extension_rays", "<u4"), ("shadow_rays", "<u4"), ("ms_tail", "<f4"), ("ms_extend", "<f4"), ("ms_shade", "<f4"), ("ms_connect", "<f4")])  # hl_bounce_profile
