"""Samples-per-pixel sharding across the GPUs of one box.

Samples are independent given (pixel, frame index) seeding (random.glsl:40-50, path_trace_rgen.glsl:191); the only
coupling in the reference is the running-mean blend (rgen:234-237).  Each rank therefore renders its own frame
indices into a per-GPU SUM buffer (HL_ACCUM_SUM), the scene and its BVH are replicated, and ONE all-reduce(sum) of
the W x H x 4 fp32 accumulation image combines them; dividing by the number of samples (fused into the tone-map
pass as sample_scale) gives the image.  Frame 0 is skipped: the reference's blend discards it.
No data-path collective other than that final reduce.
"""
from __future__ import annotations


def frame_indices(rank: int, world: int, frames_per_rank: int) -> list[int]:
    """round-robin: rank g renders num_frames = g+1, g+1+G, ... (1-based: frame 0 is discarded upstream)"""
    return [1 + rank + k * world for k in range(frames_per_rank)]


def total_samples(world: int, frames_per_rank: int) -> int:
    return world * frames_per_rank


def all_reduce_sum(tensor, dist):
    """the one collective of the path: NCCL (GPU) / gloo (CPU tests) sum of the accumulation image, in place"""
    dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor


class DeviceArray:
    """__cuda_array_interface__ view of the library's accumulation image (hl_accum_device_ptr) for torch"""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}
