"""Multi-GPU path on hardware: samples-per-pixel sharding into per-rank SUM images + ONE reduction issued by
the library (helios_b200/csrc/hl_comm.cu), against one context rendering the same frame indices.

  * hl_comm_init_all + hl_multi_gpu_reduce / hl_multi_gpu_resolve over peer memory: runs on ONE GPU too (two contexts on
    device 0 are two ranks whose images live in the same memory) — the driver's single-GPU box exercises the kernel;
  * the same with one context per GPU, and NCCL (hl_comm_init_rank + hl_accum_all_reduce / hl_accum_reduce) with one
    process per GPU: skipped below 2 GPUs.
Blend semantics kept: path_trace_rgen.glsl:219-247 (sum / count = that running mean up to fp32 rounding); tone map
tone_map.frag:35-51.  Tolerance: the only difference between N ranks and one is the fp32 summation order of N K clamped
samples in [0, 1]: |sum_N - sum_1| <= N K * 2^-24 * max partial sum, stated as 1e-5 * frames below."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

from helios_b200 import abi, multi_gpu, scenes

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def n_gpus() -> int:
    import torch

    return torch.cuda.device_count()


@pytest.fixture(scope="module")
def api():
    from helios_b200 import api as _api

    return _api


def single_context_sum(api, s, frames):
    ctx = api.Context(s.width, s.height)
    ctx.load_scene(s)
    ctx.set_accum_mode(abi.ACCUM_SUM)
    ctx.accum_clear()
    for f in frames:
        ctx.render_frame(s.push_constants(f))
    acc = ctx.read_accum()
    img = ctx.tonemap(1.0, abi.TONE_MAP_ACES, sample_scale=1.0 / len(frames))
    ctx.close()
    return acc, img


def run_group(api, s, devices, fpr):
    """one context per entry of `devices`, rank-g frames into its SUM image; returns (contexts, group, all frame indices)"""
    world = len(devices)
    ctxs = []
    for d in devices:
        c = api.Context(s.width, s.height, device=d)
        c.load_scene(s)
        c.set_accum_mode(abi.ACCUM_SUM)
        c.accum_clear()
        ctxs.append(c)
    g = api.Group(ctxs)
    for r, c in enumerate(ctxs):
        for f in multi_gpu.frame_indices(r, world, fpr):
            c.render_frame(s.push_constants(f))
    return ctxs, g, list(range(1, world * fpr + 1))


@pytest.mark.parametrize("world,root", [(2, 0), (3, 2)])
def test_peer_reduce_two_contexts_one_gpu(api, world, root):
    s = scenes.cornell_box(160, 96)
    fpr = 4
    ctxs, g, frames = run_group(api, s, [0] * world, fpr)
    own = [c.read_accum() for c in ctxs]
    g.reduce(root)
    total = ctxs[root].read_accum()
    # rank-ordered fp32 sum of the per-rank images: reproducible bit for bit on the host
    expect = own[0][..., :3].copy()
    for r in range(1, world):
        expect = expect + own[r][..., :3]
    assert np.array_equal(total[..., :3].view(np.uint32), expect.view(np.uint32))
    assert np.all(total[..., 3] == 1.0)
    for r, c in enumerate(ctxs):  # the other ranks keep their own image
        if r != root:
            assert np.array_equal(c.read_accum(), own[r])
    ref, ref_img = single_context_sum(api, s, frames)
    assert np.abs(total[..., :3] - ref[..., :3]).max() <= 1e-5 * len(frames)
    for c in ctxs:
        c.close()


def test_peer_resolve_matches_single_gpu_image(api):
    s = scenes.cornell_box(128, 128)
    ctxs, g, frames = run_group(api, s, [0, 0], 6)
    img = g.resolve(root=0, exposure=1.0, op=abi.TONE_MAP_ACES, sample_scale=1.0 / len(frames))
    ref, ref_img = single_context_sum(api, s, frames)
    d = np.abs(img.astype(np.int32) - ref_img.astype(np.int32))
    assert d.max() <= 1 and (d > 0).mean() < 0.01  # 8-bit rounding of values that differ by fp32 summation order
    # the fused kernel's image = the root's own tone-map pass over the reduced sum, bit for bit
    again = ctxs[0].tonemap(1.0, abi.TONE_MAP_ACES, sample_scale=1.0 / len(frames))
    assert np.array_equal(img, again)
    assert np.array_equal(ctxs[0].read_rgba8(), again)
    # Reinhard + device-only variant
    g.resolve(root=0, exposure=0.7, op=abi.TONE_MAP_REINHARD, sample_scale=1.0 / (2 * len(frames)), download=False)
    for c in ctxs:
        c.close()


def test_sum_mode_fused_preview(api):
    """hl_render_frame_readback in HL_ACCUM_SUM mode: the image of step k is sum / k (the rank-local progressive preview)"""
    s = scenes.cornell_box(96, 96)
    ctx = api.Context(s.width, s.height)
    ctx.load_scene(s)
    ctx.set_accum_mode(abi.ACCUM_SUM)
    ctx.accum_clear()
    host = [np.zeros((s.height, s.width, 4), np.uint8) for _ in range(5)]
    for k in range(5):
        ctx.render_frame_readback(s.push_constants(1 + k), host[k])
    ctx.synchronize()
    assert np.array_equal(host[4], ctx.tonemap(1.0, abi.TONE_MAP_ACES, sample_scale=1.0 / 5))
    ctx.accum_clear()
    ctx.render_frame_readback(s.push_constants(1), host[0])
    ctx.synchronize()
    assert np.array_equal(host[0], ctx.tonemap(1.0, abi.TONE_MAP_ACES, sample_scale=1.0))
    from helios_b200._lib import HeliosError

    with pytest.raises(HeliosError):  # tiles have no per-pixel sample count in sum mode
        ctx.render_frame_readback(s.push_constants(2, tile=(0, 0)), host[0], launch=(32, 32))
    ctx.close()


def test_group_argument_errors(api):
    from helios_b200._lib import HeliosError

    a, b = api.Context(64, 64), api.Context(64, 32)
    with pytest.raises(HeliosError):
        api.Group([a, b])  # extents differ
    with pytest.raises(HeliosError):
        api.Group([a, a])  # the same context twice
    with pytest.raises(HeliosError):
        a.accum_all_reduce()  # no communicator
    a.close(), b.close()


@pytest.mark.skipif("n_gpus() < 2", reason="needs 2 GPUs")
def test_peer_and_nccl_reduce_one_context_per_gpu(api):
    s = scenes.cornell_box(192, 128)
    world = min(n_gpus(), 4)
    ctxs, g, frames = run_group(api, s, list(range(world)), 3)
    own = [c.read_accum() for c in ctxs]
    img = g.resolve(root=0, sample_scale=1.0 / len(frames))
    total = ctxs[0].read_accum()
    expect = own[0][..., :3].copy()
    for r in range(1, world):
        expect = expect + own[r][..., :3]
    assert np.array_equal(total[..., :3].view(np.uint32), expect.view(np.uint32))
    ref, ref_img = single_context_sum(api, s, frames)
    assert np.abs(total[..., :3] - ref[..., :3]).max() <= 1e-5 * len(frames)
    assert np.abs(img.astype(np.int32) - ref_img.astype(np.int32)).max() <= 1
    # the NCCL communicator hl_comm_init_all created beside the peer group: all-reduce on every rank
    for r, c in enumerate(ctxs):
        c.write_accum(own[r])
    import ctypes as C

    lib = ctxs[0].lib
    # single process: the per-rank calls must be grouped (ncclGroupStart/End lives behind hl_multi_gpu_reduce's NCCL branch);
    # here each rank's all-reduce is issued from its own thread, as a multi-threaded host would
    import threading

    errs = []

    def one(c):
        try:
            c.accum_all_reduce()
            c.synchronize()
        except Exception as e:  # pragma: no cover
            errs.append(e)

    th = [threading.Thread(target=one, args=(c,)) for c in ctxs]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for c in ctxs:
        assert np.abs(c.read_accum()[..., :3] - expect).max() <= 1e-5 * len(frames)
    for c in ctxs:
        c.close()


def _nccl_worker(rank, world, port, out_dir, fpr):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist

    from helios_b200 import api

    dist.init_process_group("gloo", rank=rank, world_size=world)  # host plumbing only; the data path is the library's own NCCL communicator
    obj = [api.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, 0)
    s = scenes.cornell_box(160, 120)
    ctx = api.Context(s.width, s.height, device=rank)
    ctx.load_scene(s)
    ctx.set_accum_mode(abi.ACCUM_SUM)
    ctx.accum_clear()
    ctx.comm_init_rank(obj[0], world, rank)
    for f in multi_gpu.frame_indices(rank, world, fpr):
        ctx.render_frame(s.push_constants(f))
    own = ctx.read_accum()
    ctx.accum_reduce(0)
    red = ctx.read_accum()
    ctx.write_accum(own)
    ctx.accum_all_reduce()
    allred = ctx.read_accum()
    np.save(os.path.join(out_dir, f"own{rank}.npy"), own)
    np.save(os.path.join(out_dir, f"red{rank}.npy"), red)
    np.save(os.path.join(out_dir, f"all{rank}.npy"), allred)
    dist.barrier()
    ctx.comm_destroy()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.skipif("n_gpus() < 2", reason="needs 2 GPUs")
def test_nccl_one_process_per_gpu(api, tmp_path):
    import torch.multiprocessing as mp

    world, fpr = 2, 3
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_nccl_worker, args=(world, port, str(tmp_path), fpr), nprocs=world, join=True)
    own = [np.load(tmp_path / f"own{r}.npy") for r in range(world)]
    expect = own[0][..., :3] + own[1][..., :3]
    assert np.abs(np.load(tmp_path / "red0.npy")[..., :3] - expect).max() <= 1e-6
    assert np.array_equal(np.load(tmp_path / "red1.npy"), own[1])  # ncclReduce leaves the non-root image alone
    for r in range(world):
        assert np.abs(np.load(tmp_path / f"all{r}.npy")[..., :3] - expect).max() <= 1e-6
    s = scenes.cornell_box(160, 120)
    ref, _ = single_context_sum(api, s, list(range(1, world * fpr + 1)))
    assert np.abs(expect - ref[..., :3]).max() <= 1e-5 * world * fpr
