"""N > 1 host logic on CPU: world_size 2 over gloo.  Each rank renders its frame indices into a SUM buffer (here
with the kernel-logic emulator standing in for the GPU), one all-reduce combines them; the result must equal the
single-process sum over the same frames and, divided by the sample count, the reference's running mean."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, frames_per_rank, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    from helios_b200 import abi, multi_gpu, scenes
    from tests.emul import emul

    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = scenes.cornell_box(24, 24)
    e = emul.EmulScene(s)
    acc = np.zeros((s.height, s.width, 4), np.float32)
    mine = multi_gpu.frame_indices(rank, world, frames_per_rank)
    for f in mine:
        e.render_frame(s.push_constants(f), acc, accum_mode=abi.ACCUM_SUM)
    t = torch.from_numpy(acc)
    multi_gpu.all_reduce_sum(t, dist)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        np.save(os.path.join(out_dir, "sum.npy"), t.numpy())
        np.save(os.path.join(out_dir, "frames.npy"), np.array(sorted(sum(gathered, []))))
    dist.barrier()
    dist.destroy_process_group()


def test_spp_sharding_world_size_2(tmp_path):
    import torch.multiprocessing as mp

    from helios_b200 import abi, multi_gpu, scenes
    from tests.emul import emul

    world, fpr = 2, 3
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, fpr, str(tmp_path)), nprocs=world, join=True)
    total = np.load(tmp_path / "sum.npy")
    frames = np.load(tmp_path / "frames.npy").tolist()
    assert frames == list(range(1, world * fpr + 1))  # disjoint cover of frames 1..K*G, frame 0 skipped
    s = scenes.cornell_box(24, 24)
    e = emul.EmulScene(s)
    ref = np.zeros((s.height, s.width, 4), np.float32)
    for f in frames:
        e.render_frame(s.push_constants(f), ref, accum_mode=abi.ACCUM_SUM)
    assert np.allclose(total[..., :3], ref[..., :3], atol=1e-5)  # fp32 sum order differs between 1 and 2 ranks
    mean = e.render(world * fpr + 1)  # the reference's running mean over launches 0..K*G
    assert np.allclose(total[..., :3] / multi_gpu.total_samples(world, fpr), mean[..., :3], atol=2e-6)


def test_frame_indices_partition():
    from helios_b200 import multi_gpu

    for world in (1, 2, 4, 8):
        allf = sorted(sum((multi_gpu.frame_indices(r, world, 5) for r in range(world)), []))
        assert allf == list(range(1, 5 * world + 1))
