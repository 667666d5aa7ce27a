"""SURVEY.md §8 f2 on the GPU: hl_scene_update_instances refits the instance tree (k_tlas_refit) instead of rebuilding it.
Reference: the TLAS is created ALLOW_UPDATE (src/engine/resource/scene.cpp:797) and rebuilt on every change
(src/engine/gfx/renderer.cpp:147-168).  Bar: after any sequence of moves, hit IDs and (t, u, v) of the refitted tree equal
those of a context built from scratch with the same transforms, bit for bit, and both equal the oracle's IDs."""
import copy
import time

import numpy as np
import pytest

from helios_b200 import scenes
from tests.test_emul_parity import moved_instances

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from helios_b200 import api as _api

    return _api


def test_refit_equals_rebuild_and_oracle(api, oracle_mod):
    s = scenes.city_scene(n_instances=120, n_meshes=6, width=256, height=144, floors=(2, 5), detail=(1, 3))
    ctx = api.Context(s.width, s.height)
    handles = ctx.load_scene(s)
    ctx.render(s, 2)  # frames in flight / CUDA graphs exist before the first update
    cur = s
    for rnd in range(4):
        inst = moved_instances(cur, seed=100 + rnd, frac=0.6, shift=8.0)
        ctx.update_instances(inst)
        cur = copy.copy(cur)
        cur.instances = inst
        fresh = api.Context(s.width, s.height)
        fresh.load_scene(cur)
        pc = cur.push_constants(1 + rnd)
        a, b = ctx.trace_primary_ids(pc), fresh.trace_primary_ids(pc)
        for x, y in zip(a, b):
            assert np.array_equal(x.view(np.uint32), y.view(np.uint32))
        assert np.array_equal(ctx.render(cur, 3), fresh.render(cur, 3))
        fresh.close()
        if rnd == 3:
            r = oracle_mod.OracleScene(cur).trace_primary_ids(pc)
            for x, y in zip(a[:3], r[:3]):
                assert np.array_equal(x, y)
    ctx.counters()
    ctx.close()


def test_refit_rejects_a_different_instance_count(api):
    from helios_b200._lib import HeliosError

    s = scenes.city_scene(n_instances=10, n_meshes=2, width=64, height=36, floors=(2, 3), detail=(1, 2))
    ctx = api.Context(s.width, s.height)
    ctx.load_scene(s)
    with pytest.raises(HeliosError):
        ctx.update_instances(s.instances[:-1])
    ctx.update_instances(s.instances)  # the same transforms: a no-op refit
    ctx.close()


def test_refit_cost_against_rebuild(api):
    """the number DESIGN.md quotes: hl_scene_update_instances vs hl_scene_set_tables at 1023 + 1 instances (configs[3] layout)"""
    s = scenes.city_scene(width=320, height=180)
    ctx = api.Context(s.width, s.height)
    handles = ctx.load_scene(s)
    meshes = [handles[int(i["mesh_index"])] for i in s.instances]
    inst = moved_instances(s, seed=3, frac=0.5, shift=4.0)
    ctx.synchronize()
    t = {}
    for name, fn in (("set_tables", lambda: ctx.set_tables(s.materials, inst, meshes, s.submesh_info, s.lights)), ("update_instances", lambda: ctx.update_instances(inst))):
        fn()
        ctx.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            fn()
        ctx.synchronize()
        t[name] = (time.perf_counter() - t0) / 10 * 1e3
    print(f"{len(inst)} instances: hl_scene_set_tables {t['set_tables']:.3f} ms, hl_scene_update_instances {t['update_instances']:.3f} ms (host wall clock per call)")
    assert t["update_instances"] < t["set_tables"]
    ctx.close()
