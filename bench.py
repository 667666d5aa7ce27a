#!/usr/bin/env python3
"""bench.py — headline benchmark of the path-tracing hot path (BASELINE.json: Mrays/s & ms/frame).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (N = 1): BASELINE.json configs[1] — procedural 1,000,480-triangle terrain + icospheres, Hosek-Wilkie
sky + environment light + directional light, 1920x1080, 1 sample per pixel per step, max depth 8.
A "step" is one progressive frame (one hl_render_frame over the full image).
  value      all rays traced (extension + shadow, device counters) per second, scene + BVH resident in HBM
  e2e        the same metric through the C ABI the way the reference's frame loop runs it: per step the
             192-byte push-constant block comes from host memory, the frame is rendered, tone-mapped, and the
             RGBA8 image is read back to host memory (Renderer::render -> tone_map -> save path)
  roofline   the extend (traversal) kernel against the measured HBM copy bandwidth: algorithmic bytes
             A_ray = 64 + 80*ceil(log8(N/4)) + 192 per ray, time = CUDA events around the
             extend launches of the timed steps (hl_set_profiling)
  cpu_baseline  the CPU oracle (restatement of the reference's GLSL integrator, OpenMP) on a bounded sample
N > 1: samples-per-pixel sharding — every rank renders its own frame indices (weak scaling: K steps per rank)
into a per-GPU sum buffer; one NCCL all-reduce of the accumulation image at the end (inside the timed region).
--impl reference times the reference's own shaders on the host CPU: oracle/_ref/libhelios_glsl_ref.so, i.e. the
reference's GLSL files compiled as C++ (oracle/ref_glsl/; cpu_baseline.kind = "reference"), over the oracle's
traversal and texture units (the engine itself is Vulkan-RT + MSVC only and cannot run here).  If that
library is absent the restatement is timed instead (kind = "port").
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

WORKLOAD = "configs[1]: procedural 1M-triangle terrain + icospheres, Hosek-Wilkie sky + env + directional light, 1920x1080, 1 spp/step, depth 8"


def algorithmic_bytes_per_ray(n_tris: int) -> float:
    """ray read 32 + hit write 32 + one root-to-leaf chain of 80-byte nodes + one leaf of 4 x 48 B"""
    return 64.0 + 80.0 * math.ceil(math.log(max(n_tris / 4.0, 8.0), 8.0)) + 192.0


class ClockSampler:
    """samples nvidia-smi clocks and throttle reasons during the timed region"""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.proc = None
        self.lines = []
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, line in self.lines:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except Exception:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:  # region shorter than the sampling period: use whatever was seen
            for ts, line in self.lines:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except Exception:
                    pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
        out["reasons"] = sorted(reasons)
        return out


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def profiled_traffic_bytes():
    """DRAM bytes per k_extend launch from the committed `ncu --set full` capture (profiles/*_extend_traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum, mean over the captured launches of one frame), or None"""
    files = sorted((ROOT / "profiles").glob("*_extend_traffic.json"))
    if not files:
        return None
    try:
        return float(json.loads(files[-1].read_text())["mean_mb"]) * 1e6
    except Exception:
        return None


def profiled_limits():
    """what ncu says bounds k_extend (same committed capture; first launch = bounce 0): the kernel is instruction-issue bound,
    so the HBM fraction above is small by nature — these are the numbers that describe how close to ITS wall it runs"""
    files = sorted((ROOT / "profiles").glob("*_extend_traffic.json"))
    if not files:
        return None
    try:
        j = json.loads(files[-1].read_text())
        keys = ("issue_active_pct", "lanes_per_instruction", "alu_pipe_pct", "fma_pipe_pct", "dram_throughput_pct", "l2_throughput_pct", "l1_hit_pct", "l2_hit_pct", "warps_active_pct")
        out = {k: j[k][0] for k in keys if j.get(k)}
        out["source"] = j.get("source")
        return out or None
    except Exception:
        return None


def build_scene():
    from helios_b200 import scenes

    return scenes.terrain_scene()  # defaults = configs[1]


def cpu_sample(scene, oracle, sky_cf, frames=1, target_s=8.0):
    """times the oracle on horizontal bands of the same 1080p frame until >= target_s of work"""
    from oracle import oracle as orc

    accum = np.zeros((scene.height, scene.width, 4), np.float32)
    band_h = 54
    order = [10, 9, 11, 8, 12, 7, 13, 6, 14, 5, 15, 4, 16, 3, 17, 2, 18, 1, 19, 0]
    rays0 = oracle.counters.copy()
    t0 = time.time()
    bands = 0
    frame = 1
    while time.time() - t0 < target_s and frame <= 64:
        for b in order:
            pc = scene.push_constants(frame, tile=(0, b * band_h))
            oracle.render_frame(pc, accum, launch=(scene.width, band_h))
            bands += 1
            if time.time() - t0 >= target_s:
                break
        frame += 1
    dt = time.time() - t0
    rays = float((oracle.counters - rays0).sum())
    return rays / dt / 1e6, dt, f"{bands} bands of 1920x{band_h} px (20 bands = one 1080p frame, num_frames=1..{frame - 1}), {rays:.0f} rays in {dt:.1f} s"


def run_reference(args):
    """reference arm: the reference's own shaders on the host CPU, all host threads — oracle/_ref (the reference's GLSL
    compiled as C++, kind "reference") when that library exists, else the restatement (kind "port")"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc

    scene = build_scene()
    cf = orc.sky_coeffs(scene.sun_direction)
    kind = "reference" if orc.ref_lib() is not None else "port"
    o = orc.GlslRefScene(scene, sky_coeffs_override=cf) if kind == "reference" else orc.OracleScene(scene, sky_coeffs_override=cf)
    cores = os.cpu_count() or 1
    per_step = []
    sample = ""
    for s in range(args.warmup + args.steps):
        v, dt, sample = cpu_sample(scene, o, cf, target_s=3.0)
        if s >= args.warmup:
            per_step.append((v, dt))
    value = float(np.mean([v for v, _ in per_step]))
    ms = float(np.mean([dt for _, dt in per_step])) * 1e3
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "each step = bounded sample of the frame on the host CPU"},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
        return

    import torch

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        # stdout carries exactly one JSON line: NCCL's logs go to stderr, and its version banner (a plain printf at
        # NCCL_DEBUG=VERSION, which this image exports) is switched off; an explicit INFO / TRACE request is left alone
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from helios_b200 import abi, api
    from helios_b200.sky import sky_coefficients

    scene = build_scene()
    n_tris = scene.num_triangles
    ctx = api.Context(scene.width, scene.height, device=local_rank)
    t_up = time.time()
    handles = ctx.load_scene(scene, sky_coeffs=sky_coefficients(scene.sun_direction))
    ctx.synchronize()
    t_up = time.time() - t_up
    build = ctx.mesh_build_stats(handles[0])
    if world > 1:
        ctx.set_accum_mode(abi.ACCUM_SUM)
    W, K = args.warmup, args.steps

    from helios_b200 import multi_gpu

    _frames = multi_gpu.frame_indices(rank, world, W + K)  # spp sharding: rank g renders frames g+1, g+1+G, ...

    def frame_index(step):
        return _frames[step]

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (value) ----------------
    ctx.accum_clear()
    for s in range(W):
        ctx.render_frame(scene.push_constants(frame_index(s)))
    pcs = [scene.push_constants(frame_index(W + s)) for s in range(K)]
    acc = None
    if dist is not None:
        # warm-up of the one collective on a scratch image of the same size: communicator / NVLS set-up is a one-off
        # cost of the process, not of a step (measured: ~190 ms at 8 ranks when left inside the timed region)
        acc = torch.as_tensor(multi_gpu.DeviceArray(ctx.accum_device_ptr(), scene.width * scene.height * 4), device=f"cuda:{local_rank}")
        scratch = torch.zeros_like(acc)
        for _ in range(W):
            multi_gpu.all_reduce_sum(scratch, dist)
        torch.cuda.synchronize()
        del scratch
    ctx.reset_counters()
    launches0 = ctx.kernel_launches()
    barrier()
    sampler = ClockSampler(local_rank)
    t0 = time.time()
    ctx.event_record(0)
    for s in range(K):
        ctx.render_frame(pcs[s])
    if dist is not None:
        ctx.synchronize()
        multi_gpu.all_reduce_sum(acc, dist)
        torch.cuda.synchronize()
    ctx.event_record(1)
    ms_total = ctx.event_elapsed_ms(0, 1)
    barrier()
    t1 = time.time()
    clocks = sampler.stop(t0, t1)
    launches = ctx.kernel_launches() - launches0
    c = ctx.counters()
    rays = float(c["extension_rays"] + c["shadow_rays"])
    if dist is not None:
        t = torch.tensor([ms_total, rays], dtype=torch.float64, device=f"cuda:{local_rank}")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms_total, rays = float(tmax[0]), float(t[1])
    value = rays / (ms_total * 1e-3) / 1e6

    # ---------------- per-stage roofline (separate profiled pass, same frames) ----------------
    ctx.set_profiling(True)
    ext_ms = sh_ms = con_ms = frame_ms = 0.0
    ext_rays = 0
    nprof = min(K, 8)
    for s in range(nprof):
        ctx.reset_counters()
        ctx.render_frame(pcs[s])
        cc = ctx.counters()
        ext_ms += float(cc["ms_extend"])
        sh_ms += float(cc["ms_shade"])
        con_ms += float(cc["ms_connect"])
        frame_ms += float(cc["ms_frame"])
        ext_rays += int(cc["extension_rays"])
    ctx.set_profiling(False)
    a_ray = algorithmic_bytes_per_ray(n_tris)
    peak, peak_src = measured_peak_gbs()
    achieved = ext_rays * a_ray / (ext_ms * 1e-3) / 1e9 if ext_ms > 0 else 0.0
    roofline = {
        "bound": "hbm", "kernel": "k_extend", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": profiled_traffic_bytes(),
        "algorithmic_bytes_per_launch": ext_rays * a_ray / max(nprof * int(scene.max_ray_bounces), 1),
        "peak_source": peak_src, "algorithmic_bytes_per_ray": a_ray, "launches_timed": nprof * int(scene.max_ray_bounces),
        "stage_share_of_frame": {"extend": ext_ms / frame_ms, "shade": sh_ms / frame_ms, "connect": con_ms / frame_ms} if frame_ms else None,
        "ncu_bounce0": profiled_limits(),
    }

    # ---------------- end to end through the C ABI with host buffers ----------------
    host_ring = [torch.empty((scene.height, scene.width, 4), dtype=torch.uint8).pin_memory().numpy() for _ in range(8)]  # one per frame that can be in flight (HL_OPT_FRAMES_IN_FLIGHT <= 8)
    host_img = host_ring[0]
    ctx.accum_clear()
    ctx.reset_counters()
    import ctypes as C

    barrier()
    te0 = time.time()
    for s in range(K):
        pc = scene.push_constants(frame_index(W + s))  # host-side PushConstants fill (PathIntegrator::launch_rays)
        if dist is None:
            # fused accumulate + tone-map resolve pass, RGBA8 image copied to pinned host memory on the frame's own stream
            # (every step's image reaches the host; the copy of step s overlaps the rendering of step s + 1)
            ctx.render_frame_readback(pc, host_ring[s % 8], 1.0, abi.TONE_MAP_ACES)
        else:  # sum mode: separate tone-map pass (needs the sample scale)
            ctx.render_frame(pc)
            ctx._chk(ctx.lib.hl_tonemap(ctx.h, C.c_float(1.0), C.c_int(0), C.c_float(1.0 / (s + 1)), host_img.ctypes.data_as(C.c_void_p)))
    barrier()
    te = time.time() - te0
    ce = ctx.counters()
    rays_e = float(ce["extension_rays"] + ce["shadow_rays"])
    if dist is not None:
        t = torch.tensor([te, rays_e], dtype=torch.float64, device=f"cuda:{local_rank}")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        te, rays_e = float(tmax[0]), float(t[1])
    e2e = {"value": rays_e / te / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 192, "d2h_bytes_per_step": scene.width * scene.height * 4, "ms_per_step": te / K * 1e3}

    # ---------------- CPU baseline (rank 0, N = 1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as orc

        cf = orc.sky_coeffs(scene.sun_direction)
        kind = "reference" if orc.ref_lib() is not None else "port"  # oracle/_ref = the reference's GLSL compiled as C++
        o = orc.GlslRefScene(scene, sky_coeffs_override=cf) if kind == "reference" else orc.OracleScene(scene, sky_coeffs_override=cf)
        v, dt, sample = cpu_sample(scene, o, cf, target_s=10.0)
        cpu = {"value": v, "unit": "Mrays/s", "cores": os.cpu_count() or 1, "kind": kind, "sample": sample}

    if rank == 0:
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_total / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": WORKLOAD, "triangles": n_tris, "resolution": [scene.width, scene.height], "spp_per_step": 1, "max_ray_bounces": int(scene.max_ray_bounces),
                "rays_per_step": rays / K / world, "parallelism": f"spp-sharded x{world}" if world > 1 else "single GPU",
                "l2": "per-step working set (ray queues ~400 MB + scene ~200 MB) exceeds the 126 MB L2; no flush between steps",
                "bvh_build_ms": float(build["ms_build"]), "scene_upload_and_build_s": t_up, "wide_nodes": int(build["wide_nodes"]),
            },
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
