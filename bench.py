#!/usr/bin/env python3
"""bench.py — headline benchmark of the path-tracing hot path (BASELINE.json: Mrays/s & ms/frame).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 1|4]

Workload (default, --config 1): BASELINE.json configs[1] — procedural 1,000,480-triangle terrain + icospheres,
Hosek-Wilkie sky + environment light + directional light, 1920x1080, 1 sample per pixel per step, max depth 8.
A "step" is one progressive frame (one hl_render_frame over the full image).
  value      all rays traced (extension + shadow, device counters) per second, scene + BVH resident in HBM, CUDA events on
             the library's stream, max over ranks
  e2e        the same metric through the C ABI the way the reference's frame loop runs it: per step the 192-byte
             push-constant block comes from host memory, the frame is rendered, tone-mapped, and the RGBA8 image is read
             back to pinned host memory (Renderer::render -> tone_map -> save path); host wall clock between barriers
  roofline   the extend (traversal) kernel against the measured HBM copy bandwidth: algorithmic bytes
             A_ray = 64 + 80*ceil(log8(N/4)) + 192 per ray over the non-empty k_extend launches of a
             profiled pass (hl_set_profiling: one CUDA-event pair per launch, k_tail excluded)
  cpu_baseline  the reference's shaders on the host CPU (oracle/_ref, kind "reference"; else the restatement, "port")
N > 1 (torchrun, one process per GPU): samples-per-pixel sharding — rank g renders frame indices g+1, g+1+G, ... into its
own SUM image (weak scaling: K steps per rank); the images are combined by ONE NCCL reduction issued by the library itself
(hl_comm_init_rank + hl_accum_all_reduce, helios_b200/csrc/hl_comm.cu).  The device leg times K frames + the all-reduce;
the e2e leg times, per rank, K x (host push constants -> frame -> rank-local preview image read back to the host), then
the reduction to rank 0, ONE tone map with 1 / (N K) and the final image's read-back on rank 0.  `parity` compares that
reduced image with the running mean of the same N K frames rendered on rank 0 alone.
--config 4: BASELINE.json configs[3] — instanced 19.8M-triangle city, 3840x2160, --spp (default 256) samples divided
among the ranks (STRONG scaling), same reduction; one step = one sample per pixel on one rank.
--impl reference times the reference's own shaders on the host CPU over the same frames: oracle/_ref/libhelios_glsl_ref.so,
i.e. the reference's GLSL files compiled as C++ (oracle/ref_glsl/), on all host threads, one full frame per step (the
engine itself is Vulkan-RT + MSVC only and cannot run here); kind = "port" if only the restatement exists.
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

WORKLOADS = {
    1: "configs[1]: procedural 1M-triangle terrain + icospheres, Hosek-Wilkie sky + env + directional light, 1920x1080, 1 spp/step, depth 8",
    4: "configs[3]: instanced 19.8M-triangle city (1023 instances of 32 meshes + ground), 3840x2160, spp sharded across the GPUs, depth 8",
}


def algorithmic_bytes_per_ray(n_tris: int, n_instances: int = 1) -> float:
    """ray read 32 + hit write 32 + one root-to-leaf chain of 80-byte nodes + one leaf of 4 x 48 B;
    two-level scenes add the instance-tree chain and one 3x4 transform"""
    a = 64.0 + 80.0 * math.ceil(math.log(max(n_tris / 4.0, 8.0), 8.0)) + 192.0
    if n_instances > 1:
        a += 80.0 * math.ceil(math.log(max(n_instances, 8.0), 8.0)) + 48.0
    return a


class ClockSampler:
    """samples nvidia-smi clocks and throttle reasons during the timed region"""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_indices):
        """one nvidia-smi poller for all the GPUs of the job (started by rank 0 only: N pollers at 10 Hz were N processes
        taking driver locks while N ranks enqueue ~700 launches per second each)"""
        self.proc = None
        self.lines = []
        if not gpu_indices:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--id=" + ",".join(str(g) for g in gpu_indices), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
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
            out["samples"] = len(sm)
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


def profiled_static():
    """what the committed `ncu --set full` capture of k_extend says (profiles/*_extend_traffic.json, newest): DRAM bytes per
    launch and the counters that describe the wall this kernel really runs into (it is instruction-issue bound).  STATIC:
    read from the repository, not measured by this run — labelled as such in the line."""
    files = sorted((ROOT / "profiles").glob("*_extend_traffic.json"))
    if not files:
        return None, None
    try:
        j = json.loads(files[-1].read_text())
        keys = ("issue_active_pct", "lanes_per_instruction", "alu_pipe_pct", "fma_pipe_pct", "dram_throughput_pct", "l2_throughput_pct", "l1_hit_pct", "l2_hit_pct", "warps_active_pct")
        out = {k: j[k][0] for k in keys if j.get(k)}
        out["source"] = f"static: {files[-1].name} ({j.get('source')}), bounce-0 launch"
        return float(j["mean_mb"]) * 1e6, out
    except Exception:
        return None, None


def build_scene(config: int):
    from helios_b200 import scenes

    if config == 4:
        return scenes.city_scene(width=3840, height=2160)
    return scenes.terrain_scene()  # defaults = configs[1]


def shared_config(config: int, scene, args):
    """the keys both arms print (the driver compares them)"""
    c = {"workload": WORKLOADS[config], "triangles": int(scene.num_triangles), "resolution": [int(scene.width), int(scene.height)], "spp_per_step": 1,
         "max_ray_bounces": int(scene.max_ray_bounces),
         "l2": "per-step working set (ray queues ~400 MB at 1080p + scene >= 200 MB) exceeds the 126 MB L2; no flush between steps"}
    if config == 4:
        c["spp_total"] = int(args.spp)
    return c


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def make_cpu_scene(scene, threads: int):
    """the CPU arm: the reference's shaders (oracle/_ref) when that library exists, else the restatement; OpenMP over `threads`
    (torchrun exports OMP_NUM_THREADS=1 to its workers: the thread count is set explicitly)"""
    os.environ["OMP_NUM_THREADS"] = str(threads)
    from oracle import oracle as orc

    orc.set_threads(threads)
    cf = orc.sky_coeffs(scene.sun_direction) if scene.sun_direction is not None else None
    kind = "reference" if orc.ref_lib() is not None else "port"
    o = orc.GlslRefScene(scene, sky_coeffs_override=cf) if kind == "reference" else orc.OracleScene(scene, sky_coeffs_override=cf)
    return o, kind


def cpu_frames(scene, o, frames, target_s=None, band_h=None):
    """renders `frames` (push-constant frame indices) on the CPU, full frames — or horizontal bands of a frame until
    target_s seconds are spent when the frame is too big for the budget; returns (Mrays/s, seconds, rays, description)"""
    accum = np.zeros((scene.height, scene.width, 4), np.float32)
    rays0 = o.counters.copy()
    t0 = time.time()
    done = 0
    if band_h is None:
        for f in frames:
            o.render_frame(scene.push_constants(f), accum)
            done += 1
            if target_s is not None and time.time() - t0 >= target_s:
                break
        what = f"{done} full {scene.width}x{scene.height} frame(s)"
    else:
        nb = scene.height // band_h
        order = sorted(range(nb), key=lambda b: abs(b - nb // 2))  # centre out: representative of the frame's mix
        for f in frames:
            for b in order:
                o.render_frame(scene.push_constants(f, tile=(0, b * band_h)), accum, launch=(scene.width, band_h))
                done += 1
                if time.time() - t0 >= target_s:
                    break
            if time.time() - t0 >= target_s:
                break
        what = f"{done} bands of {scene.width}x{band_h} px ({nb} bands = one frame)"
    dt = time.time() - t0
    rays = float((o.counters - rays0).sum())
    return rays / dt / 1e6, dt, rays, f"{what}, {rays:.0f} rays in {dt:.2f} s"


def run_reference(args):
    """reference arm: the reference's own shaders on the host CPU, all host threads, the same frames as the GPU arm —
    one full frame per step for --config 1; a bounded band sample of the 4K frame per step for --config 4"""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    scene = build_scene(args.config)
    threads = host_threads()
    o, kind = make_cpu_scene(scene, threads)
    per_step, sample = [], ""
    for s in range(args.warmup + args.steps):
        if args.config == 4:
            v, dt, rays, sample = cpu_frames(scene, o, [1 + s], target_s=3.0, band_h=54)
        else:
            v, dt, rays, sample = cpu_frames(scene, o, [1 + s])
        if s >= args.warmup:
            per_step.append((rays, dt))
    rays = sum(r for r, _ in per_step)
    secs = sum(d for _, d in per_step)
    value = rays / secs / 1e6
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": secs / max(len(per_step), 1) * 1e3, "higher_is_better": True, "scaling": "strong" if args.config == 4 else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": shared_config(args.config, scene, args),
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": kind, "sample": f"per step: {sample}"},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", type=int, default=1, choices=(1, 4))
    ap.add_argument("--spp", type=int, default=256, help="--config 4: samples per pixel in total (divided among the ranks)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sustain-s", type=float, default=2.0, help="length of the extra sustained-throughput loop (N = 1, config 1), 0 = off")
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
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))  # plumbing: barrier, id broadcast, max over ranks
    from helios_b200 import abi, api, multi_gpu
    from helios_b200.sky import sky_coefficients

    dev = f"cuda:{local_rank}"
    scene = build_scene(args.config)
    n_tris = scene.num_triangles
    n_px = scene.width * scene.height
    ctx = api.Context(scene.width, scene.height, device=local_rank)
    t_up = time.time()
    handles = ctx.load_scene(scene, sky_coeffs=sky_coefficients(scene.sun_direction) if scene.sun_direction is not None else None)
    ctx.synchronize()
    t_up = time.time() - t_up
    builds = [ctx.mesh_build_stats(h) for h in handles]
    if world > 1:
        ctx.set_accum_mode(abi.ACCUM_SUM)
        # the data-path communicator belongs to the library (hl_comm.cu): rank 0 makes the id, torch.distributed carries its 128 bytes
        uid = torch.zeros(abi.COMM_ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(api.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.comm_init_rank(bytes(uid.cpu().numpy().tobytes()), world, rank)
    W = args.warmup
    if args.config == 4:
        assert args.spp % world == 0, "--spp must be divisible by the number of GPUs"
        K = args.spp // world  # strong scaling: the samples are divided
    else:
        K = args.steps
    _frames = multi_gpu.frame_indices(rank, world, W + K)  # spp sharding: rank g renders frames g+1, g+1+G, ...

    def frame_index(step):
        return _frames[step]

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_sum(vals_max, vals_sum):
        if dist is None:
            return vals_max, vals_sum
        a = torch.tensor(vals_max, dtype=torch.float64, device=dev)
        b = torch.tensor(vals_sum, dtype=torch.float64, device=dev)
        dist.all_reduce(a, op=dist.ReduceOp.MAX)
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
        return [float(x) for x in a], [float(x) for x in b]

    # ---------------- device-resident throughput (value) ----------------
    ctx.accum_clear()
    for s in range(W):
        ctx.render_frame(scene.push_constants(frame_index(s)))
    pcs = [scene.push_constants(frame_index(W + s)) for s in range(K)]
    collective = None
    if dist is not None:
        # warm-up of the one collective (communicator channels / NVLS set-up are a one-off cost of the process, ~190 ms at 8
        # ranks), then the collective timed ALONE: barrier first, so the number is the reduction, not straggler wait
        for _ in range(W):
            ctx.accum_all_reduce()
            ctx.accum_reduce(0)  # (ncclReduce connects its own channels on first use: ~1 s at 8 ranks when left inside the e2e region)
        ms_coll = []
        for _ in range(5):
            barrier()
            ctx.event_record(2)
            ctx.accum_all_reduce()
            ctx.event_record(3)
            ms_coll.append(ctx.event_elapsed_ms(2, 3))
        (ms_c,), _ = max_sum([float(np.median(ms_coll))], [0.0])
        nbytes = n_px * 16
        collective = {"op": "ncclAllReduce(sum, f32) of the accumulation image, in place, timed alone after a barrier (median of 5, max over ranks)",
                      "bytes": nbytes, "ms": ms_c, "algbw_GBps": nbytes / ms_c / 1e6, "busbw_GBps": nbytes / ms_c / 1e6 * 2.0 * (world - 1) / world}
    ctx.accum_clear()
    ctx.reset_counters()
    launches0 = ctx.kernel_launches()
    barrier()
    sampler = ClockSampler(list(range(world)) if rank == 0 else [])
    if rank == 0:
        time.sleep(0.3)  # (the poller's NVML start-up stays outside the timed region)
    barrier()
    t0 = time.time()
    ctx.event_record(0)
    for s in range(K):
        ctx.render_frame(pcs[s])
    t_enq = time.time() - t0
    ctx.event_record(6)
    if dist is not None:
        ctx.accum_all_reduce()  # asynchronous, on the library's stream behind the frames in flight
    ctx.event_record(1)
    ms_total = ctx.event_elapsed_ms(0, 1)
    if os.environ.get("HL_BENCH_DEBUG"):
        print(f"[rank {rank}] device leg: {K} frames enqueued in {t_enq * 1e3:.2f} ms host time, frames {ctx.event_elapsed_ms(0, 6):.2f} ms + reduce {ctx.event_elapsed_ms(6, 1):.3f} ms on the device; "
              f"cpus {host_threads()} of {os.cpu_count()}", file=sys.stderr, flush=True)
    barrier()
    t1 = time.time()
    clocks = sampler.stop(t0, t1)
    launches = ctx.kernel_launches() - launches0
    c = ctx.counters()
    rays = float(c["extension_rays"] + c["shadow_rays"])
    (ms_total,), (rays,) = max_sum([ms_total], [rays])
    value = rays / (ms_total * 1e-3) / 1e6

    # ---------------- per-stage roofline (separate profiled pass, same frames) ----------------
    ctx.set_profiling(True)
    ext_ms = sh_ms = con_ms = tail_ms = frame_ms = 0.0
    ext_rays = ext_launches = 0
    b0_ms = b0_rays = 0.0
    nprof = min(K, 8)
    for s in range(nprof):
        ctx.render_frame(pcs[s])
        cc = ctx.counters()
        frame_ms += float(cc["ms_frame"])
        prof, _, _ = ctx.bounce_profile()
        for b, p in enumerate(prof):
            tail_ms += float(p["ms_tail"])
            sh_ms += float(p["ms_shade"])
            con_ms += float(p["ms_connect"])
            if int(p["extension_rays"]) > 0:  # a launch that found an empty queue (after Russian roulette / the tail kernel) is not a timed launch
                ext_ms += float(p["ms_extend"])
                ext_rays += int(p["extension_rays"])
                ext_launches += 1
                if b == 0:
                    b0_ms += float(p["ms_extend"])
                    b0_rays += int(p["extension_rays"])
    ctx.set_profiling(False)
    a_ray = algorithmic_bytes_per_ray(n_tris, len(scene.instances))
    peak, peak_src = measured_peak_gbs()
    achieved = ext_rays * a_ray / (ext_ms * 1e-3) / 1e9 if ext_ms > 0 else 0.0
    traffic_static, ncu_static = profiled_static()
    roofline = {
        "bound": "hbm", "kernel": "k_extend", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic_static if args.config == 1 else None, "traffic_source": "static: committed ncu capture of the bounce-0 launch, not measured by this run" if args.config == 1 else None,
        "algorithmic_bytes_per_launch": ext_rays * a_ray / max(ext_launches, 1), "ms_per_launch": ext_ms / max(ext_launches, 1),
        "peak_source": peak_src, "algorithmic_bytes_per_ray": a_ray, "launches_timed": ext_launches,
        "bounce0": {"rays_per_launch": b0_rays / nprof, "ms_per_launch": b0_ms / nprof, "achieved": (b0_rays * a_ray / (b0_ms * 1e-3) / 1e9) if b0_ms > 0 else None},
        "stage_share_of_frame": {"extend": ext_ms / frame_ms, "tail": tail_ms / frame_ms, "shade": sh_ms / frame_ms, "connect": con_ms / frame_ms} if frame_ms else None,
        "stated_bound": "instruction issue (the BVH is L1/L2-served: DRAM traffic is ~0.25x the algorithmic bytes); see ncu_static.issue_active_pct / lanes_per_instruction",
        "ncu_static": ncu_static if args.config == 1 else None,
    }

    # ---------------- end to end through the C ABI with host buffers ----------------
    host_ring = [torch.empty((scene.height, scene.width, 4), dtype=torch.uint8).pin_memory().numpy() for _ in range(8)]  # one per frame that can be in flight (HL_OPT_FRAMES_IN_FLIGHT <= 8)
    final_img = torch.empty((scene.height, scene.width, 4), dtype=torch.uint8).pin_memory().numpy()
    ctx.accum_clear()
    ctx.reset_counters()
    barrier()
    te0 = time.time()
    for s in range(K):
        pc = scene.push_constants(frame_index(W + s))  # host-side PushConstants fill (PathIntegrator::launch_rays)
        # fused accumulate + tone-map resolve pass, RGBA8 image copied to pinned host memory on the frame's own copy stream
        # (every step's image reaches the host; the copy of step s overlaps the rendering of step s + 1).  N > 1: the image is
        # the rank's own progressive preview (sum / frames so far)
        ctx.render_frame_readback(pc, host_ring[s % 8], 1.0, abi.TONE_MAP_ACES)
    t_e2e_enq = time.time() - te0
    if dist is not None:
        ctx.accum_reduce(0)  # the one collective, asynchronous on the library's stream
        if rank == 0:
            ctx.tonemap(1.0, abi.TONE_MAP_ACES, sample_scale=1.0 / (world * K), out=final_img)  # final picture on the host (synchronises)
    barrier()
    te = time.time() - te0
    if os.environ.get("HL_BENCH_DEBUG"):
        print(f"[rank {rank}] e2e leg: {K} steps enqueued in {t_e2e_enq * 1e3:.2f} ms host time, whole region {te * 1e3:.2f} ms", file=sys.stderr, flush=True)
    ce = ctx.counters()
    rays_e = float(ce["extension_rays"] + ce["shadow_rays"])
    (te,), (rays_e,) = max_sum([te], [rays_e])
    e2e = {"value": rays_e / te / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 192, "d2h_bytes_per_step": n_px * 4 + (n_px * 4 // K if world > 1 else 0), "ms_per_step": te / K * 1e3,
           "what": "per rank and step: 192 B push constants from host memory -> frame -> fused resolve + tone map -> RGBA8 image to pinned host memory"
                   + ("; then ncclReduce to rank 0 + one tone map (1 / (N K)) + the final image's read-back, all inside the timed region" if world > 1 else "")}

    # ---------------- N > 1: the reduced image against one GPU rendering the same N K frames ----------------
    parity = None
    if dist is not None:
        reduced = ctx.read_accum() if rank == 0 else None  # rank 0 holds the sum after hl_accum_reduce
        barrier()
        if rank == 0:
            ctx.accum_clear()
            every = sorted(f for r in range(world) for f in multi_gpu.frame_indices(r, world, W + K)[W:])
            # the reference's running mean weights frame f with 1/f (path_trace_rgen.glsl:234-237): renumbering is not allowed, so the
            # single-GPU image of THESE frames is their plain mean — rendered in SUM mode on one GPU and divided once
            for f in every:
                ctx.render_frame(scene.push_constants(f))
            single = ctx.read_accum()[..., :3] / float(len(every))
            multi = reduced[..., :3] / float(world * K)
            d = np.abs(multi - single)
            img1 = ctx.tonemap(1.0, abi.TONE_MAP_ACES, sample_scale=1.0 / len(every))
            parity = {"what": "max |reduced / (N K) - single-GPU sum / (N K)| over all pixels and channels, rank 0, same frame indices (fp32 summation order is the only difference)",
                      "frames": len(every), "max_abs_diff": float(d.max()), "mean_abs_diff": float(d.mean()), "mean_radiance": float(single.mean()),
                      "rgba8_pixels_differing": int((img1[..., :3] != final_img[..., :3]).any(-1).sum()), "rgba8_max_diff": int(np.abs(img1[..., :3].astype(np.int32) - final_img[..., :3].astype(np.int32)).max())}
        barrier()

    # ---------------- sustained throughput: does the number hold for seconds? (N = 1) ----------------
    sustained = None
    if world == 1 and args.config == 1 and args.sustain_s > 0:
        n_sus = max(int(args.sustain_s / (ms_total / K * 1e-3)), K)
        ctx.set_accum_mode(abi.ACCUM_RUNNING_MEAN)
        ctx.accum_clear()
        ctx.reset_counters()
        barrier()
        samp2 = ClockSampler([local_rank])
        time.sleep(0.3)
        ts0 = time.time()
        ctx.event_record(4)
        for s in range(n_sus):
            ctx.render_frame(pcs[s % K])
        ctx.event_record(5)
        ms_sus = ctx.event_elapsed_ms(4, 5)
        ts1 = time.time()
        cs = ctx.counters()
        sustained = {"frames": n_sus, "seconds": ms_sus * 1e-3, "value": float(cs["extension_rays"] + cs["shadow_rays"]) / ms_sus / 1e3, "unit": "Mrays/s", "ms_per_step": ms_sus / n_sus,
                     "clocks": samp2.stop(ts0, ts1)}

    # ---------------- CPU baseline (rank 0, N = 1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        o, kind = make_cpu_scene(scene, threads)
        if args.config == 4:
            v, dt, _, sample = cpu_frames(scene, o, range(1, 65), target_s=10.0, band_h=54)
        else:
            v, dt, _, sample = cpu_frames(scene, o, range(1, 65), target_s=10.0)
        cpu = {"value": v, "unit": "Mrays/s", "cores": threads, "kind": kind, "sample": sample}

    if rank == 0:
        cfg = shared_config(args.config, scene, args)
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_total / K,
            "higher_is_better": True, "scaling": "strong" if args.config == 4 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "details": {"rays_per_step": rays / K / world, "parallelism": f"spp-sharded x{world}, one ncclAllReduce per image (hl_accum_all_reduce)" if world > 1 else "single GPU",
                        "bvh_build_ms": float(sum(float(b["ms_build"]) for b in builds)), "scene_upload_and_build_s": t_up, "wide_nodes": int(sum(int(b["wide_nodes"]) for b in builds)),
                        "sah_cost_mesh0": float(builds[0]["sah_cost"]), "ms_total": ms_total},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        if collective is not None:
            line["collective"] = collective
        if parity is not None:
            line["parity"] = parity
        if sustained is not None:
            line["sustained"] = sustained
        print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
