#!/usr/bin/env python
"""bench.py -- headline benchmark of the spectral render path (BASELINE.json: samples/s at 1080p
Cornell box).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...    # the reference's CPU path (oracle)

Workload (config[1] of BASELINE.json): Cornell box (main.rs:1538-1635), 1920x1080, 32 spectral samples,
30 bounces.  One STEP = one srt_render_frames call of --frames-per-step frames (default 64), i.e. 64 samples
for every pixel = 132.7 M samples; the default K=16 steps are exactly the 1024 spp of the named config.
With N GPUs every rank renders its own --frames-per-step frames per step (frame-sharded, weak scaling), and
the spectral accumulation buffers are summed onto rank 0 with one NCCL reduce inside the timed region.

Prints ONE JSON line on rank 0 (see the keys at the bottom of main()).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version line to stdout when
# NCCL_DEBUG=VERSION is set in the environment), so everything else is sent to stderr and the JSON line goes to the
# original stdout.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


WIDTH, HEIGHT, N_LAMBDA, BOUNCES = 1920, 1080, 32, 30
SCENE = "cornell"

# SURVEY.md 8(d) cost table: f32 lane-operations per event
COST = dict(slab=28, shape_sphere=46, shape_plain=31, shape_rotated=70, raygen=75, hit_common=36 + 24,
            normal_plain=15, normal_sphere=18, normal_rotated=51, per_light=45, diffuse_cont=80, specular_cont=72,
            resolve=15 + 8)


def ops_per_sample(c: dict, n_lambda: int, n_lights: int) -> float:
    """Algorithmic f32 operations per sample from the ORACLE's event counters."""
    s = c["samples"]
    shapes = c["shape_sphere"] + c["shape_plain"] + c["shape_rotated"]
    normal = (COST["normal_sphere"] * c["shape_sphere"] + COST["normal_plain"] * c["shape_plain"] +
              COST["normal_rotated"] * c["shape_rotated"]) / max(1, shapes)
    diffuse_hits = c["hits"] - c["spec_hits"]
    ops = (c["slab_tests"] * COST["slab"] + c["shape_sphere"] * COST["shape_sphere"] +
           c["shape_plain"] * COST["shape_plain"] + c["shape_rotated"] * COST["shape_rotated"] +
           s * COST["raygen"] + c["hits"] * (COST["hit_common"] + normal) +
           c["rays_shadow"] * COST["per_light"] + c["lit"] * 4 * n_lambda +
           diffuse_hits * (COST["diffuse_cont"] + 2 * n_lambda) + c["spec_hits"] * (COST["specular_cont"] + n_lambda) +
           c["hits"] * n_lambda + s * (6 * n_lambda + COST["resolve"]))
    return ops / s


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(prefix="clocks_", suffix=".csv")
        self.idx = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.proc.wait()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_run(n_frames: int, threads: int, first_frame: int = 0, counters: bool = False, tight: bool = False):
    """The reference's CPU path (oracle/oracle.cpp, platform libm, row-per-task pool) on n_frames of the
    bench workload.  Returns (seconds, samples, counters|None, image).  The oracle is only ever the baseline / the
    checker here."""
    import oracle as O
    O.build()
    O.set_modes(O.MATH_NATIVE, O.RNG_PCG3D)
    sc = O.Scene(N_LAMBDA, SCENE, tight=tight)
    if counters:
        O.counters_reset()
    t0 = time.perf_counter()
    img = sc.render(WIDTH, HEIGHT, n_frames, first_frame=first_frame, intended_frames=1024, max_bounces=BOUNCES,
                    threads=threads)
    dt = time.perf_counter() - t0
    return dt, n_frames * WIDTH * HEIGHT, (O.counters() if counters else None), img


def run_reference(args, rank: int, world: int):
    """--impl reference: the reference's own CPU implementation on all host threads; each step is a bounded
    sample (1 frame of the 1080p Cornell box = 2.07 M samples).  Under torchrun only rank 0 works."""
    if rank != 0:
        return
    import oracle as O
    threads = O.hardware_threads()
    for w in range(args.warmup):
        cpu_reference_run(1, threads, first_frame=w)
    t = 0.0
    samples = 0
    for k in range(args.steps):
        dt, n, _, _ = cpu_reference_run(1, threads, first_frame=args.warmup + k)
        t += dt
        samples += n
    value = samples / t
    line = {
        "impl": "reference", "metric": "samples/s at 1080p Cornell box", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1, reference=True),
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps x 1 frame of the 1920x1080 Cornell box (2.07 M samples each), "
                                   "C++ restatement of the Rust reference (no Rust toolchain in this image)"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, world: int, reference: bool = False) -> dict:
    return {"workload": f"Cornell box (main.rs:1538-1635) {WIDTH}x{HEIGHT}, {N_LAMBDA} spectral samples, "
                        f"{BOUNCES} bounces; 1 step = {1 if reference else args.frames_per_step} frame(s) per rank",
            "scene": SCENE, "width": WIDTH, "height": HEIGHT, "n_lambda": N_LAMBDA, "max_bounces": BOUNCES,
            "frames_per_step": 1 if reference else args.frames_per_step,
            "spp_total": (1 if reference else args.frames_per_step) * args.steps * world,
            "integrator": "cpu" if reference else ("resident" if args.integrator == 1 else "wavefront"),
            "rng": "pcg3d_reference" if args.rng == 0 else "philox", "math": "fast" if args.math == 0 else "exact",
            "parallelism": f"frames sharded over {world} GPU(s), NCCL reduce of spectral accumulation buffers",
            "cache": "working set (265 MB accumulation buffer per GPU) exceeds the 126 MB L2; no explicit flush"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames-per-step", type=int, default=64)
    ap.add_argument("--integrator", type=int, default=1, help="0 wavefront, 1 resident")
    ap.add_argument("--rng", type=int, default=0, help="0 pcg3d (reference), 1 philox")
    ap.add_argument("--math", type=int, default=0, help="0 fast (CUDA f32 libm), 1 exact")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3  # timing rule: W >= 3

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import spectral_raytracer_b200 as srt
    from spectral_raytracer_b200 import scenes
    from spectral_raytracer_b200.distributed import accum_as_tensor, reduce_sum_

    if not torch.cuda.is_available() or srt.native.lib().srt_device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device -- the render backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    F, K, W = args.frames_per_step, args.steps, args.warmup
    npix = WIDTH * HEIGHT
    total_frames = (K + W) * F * world
    flat = scenes.preset(SCENE, N_LAMBDA)
    r = srt.Renderer(flat, WIDTH, HEIGHT, max_bounces=BOUNCES, intended_frames=max(1024, total_frames), rng=args.rng,
                     math=args.math, integrator=args.integrator, device=local_rank)
    r.set_profiling(args.integrator == 0)
    acc_t = accum_as_tensor(r)
    host_img = torch.empty((HEIGHT, WIDTH, 4), dtype=torch.float32).pin_memory()
    host_np = host_img.numpy()

    def frame_base(step):  # rank-th block of F frames of global step `step`
        return (step * world + rank) * F

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (`value`)
    for s in range(W):
        r.render_frames(frame_base(s), F)
    if world > 1:  # warm NCCL
        dist.reduce(acc_t.clone(), dst=0)
    r.clear()
    r.reset_counters()
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    wall0 = time.perf_counter()
    render_ms, launches, stage_ms, stage_n = 0.0, 0, [0.0, 0.0, 0.0], [0, 0, 0]
    for s in range(K):
        r.render_frames(frame_base(W + s), F)
        ms, n = r.last_render_stats()
        render_ms += ms
        launches += n
        sm, sn = r.last_stage_times()
        stage_ms = [a + b for a, b in zip(stage_ms, sm)]
        stage_n = [a + b for a, b in zip(stage_n, sn)]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    if world > 1:
        dist.reduce(acc_t, dst=0)  # spectral accumulation buffers -> rank 0 over NVLink
        r.frames_accumulated = K * F * world
    ev1.record()
    barrier()
    reduce_ms = ev0.elapsed_time(ev1)
    wall = time.perf_counter() - wall0
    clk = clocks.stop()
    t_rank = torch.tensor([render_ms + reduce_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_rank, op=dist.ReduceOp.MAX)
    total_ms = float(t_rank.item())
    counters = r.counters()
    csum = torch.tensor([counters[k] for k in ("samples", "rays_primary", "rays_continuation", "rays_shadow", "lit")],
                        dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(csum)
    samples, rays = float(csum[0]), float(csum[1] + csum[2] + csum[3])
    assert samples == K * F * npix * world, (samples, K * F * npix * world)
    value = samples / (total_ms * 1e-3)

    # ---------------- end to end through the C ABI with host buffers (`e2e`)
    # per step: render F frames, resolve spectrum -> RGB on the device and read the RGBA f32 image back
    # into pinned host memory (what App::render does every frame with FrameUpdate, main.rs:1343-1348);
    # host -> device per step = the kernel-parameter block (scene) of every launch + the control block.
    r.clear()
    barrier()
    e0 = time.perf_counter()
    e2e_launches = 0
    for s in range(K):
        r.render_frames(frame_base(W + s), F)
        e2e_launches += r.last_render_stats()[1] + 1
        r.resolve_rgba_f32(host_np)
    if world > 1:
        total = reduce_sum_(acc_t, K * F, dst=0)
        if rank == 0:
            r.frames_accumulated = total
            r.resolve_rgba_f32(host_np)
    barrier()
    e_rank = torch.tensor([time.perf_counter() - e0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e_rank, op=dist.ReduceOp.MAX)
    e2e_value = samples / float(e_rank.item())
    scene_param_bytes = int(srt.native.lib().srt_launch_param_bytes())  # sizeof(SceneParams), passed by value with every launch
    h2d = int(e2e_launches / K * scene_param_bytes + 32)
    d2h = int(host_np.nbytes + 32 * max(1, e2e_launches // (3 * K)))

    # ---------------- CPU baseline (rank 0, N=1): bounded sample of the same workload
    cpu = None
    cpu_tight = None
    oc = None
    rmse = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle as O
        threads = O.hardware_threads()
        dt, n, oc, _ = cpu_reference_run(1, threads, counters=True)
        frames = int(min(24, max(1, 12.0 / dt)))
        dt2, n2, _, cpu_img = cpu_reference_run(frames, threads, first_frame=1)
        # (the reference blends frame f with ratio 1 / (f + 1), custom_image.rs:71-76: started at frame 1 on an empty image the
        # running mean counts an all-zero frame 0 -- undo that to get the mean of the frames actually rendered)
        cpu_img = cpu_img * np.float32((frames + 1) / frames)
        # RMSE vs the CPU reference (BASELINE.json's metric): the same frames on the GPU, production settings, against
        # the image the CPU baseline just rendered; relative RMSE on linear f32 RGB (SURVEY 8d gate 3).  The noise floor
        # beside it is the same statistic for DISJOINT frames (what two independent estimates of this length differ by).
        def rel_rmse(a, b):
            ok = np.isfinite(a).all(axis=2) & np.isfinite(b).all(axis=2)
            a, b = a[ok].astype(np.float64), b[ok].astype(np.float64)
            return float(np.sqrt(np.mean((a - b) ** 2)) / np.mean(b))
        with srt.Renderer(flat, WIDTH, HEIGHT, intended_frames=1024, max_bounces=BOUNCES, integrator=args.integrator,
                          device=local_rank) as rr:
            rr.render_frames(1, frames)
            same = rr.resolve_rgba_f32()[..., :3]
            rr.clear()
            rr.render_frames(1 + frames, frames)
            other = rr.resolve_rgba_f32()[..., :3]
        # the same port without the reference's avoidable cost items (oracle.cpp -DORACLE_TIGHT, bit-identical images)
        frames_t = max(1, frames // 2)
        dt3, n3, _, _ = cpu_reference_run(frames_t, threads, first_frame=1, tight=True)
        cpu_tight = {"value": n3 / dt3, "unit": "samples/s", "cores": threads, "kind": "port",
                     "sample": f"{frames_t} frames ({n3 / 1e6:.1f} M samples, {dt3:.1f} s) of the same workload with the tight build of "
                               "the port: spectra stored 32 wide, colour weights computed once, closest hit tracked without heap "
                               "vector + sort, reciprocals hoisted -- same arithmetic, bit-identical images"}
        rmse = {"rel_rmse_vs_cpu": rel_rmse(same, cpu_img[..., :3]), "noise_floor": rel_rmse(other, cpu_img[..., :3]),
                "mean_ratio": float(np.nanmean(same) / np.nanmean(cpu_img[..., :3])), "frames": frames,
                "what": "1920x1080 Cornell box, the same frame ids on the GPU (production math) and on the CPU reference port; "
                        "relative RMSE of linear RGB; noise_floor = disjoint frame ids"}
        cpu = {"value": (n + n2) / (dt + dt2), "unit": "samples/s", "cores": threads, "kind": "port",
               "sample": f"{frames + 1} frames of the same 1920x1080 Cornell box ({(n + n2) / 1e6:.1f} M samples, "
                         f"{dt + dt2:.1f} s), C++ restatement of the Rust reference with its cost structure "
                         "(row-per-task pool, 528-byte spectra, per-ray Vec + sort, per-sample get_rgb_early)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- rooflines
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    n_launch = max(1, launches if args.integrator == 1 else stage_n[2])
    kernel_ms = (render_ms if args.integrator == 1 else stage_ms[2]) / n_launch
    # Algorithmic HBM bytes of the dominant kernel, from THIS run's counters (rank 0).
    #  * SURVEY.md 8(d)'s figure for a wavefront whose whole path state lives in HBM: 576 B per path-bounce (ray 32 B +
    #    throughput 4*n_lambda + radiance 4*n_lambda, read and written) + 4*n_lambda per accumulation write (quoted beside).
    #  * what THIS wavefront's k_shade moves: per path-bounce ray 32 B + hit 8 B read, throughput 4*n_lambda
    #    read unless the path is fresh; ray 32 B + throughput 4*n_lambda written if the path goes on; radiance is not
    #    carried per path -- a lit event adds 4*n_lambda to the pixel's record (read + write, the buffer exceeds L2).
    #  * the RESIDENT integrator keeps ray and throughput on chip; what it must move is the accumulation buffer, which
    #    is larger than L2 and whose every pixel record (4*n_lambda B) is read and written once per frame.
    bounces_rank0 = counters["rays_primary"] + counters["rays_continuation"]
    state_bytes = bounces_rank0 * 2 * (32 + 4 * N_LAMBDA + 4 * N_LAMBDA) + counters["lit"] * 4 * N_LAMBDA
    shade_bytes = (bounces_rank0 * 40 + counters["rays_continuation"] * 4 * N_LAMBDA +
                   counters["rays_continuation"] * (32 + 4 * N_LAMBDA) + counters["lit"] * 2 * 4 * N_LAMBDA)
    accum_bytes = counters["samples"] * 2 * 4 * N_LAMBDA
    alg_bytes = accum_bytes if args.integrator == 1 else shade_bytes
    achieved_gbs = alg_bytes / n_launch / (kernel_ms * 1e-3) / 1e9
    traffic = None
    try:  # measured DRAM bytes of the dominant kernel (one ncu --set full capture, committed under profiles/)
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        key = "k_resident" if args.integrator == 1 else "k_shade"
        if key in tr:
            traffic = tr[key]["dram_bytes_per_sample"] * counters["samples"] / n_launch
    except OSError:
        pass
    roofline = {"bound": "hbm", "kernel": "k_resident" if args.integrator == 1 else "k_shade",
                "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                "traffic": traffic, "algorithmic_bytes_per_launch": alg_bytes / n_launch, "peak_source": peak_src,
                "algorithmic_bytes_per_sample": alg_bytes / max(1, counters["samples"]),
                "survey_8d_hbm_resident_state_bytes_per_sample": state_bytes / max(1, counters["samples"]),
                "note": ("resident integrator: path state stays on chip, the algorithmic HBM traffic is the accumulation buffer "
                         "(one 128-byte pixel record read + written per sample); the kernel is bound by instruction issue, see "
                         "roofline_fp32 (north star: FP32 issue rate).  SURVEY 8(d)'s figure for an HBM-resident wavefront state "
                         "is given beside it: at this throughput it would need more than the HBM peak."
                         if args.integrator == 1 else
                         "wavefront k_shade: ray + hit + throughput read, ray + throughput written for surviving paths, 256 B per lit "
                         "event (accumulation record read + write); SURVEY 8(d)'s figure (radiance carried per path too) beside it")}
    sm_mhz = clk.get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
    fp32 = None
    if oc is not None:
        ops = ops_per_sample(oc, N_LAMBDA, 1)
        peak_ops = 148 * 128 * sm_mhz * 1e6
        fp32 = {"bound": "fp32_issue", "ops_per_sample": ops, "achieved": value * ops / 1e12,
                "peak": peak_ops / 1e12, "unit": "Tlane-op/s", "frac": value * ops / peak_ops,
                "note": "SURVEY 8(d) cost table x oracle event counters on the same config; peak = 148 SMs x 128 lanes "
                        "x SM clock observed during the run"}

    line = {
        "metric": "samples/s at 1080p Cornell box", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "mrays_per_s": rays / (total_ms * 1e-3) / 1e6,
        "rays_per_sample": rays / samples,
        "reduce_ms": reduce_ms, "wall_s_timed_region": wall,
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "what": "per step: srt_render_frames + srt_resolve_rgba_f32 into pinned host memory"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "roofline_fp32": fp32,
        "cpu_baseline": cpu,
        "cpu_baseline_tight": cpu_tight,
        "rmse": rmse,
    }
    emit(line)
    r.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
