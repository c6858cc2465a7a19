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

Prints ONE JSON line on rank 0.  Beside the contract's keys it carries, measured in the same run:
  roofline       FP32-issue roofline of the dominant kernel (the bound SURVEY.md 8d names); roofline_hbm beside it
  exact_math / philox   the same workload in the sample-exact math mode / with the north star's Philox RNG
  configs        BASELINE.json's other configs (C0, C2, C3, C4), each rank rendering its shard, reduce included
  strong         the named config as a fixed job: 1024 spp split over the N ranks, reduce + resolve inside the clock
  reduce_check   (N >= 2) srt_reduce of libsrt_nccl.so against the torch.distributed path, and its time
  rmse           the converged-image gate: 480x270, 1024 spp, production math vs the committed oracle image
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import re
import subprocess
import sys
import tempfile
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
SPP_NAMED = 1024  # BASELINE.json config[1]

# SURVEY.md 8(d) cost table: f32 lane-operations per event
COST = dict(slab=28, shape_sphere=46, shape_plain=31, shape_rotated=70, raygen=75, hit_common=36 + 24,
            normal_plain=15, normal_sphere=18, normal_rotated=51, per_light=45, diffuse_cont=80, specular_cont=72,
            resolve=15 + 8)


def ops_per_sample(c: dict, n_lambda: int) -> float:
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


def cpu_reference_run(n_frames: int, threads: int, first_frame: int = 0, counters: bool = False, tight: bool = False,
                      width: int = WIDTH, height: int = HEIGHT):
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
    img = sc.render(width, height, n_frames, first_frame=first_frame, intended_frames=SPP_NAMED, max_bounces=BOUNCES,
                    threads=threads)
    dt = time.perf_counter() - t0
    return dt, n_frames * width * height, (O.counters() if counters else None), img


def workload_config() -> dict:
    """What is measured -- the same dictionary in both arms (what differs between them is in `arm`)."""
    return {"workload": f"Cornell box (main.rs:1538-1635) {WIDTH}x{HEIGHT}, {N_LAMBDA} spectral samples, {BOUNCES} bounces, "
                        f"{SPP_NAMED} spp (BASELINE.json config 1)",
            "scene": SCENE, "width": WIDTH, "height": HEIGHT, "n_lambda": N_LAMBDA, "max_bounces": BOUNCES,
            "spp": SPP_NAMED, "rng": "pcg3d_reference",
            "cache": "working set (265 MB spectral accumulation buffer per GPU) exceeds the 126 MB L2; no explicit flush"}


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
        "config": workload_config(),
        "arm": {"frames_per_step": 1, "integrator": "cpu", "math": "platform libm (glibc)",
                "parallelism": f"{threads} host threads, one task per image row (main.rs:1286-1307)"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps x 1 frame of the 1920x1080 Cornell box (2.07 M samples each), "
                                   "C++ restatement of the Rust reference (no Rust toolchain in this image)"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def newest_ncu_traffic(kernel: str):
    """DRAM bytes per sample of `kernel` from the NEWEST profiles/rNN_<kernel>_vMM_ncu_summary.txt (one ncu --set full
    capture per kernel change, scripts/ncu_summary.py --samples N writes the sample count of the captured launch)."""
    best = None
    for path in glob.glob(os.path.join(ROOT, "profiles", f"r*_{kernel}_*ncu_summary.txt")):
        m = re.search(r"r(\d+)_" + re.escape(kernel) + r"_v(\d+)", os.path.basename(path))
        if not m:
            continue
        key = (int(m.group(1)), int(m.group(2)))
        if best is None or key > best[0]:
            best = (key, path)
    if best is None:
        return None, None
    rd = wr = samples = None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    for line in open(best[1]):
        f = line.split()
        if len(f) >= 3 and f[0] == "dram__bytes_read.sum" and f[2] in unit:
            rd = float(f[1]) * unit[f[2]]
        elif len(f) >= 3 and f[0] == "dram__bytes_write.sum" and f[2] in unit:
            wr = float(f[1]) * unit[f[2]]
        elif len(f) >= 2 and f[0] == "samples_in_launch":
            samples = float(f[1])
    if rd is None or wr is None or not samples:
        return None, os.path.relpath(best[1], ROOT)
    return (rd + wr) / samples, os.path.relpath(best[1], ROOT)


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
    ap.add_argument("--no-extras", action="store_true", help="headline numbers only (skip configs / strong / modes)")
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
    from spectral_raytracer_b200.distributed import accum_as_tensor, frame_shard, reduce_contexts_

    if not torch.cuda.is_available() or srt.native.lib().srt_device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device -- the render backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # host-side group: ranks that wait for rank 0's single-process checks must not spin in an NCCL kernel on the
        # devices rank 0 is measuring (a waiting NCCL barrier made srt_reduce look 6x slower: 2.75 instead of 0.45 ms
        # for the 265 MB buffers, 20 ms instead of 1.7 ms for 1.06 GB on four devices)
        host_group = dist.new_group(backend="gloo")

    F, K, W = args.frames_per_step, args.steps, args.warmup
    npix = WIDTH * HEIGHT
    total_frames = (K + W) * F * world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(xs):
        t = torch.tensor(list(xs), dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t)
        return [float(v) for v in t]

    def timed_render(r, n_steps, n_warm, frames, *, reduce=True):
        """n_warm untimed + n_steps timed steps of `frames` frames on every rank (each rank its own frame ids), device
        time from the context's CUDA events; with N > 1 one reduce of the accumulation buffers onto rank 0 inside the
        clock.  Returns whole-job numbers (times: max over ranks)."""
        acc = accum_as_tensor(r)
        base = lambda s: (s * world + rank) * frames  # noqa: E731
        for s in range(n_warm):
            r.render_frames(base(s), frames)
        if world > 1 and reduce:
            dist.reduce(acc.clone(), dst=0)  # warm NCCL for this message size
        r.clear()
        r.reset_counters()
        barrier()
        ms, launches = 0.0, 0
        for s in range(n_steps):
            r.render_frames(base(n_warm + s), frames)
            m, n = r.last_render_stats()
            ms += m
            launches += n
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        if world > 1 and reduce:
            dist.reduce(acc, dst=0)
            if rank == 0:
                r.frames_accumulated = n_steps * frames * world
        ev1.record()
        barrier()
        reduce_ms = ev0.elapsed_time(ev1) if world > 1 and reduce else 0.0
        total_ms = max_over_ranks(ms + reduce_ms)
        c = r.counters()
        samples, prim, cont, shad = sum_over_ranks([c["samples"], c["rays_primary"], c["rays_continuation"], c["rays_shadow"]])
        return {"samples": samples, "rays": prim + cont + shad, "ms": total_ms, "render_ms_rank": ms, "reduce_ms": reduce_ms,
                "launches": launches, "counters_rank": c}

    flat = scenes.preset(SCENE, N_LAMBDA)
    r = srt.Renderer(flat, WIDTH, HEIGHT, max_bounces=BOUNCES, intended_frames=max(SPP_NAMED, total_frames), rng=args.rng,
                     math=args.math, integrator=args.integrator, device=local_rank)
    r.set_profiling(args.integrator == 0)
    acc_t = accum_as_tensor(r)

    def frame_base(step):  # rank-th block of F frames of global step `step`
        return (step * world + rank) * F

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
    total_ms = max_over_ranks(render_ms + reduce_ms)
    counters = r.counters()
    samples, prim, cont, shad, lit = sum_over_ranks([counters[k] for k in ("samples", "rays_primary", "rays_continuation",
                                                                            "rays_shadow", "lit")])
    rays = prim + cont + shad
    assert samples == K * F * npix * world, (samples, K * F * npix * world)
    value = samples / (total_ms * 1e-3)

    # ---------------- end to end through the C ABI with host buffers (`e2e`)
    # App::render's protocol (main.rs:1338-1357) through srt_render_progressive: per step F frames, then the RGBA8
    # image of everything accumulated so far (AppActions::FrameUpdate: From<CustomImage> for DynamicImage) is resolved
    # on the device and copied into pinned host memory on a second stream while the next step renders, and handed to
    # the host callback; at the end the f32 image (CustomImage.data) is read back as well.  host -> device per step =
    # the kernel-parameter block (scene) of every launch + the control block.
    r.clear()
    host_img = torch.empty((HEIGHT, WIDTH, 4), dtype=torch.float32).pin_memory()
    host_np = host_img.numpy()
    seen = []
    checksum = [0]

    def on_update(done, total, img):
        seen.append(done)
        checksum[0] += int(img[::97, ::89, :3].sum())  # the host really reads the delivered image
        return False

    r.render_progressive(0, 1, 1, lambda *a: False, preview=True)  # (allocates the preview buffers / second stream: untimed)
    r.resolve_rgba_f32(host_np)                                    # (... and the f32 resolve buffer)
    r.clear()
    seen.clear()
    barrier()
    l0 = r.counters()["kernel_launches"]
    e0 = time.perf_counter()
    r.render_progressive((W * world + rank * K) * F, K * F, F, on_update, preview=True)  # (frame ids disjoint across ranks)
    if world > 1:
        reduce_contexts_(r, dst=0)
    if rank == 0:
        r.resolve_rgba_f32(host_np)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - e0)
    assert len(seen) == K and seen[-1] == K * F
    e2e_value = samples / e2e_s
    e2e_launches = r.counters()["kernel_launches"] - l0
    scene_param_bytes = int(srt.native.lib().srt_launch_param_bytes())  # sizeof(SceneParams), passed by value with every launch
    h2d = int(e2e_launches / K * scene_param_bytes + 32)
    d2h = int(WIDTH * HEIGHT * 4 + host_np.nbytes / K)

    extras = {}
    if not args.no_extras:
        # ---------------- the same workload in the other modes (device-timed, 4 steps each)
        for key, kw in (("exact_math", dict(math=1, rng=0)), ("philox", dict(math=0, rng=1))):
            with srt.Renderer(flat, WIDTH, HEIGHT, max_bounces=BOUNCES, intended_frames=SPP_NAMED, integrator=args.integrator,
                              device=local_rank, **kw) as rr:
                t = timed_render(rr, 4, 1, F, reduce=False)
                extras[key] = {"value": t["samples"] / (t["ms"] * 1e-3), "unit": "samples/s", "steps": 4, "frames_per_step": F,
                               "what": ("SRT_MATH_EXACT: correctly rounded sin / cos / asin, the reference's operation order in the light "
                                        "term -- the mode whose samples equal the oracle's one by one" if key == "exact_math" else
                                        "SRT_RNG_PHILOX: Philox4x32-10 keyed (pixel, frame, bounce), the north star's RNG")}

        # ---------------- every power-of-two spectral width (spectrum.rs:37-38 allows 8..=128), both integrators
        extras["widths"] = {}
        for nl in (8, 16, 32, 64, 128):
            fl = scenes.preset(SCENE, nl)
            row = {}
            for name, integ in (("resident", 1), ("wavefront", 0)):
                with srt.Renderer(fl, WIDTH, HEIGHT, max_bounces=BOUNCES, intended_frames=SPP_NAMED, integrator=integ, device=local_rank) as rr:
                    t = timed_render(rr, 2, 1, 16, reduce=False)
                    row[name] = t["samples"] / (t["ms"] * 1e-3)
            extras["widths"][str(nl)] = row
        extras["widths"]["what"] = ("samples/s of the 1080p Cornell box at n_lambda spectral samples (2 steps x 16 frames per rank, no reduce): the "
                                    "resident integrator exists for every legal width, AUTO picks it for linear-scan scenes")

        # ---------------- BASELINE.json's other configs, every rank its shard of frames, reduce inside the clock
        cfgs = [("C0", "default scene 400x300, 64 iterations (the reference workload)", "default", 0, 400, 300, 64, 64, 4),
                ("C2", "prism dispersion (extension) 1920x1080, 4096 spp", "prism", 0, 1920, 1080, 4096, 32, 3),
                ("C3", "Cornell box 3840x2160, 16384 spp", "cornell", 0, 3840, 2160, 16384, 8, 3),
                ("C4", "10 000 random spheres (BVH) 1920x1080, 1024 spp", "spheres", 10000, 1920, 1080, 1024, 64, 3)]
        extras["configs"] = {}
        for cid, desc, preset, arg, w, h, spp, frames, steps in cfgs:
            fl = scenes.preset(preset, N_LAMBDA, arg)
            with srt.Renderer(fl, w, h, max_bounces=BOUNCES, intended_frames=spp, device=local_rank) as rr:
                t = timed_render(rr, steps, 1, frames)
                sps = t["samples"] / (t["ms"] * 1e-3)
                extras["configs"][cid] = {"config": desc, "samples_per_s": sps, "mrays_per_s": t["rays"] / (t["ms"] * 1e-3) / 1e6,
                                          "frames_per_rank_and_step": frames, "steps": steps, "ms": t["ms"], "reduce_ms": t["reduce_ms"],
                                          "reduce_mbytes": w * h * N_LAMBDA * 4 / 1e6 if world > 1 else 0.0,
                                          "projected_seconds_for_config": spp * w * h / sps}

        # ---------------- the named config as a FIXED job: 1024 spp split over the ranks, reduce + resolve in the clock
        with srt.Renderer(flat, WIDTH, HEIGHT, max_bounces=BOUNCES, intended_frames=SPP_NAMED, integrator=args.integrator,
                          device=local_rank) as rr:
            rr.render_frames(SPP_NAMED, 8)  # warm (first launch, and the context's resolve buffer, outside the clock)
            if rank == 0:
                rr.resolve_rgba_f32(host_np)
            rr.clear()
            first, count = frame_shard(0, SPP_NAMED, rank, world)
            barrier()
            t0 = time.perf_counter()
            done = 0
            dev_ms = 0.0
            while done < count:
                n = min(F, count - done)
                rr.render_frames(first + done, n)
                dev_ms += rr.last_render_stats()[0]
                done += n
            t_render = time.perf_counter() - t0
            if world > 1:
                reduce_contexts_(rr, dst=0)
            t_reduce = time.perf_counter() - t0
            if rank == 0:
                assert rr.frames_accumulated == SPP_NAMED
                rr.resolve_rgba_f32(host_np)
            t_resolve = time.perf_counter() - t0
            barrier()
            strong_s = max_over_ranks(time.perf_counter() - t0)
            phases = torch.tensor([t_render, t_reduce, t_resolve], dtype=torch.float64, device="cuda")
            if world > 1:
                gathered = [torch.zeros_like(phases) for _ in range(world)]
                dist.all_gather(gathered, phases)
            else:
                gathered = [phases]
            extras["strong"] = {"what": f"{SPP_NAMED} spp of the 1080p Cornell box split over {world} rank(s): render in launches of <= {F} "
                                        "frames, NCCL reduce onto rank 0, resolve + read-back of the f32 image there; wall clock between "
                                        "barriers, max over ranks",
                                "seconds": strong_s, "value": SPP_NAMED * npix / strong_s, "unit": "samples/s",
                                "render_ms_max": max_over_ranks(dev_ms), "frames_per_rank": count,
                                "per_rank_s_after_render_reduce_resolve": [[round(float(v), 5) for v in g] for g in gathered]}

        # ---------------- srt_reduce (libsrt_nccl.so: one process, one context per device) against the torch path
        if world > 1:
            barrier()
            dist.barrier(group=host_group)
            if rank == 0:
                try:
                    from spectral_raytracer_b200 import reduce_contexts
                    w2, h2, n2 = 480, 270, 8
                    with srt.Renderer(flat, w2, h2, intended_frames=n2, device=0) as whole:
                        whole.render_frames(0, n2)
                        want = whole.resolve_rgba_f32()
                    ctxs = [srt.Renderer(flat, w2, h2, intended_frames=n2, device=d) for d in range(world)]
                    for d, c in enumerate(ctxs):
                        f0, cnt = frame_shard(0, n2, d, world)
                        if cnt:
                            c.render_frames(f0, cnt)
                    reduce_contexts(ctxs)
                    got = ctxs[0].resolve_rgba_f32()
                    ok = bool(ctxs[0].frames_accumulated == n2 and np.allclose(got, want, rtol=1e-5, atol=1e-6) and
                              all(c.frames_accumulated == 0 for c in ctxs[1:]))
                    for c in ctxs:
                        c.close()
                    # the time of the real message sizes (communicators are cached after the first call)
                    times = {}
                    for label, (w3, h3) in (("1080p_265MB", (1920, 1080)), ("4k_1062MB", (3840, 2160))):
                        ctxs = [srt.Renderer(flat, w3, h3, intended_frames=4, device=d) for d in range(world)]
                        for c in ctxs:
                            c.render_frames(0, 1)
                        reduce_contexts(ctxs)  # creates the communicators
                        best = 1e30
                        for _ in range(3):
                            best = min(best, reduce_contexts(ctxs))
                        times[label] = best
                        for c in ctxs:
                            c.close()
                    extras["reduce_check"] = {"ok": ok, "max_abs_diff": float(np.abs(got - want).max()), "devices": world,
                                              "srt_reduce_ms": times,
                                              "what": "srt_reduce (libsrt_nccl.so, single process, one context per device, cached communicators) "
                                                      "vs one context rendering all frames; then device time of reducing full-size buffers"}
                except Exception as e:  # never lose the headline line over the extra check
                    extras["reduce_check"] = {"ok": False, "error": repr(e)}
            dist.barrier(group=host_group)  # (the other ranks wait here on the host, their GPUs idle)
            barrier()

    # ---------------- converged-image gate + CPU baseline (rank 0)
    cpu = cpu_tight = oc = rmse = None
    if rank == 0:
        gpath = os.path.join(ROOT, "tests", "golden", "converged_cornell_480x270_1024spp.npz")
        if os.path.exists(gpath) and not args.no_extras:
            g = np.load(gpath)
            gw, gh, gspp = int(g["width"]), int(g["height"]), int(g["spp"])
            goc = dict(zip([str(k) for k in g["counter_names"]], [int(v) for v in g["counter_values"]]))
            with srt.Renderer(flat, gw, gh, intended_frames=gspp, max_bounces=BOUNCES, integrator=args.integrator, rng=args.rng,
                              math=args.math, device=local_rank) as rr:
                rr.render_frames(0, gspp)
                got = rr.resolve_rgba_f32()[..., :3].astype(np.float64)
                gc = rr.counters()
            want = g["rgb"].astype(np.float64)
            rmse = {"rel_rmse_vs_cpu": float(np.sqrt(np.mean((got - want) ** 2)) / want.mean()),
                    "mean_ratio": float(got.mean() / want.mean()),
                    "self_hit_rate": gc["self_hits"] / gc["hits"], "cpu_self_hit_rate": goc["self_hits"] / goc["hits"],
                    "thresholds": {"rel_rmse": 0.03, "mean_ratio": 0.002, "self_hit_rate_rel": 0.01},
                    "what": f"Cornell box {gw}x{gh}, {gspp} spp: this run's GPU image (the bench's math / rng mode) against the committed "
                            "image of the CPU reference port with the same pcg3d keys (tests/golden/make_converged_cornell.py); relative RMSE "
                            "of linear RGB, ratio of the means, rate of rounding-level self-intersections (tests/test_gpu_scale_parity.py "
                            "asserts the same thresholds)"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle as O
        threads = O.hardware_threads()
        dt, n, oc, _ = cpu_reference_run(1, threads, counters=True)
        frames = int(min(24, max(1, 12.0 / dt)))
        dt2, n2, _, _ = cpu_reference_run(frames, threads, first_frame=1)
        # the same port without the reference's avoidable cost items (oracle.cpp -DORACLE_TIGHT, bit-identical images)
        frames_t = max(1, frames // 2)
        dt3, n3, _, _ = cpu_reference_run(frames_t, threads, first_frame=1, tight=True)
        cpu_tight = {"value": n3 / dt3, "unit": "samples/s", "cores": threads, "kind": "port",
                     "sample": f"{frames_t} frames ({n3 / 1e6:.1f} M samples, {dt3:.1f} s) of the same workload with the tight build of "
                               "the port: spectra stored 32 wide, colour weights computed once, closest hit tracked without heap "
                               "vector + sort, reciprocals hoisted -- same arithmetic, bit-identical images"}
        cpu = {"value": (n + n2) / (dt + dt2), "unit": "samples/s", "cores": threads, "kind": "port",
               "sample": f"{frames + 1} frames of the same 1920x1080 Cornell box ({(n + n2) / 1e6:.1f} M samples, "
                         f"{dt + dt2:.1f} s), C++ restatement of the Rust reference with its cost structure "
                         "(row-per-task pool, 528-byte spectra, per-ray Vec + sort, per-sample get_rgb_early)"}
    elif rank == 0:
        # event counters for the algorithmic op count only: one small frame of the same scene on the CPU port
        _, _, oc, _ = cpu_reference_run(1, 0, counters=True, width=480, height=270)

    if rank != 0:
        r.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- rooflines
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    kernel = "k_resident" if args.integrator == 1 else "k_shade"
    n_launch = max(1, launches if args.integrator == 1 else stage_n[2])
    kernel_ms = (render_ms if args.integrator == 1 else stage_ms[2]) / n_launch
    samples_per_launch = counters["samples"] / n_launch
    # (1) FP32 issue -- the bound SURVEY.md 8(d) names: algorithmic lane-ops per sample = oracle event counters x cost table
    sm_mhz = clk.get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
    ops = ops_per_sample(oc, N_LAMBDA)
    peak_ops = 148 * 128 * sm_mhz * 1e6
    kernel_sps = samples_per_launch / (kernel_ms * 1e-3)  # rank 0's kernel alone
    traffic_per_sample, traffic_src = newest_ncu_traffic(kernel)
    roofline = {"bound": "fp32_issue", "kernel": kernel, "achieved": kernel_sps * ops / 1e12, "peak": peak_ops / 1e12, "unit": "Tlane-op/s",
                "frac": kernel_sps * ops / peak_ops,
                "traffic": traffic_per_sample * samples_per_launch if traffic_per_sample else None,
                "traffic_source": traffic_src, "traffic_what": "measured DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum of "
                                                               "the newest ncu --set full summary under profiles/, scaled to this launch's samples)",
                "ops_per_sample": ops, "kernel_ms_per_launch": kernel_ms, "samples_per_launch": samples_per_launch,
                "peak_source": f"148 SMs x 128 lanes x SM clock observed during the run ({sm_mhz:.0f} MHz)",
                "note": "SURVEY 8(d) cost table x the CPU port's event counters on the same scene (1 op = one f32 add / mul / min / max / "
                        "compare / select; div, sqrt, rcp and libm calls count 1); achieved = samples/s of the kernel on rank 0, timed with "
                        "CUDA events on its own stream, x ops per sample"}
    # (2) HBM, beside it.  Algorithmic bytes of the dominant kernel from THIS run's counters (rank 0):
    #  * resident integrator: ray and throughput stay on chip; what it must move is the accumulation buffer, larger than L2 --
    #    every pixel record (4*n_lambda B) is read and written once per frame;
    #  * wavefront k_shade: per path-bounce ray 32 B + hit 8 B read, throughput 4*n_lambda read unless fresh; ray +
    #    throughput written if the path goes on; a lit event adds 4*n_lambda to the pixel's record (read + write);
    #  * SURVEY.md 8(d)'s figure for a wavefront whose whole path state lives in HBM is quoted beside it.
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    bounces_rank0 = counters["rays_primary"] + counters["rays_continuation"]
    state_bytes = bounces_rank0 * 2 * (32 + 4 * N_LAMBDA + 4 * N_LAMBDA) + counters["lit"] * 4 * N_LAMBDA
    shade_bytes = (bounces_rank0 * 40 + counters["rays_continuation"] * 4 * N_LAMBDA +
                   counters["rays_continuation"] * (32 + 4 * N_LAMBDA) + counters["lit"] * 2 * 4 * N_LAMBDA)
    accum_bytes = counters["samples"] * 2 * 4 * N_LAMBDA
    alg_bytes = accum_bytes if args.integrator == 1 else shade_bytes
    achieved_gbs = alg_bytes / n_launch / (kernel_ms * 1e-3) / 1e9
    roofline_hbm = {"bound": "hbm", "kernel": kernel, "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                    "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                    "algorithmic_bytes_per_sample": alg_bytes / max(1, counters["samples"]),
                    "measured_dram_bytes_per_sample": traffic_per_sample,
                    "survey_8d_hbm_resident_state_bytes_per_sample": state_bytes / max(1, counters["samples"]),
                    "note": "not the binding roofline: the resident integrator keeps the path state on chip" if args.integrator == 1 else
                            "wavefront k_shade: ray + hit + throughput read, ray + throughput written for surviving paths, 256 B per lit event"}

    line = {
        "metric": "samples/s at 1080p Cornell box", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(),
        "arm": {"frames_per_step": F, "spp_total": F * K * world, "integrator": "resident" if args.integrator == 1 else "wavefront",
                "rng": "pcg3d_reference" if args.rng == 0 else "philox", "math": "fast (CUDA f32 libm)" if args.math == 0 else "exact",
                "parallelism": f"frames sharded over {world} GPU(s), one NCCL reduce of the spectral accumulation buffers"},
        "mrays_per_s": rays / (total_ms * 1e-3) / 1e6,
        "rays_per_sample": rays / samples,
        "reduce_ms": reduce_ms, "wall_s_timed_region": wall,
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "what": "srt_render_progressive: per step F frames + the RGBA8 image of the frames so far (FrameUpdate, main.rs:1343-1348) "
                        "resolved on the device, copied to pinned host memory on a second stream and handed to a host callback; "
                        "at the end srt_resolve_rgba_f32 into pinned host memory (amortised over the steps in d2h)",
                "host_checksum": checksum[0]},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "roofline_hbm": roofline_hbm,
        "cpu_baseline": cpu,
        "cpu_baseline_tight": cpu_tight,
        "rmse": rmse,
    }
    line.update(extras)
    emit(line)
    r.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
