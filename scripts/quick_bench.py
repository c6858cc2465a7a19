"""Developer timing script (not the contract bench -- see bench.py): renders N frames of a preset
and prints samples/s, rays/s and the per-stage split."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

import spectral_raytracer_b200 as srt  # noqa: E402
from spectral_raytracer_b200 import scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="cornell")
    ap.add_argument("--arg", type=int, default=0)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--pool", type=int, default=0)
    ap.add_argument("--math", type=int, default=0)
    ap.add_argument("--rng", type=int, default=0)
    ap.add_argument("--accel", type=int, default=0)
    ap.add_argument("--integrator", type=int, default=0)
    ap.add_argument("--profile", type=int, default=1)
    a = ap.parse_args()
    flat = scenes.preset(a.scene, 32, a.arg)
    with srt.Renderer(flat, a.width, a.height, intended_frames=1024, pool_paths=a.pool, math=a.math, rng=a.rng,
                      accel=a.accel, integrator=a.integrator) as r:
        r.set_profiling(bool(a.profile))
        r.render_frames(0, 2)  # warm-up
        for rep in range(a.reps):
            r.reset_counters()
            t0 = time.perf_counter()
            r.render_frames(rep * a.frames, a.frames)
            wall = time.perf_counter() - t0
            ms, launches = r.last_render_stats()
            c = r.counters()
            stage_ms, stage_n = r.last_stage_times()
            rays = c["rays_primary"] + c["rays_continuation"] + c["rays_shadow"]
            print(json.dumps({
                "scene": a.scene, "samples_per_s": c["samples"] / (ms * 1e-3), "mrays_per_s": rays / (ms * 1e-3) / 1e6,
                "device_ms": ms, "wall_ms": wall * 1e3, "launches": launches, "iterations": c["iterations"],
                "rays_per_sample": rays / c["samples"], "hits_per_sample": c["hits"] / c["samples"],
                "self_hit_frac": c["self_hits"] / max(1, c["hits"]), "stage_ms": stage_ms, "stage_launches": stage_n,
            }))
        img = r.resolve_rgba_f32()
        print("mean rgb", img[..., :3].reshape(-1, 3).mean(axis=0), "nan px", int(np.isnan(img[..., 0]).sum()))


if __name__ == "__main__":
    main()
