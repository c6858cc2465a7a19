"""Throughput of every BASELINE.json config on one GPU (a few frames each; the spp of a config only sets how long
it runs).  Writes one JSON object per config; bench.py stays the contract benchmark for the headline config."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import spectral_raytracer_b200 as srt  # noqa: E402
from spectral_raytracer_b200 import scenes  # noqa: E402

CONFIGS = [
    ("C0 default scene 400x300 (reference workload)", "default", 0, 400, 300, 64, 64),
    ("C1 Cornell box 1920x1080", "cornell", 0, 1920, 1080, 1024, 32),
    ("C2 prism (dispersion extension) 1920x1080", "prism", 0, 1920, 1080, 4096, 32),
    ("C3 Cornell box 3840x2160", "cornell", 0, 3840, 2160, 16384, 8),
    ("C4 10k random spheres (BVH) 1920x1080", "spheres", 10000, 1920, 1080, 1024, 64),
]


def main():
    out = []
    only = os.environ.get("SRT_SWEEP_ONLY")  # e.g. "C4" or "C0,C4"
    for name, preset, arg, w, h, spp, frames in CONFIGS:
        if only and name.split()[0] not in only.split(","):
            continue
        flat = scenes.preset(preset, 32, arg)
        with srt.Renderer(flat, w, h, intended_frames=spp) as r:
            r.render_frames(0, 2)
            best = None
            for rep in range(3):
                r.reset_counters()
                r.render_frames(2 + rep * frames, frames)
                ms, launches = r.last_render_stats()
                c = r.counters()
                rays = c["rays_primary"] + c["rays_continuation"] + c["rays_shadow"]
                row = {"config": name, "frames_timed": frames, "spp_of_config": spp, "samples_per_s": c["samples"] / (ms * 1e-3),
                       "mrays_per_s": rays / (ms * 1e-3) / 1e6, "rays_per_sample": rays / c["samples"], "device_ms": ms,
                       "kernel_launches": launches, "projected_seconds_for_config": spp * w * h / (c["samples"] / (ms * 1e-3))}
                if best is None or row["samples_per_s"] > best["samples_per_s"]:
                    best = row
            out.append(best)
            print(json.dumps(best), flush=True)
    return out


if __name__ == "__main__":
    main()
