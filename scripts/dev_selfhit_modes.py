"""Developer check: event counters of the 10 000-sphere scene in exact and production math (same keys), beside the
committed oracle render.  python scripts/dev_selfhit_modes.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spectral_raytracer_b200 as srt  # noqa: E402
from spectral_raytracer_b200 import scenes  # noqa: E402

g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "converged_spheres10k_160x90_256spp.npz"))
oc = dict(zip([str(k) for k in g["counter_names"]], [int(v) for v in g["counter_values"]]))
print("oracle", {k: oc[k] for k in ("hits", "self_hits", "rays_continuation", "rays_shadow", "lit", "spec_hits", "spec_dropped", "misses")})
flat = scenes.preset("spheres", 32, 10000)
imgs = {}
for name, math in (("exact", srt.MATH_EXACT), ("fast", srt.MATH_FAST)):
    with srt.Renderer(flat, 160, 90, intended_frames=256, max_bounces=30, math=math) as r:
        r.render_frames(0, 256)
        c = r.counters()
        imgs[name] = r.resolve_rgba_f32()[..., :3]
    print(name, {k: c[k] for k in ("hits", "self_hits", "rays_continuation", "rays_shadow", "shadow_skipped", "lit", "spec_hits", "spec_dropped", "misses")},
          "self-hit rate", c["self_hits"] / c["hits"])
want = g["rgb"]
for k, v in imgs.items():
    print(k, "mean ratio vs oracle", float(v.mean(dtype=np.float64) / want.mean(dtype=np.float64)),
          "rel rmse", float(np.sqrt(np.mean((v - want) ** 2)) / want.mean()))
