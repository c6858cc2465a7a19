"""Developer check: the resident kernel's pair mode on/off at every spectral capacity (Cornell box, 1080p).
SRT_RESIDENT_PAIR is read at srt_create.   python scripts/dev_pair_widths.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spectral_raytracer_b200 as srt  # noqa: E402
from spectral_raytracer_b200 import scenes  # noqa: E402

for nl in (8, 16, 24, 32, 64, 128):
    row = []
    for pair in ("0", "1"):
        os.environ["SRT_RESIDENT_PAIR"] = pair
        flat = scenes.preset("cornell", nl)
        with srt.Renderer(flat, 1920, 1080, intended_frames=1024, integrator=srt.INTEGRATOR_RESIDENT) as r:
            r.render_frames(0, 2)
            best = 0.0
            for rep in range(3):
                r.reset_counters()
                r.render_frames(2 + 16 * rep, 16)
                ms, _ = r.last_render_stats()
                best = max(best, r.counters()["samples"] / (ms * 1e-3))
        row.append(best / 1e9)
    print(f"n_lambda {nl:4d}: pair off {row[0]:.3f}  on {row[1]:.3f} G samples/s  ({100 * (row[1] / row[0] - 1):+.1f} %)", flush=True)
