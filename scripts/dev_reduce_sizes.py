"""Developer check (needs >= 2 GPUs): device time of srt_reduce by buffer size and path.  python scripts/dev_reduce_sizes.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spectral_raytracer_b200 as srt  # noqa: E402
from spectral_raytracer_b200 import scenes  # noqa: E402

n_dev = srt.native.lib().srt_device_count()
flat = scenes.preset("cornell", 32)
for mode in ("nccl",):
    pass
    for w, h in ((16, 16), (480, 270), (1920, 1080), (3840, 2160)):
        ctxs = [srt.Renderer(flat, w, h, intended_frames=4, device=d) for d in range(n_dev)]
        for c in ctxs:
            c.render_frames(0, 1)
        srt.reduce_contexts(ctxs)
        best, wall = 1e30, 1e30
        for _ in range(5):
            t0 = time.perf_counter()
            ms = srt.reduce_contexts(ctxs)
            wall = min(wall, (time.perf_counter() - t0) * 1e3)
            best = min(best, ms)
        print(f"{mode} {n_dev} devices {w}x{h} ({w * h * 128 / 1e6:.1f} MB): device {best:.3f} ms, host wall {wall:.3f} ms", flush=True)
        for c in ctxs:
            c.close()
