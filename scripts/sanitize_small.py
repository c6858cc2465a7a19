"""Small renders through every kernel family, for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool memcheck python scripts/sanitize_small.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

import spectral_raytracer_b200 as srt  # noqa: E402
from spectral_raytracer_b200 import scenes  # noqa: E402


def run(name, nl, arg=0, w=64, h=48, frames=2, **kw):
    flat = scenes.preset(name, nl, arg)
    with srt.Renderer(flat, w, h, intended_frames=8, **kw) as r:
        r.render_frames(0, frames)
        r.primary_ids(0)
        img = r.resolve_rgba_f32()
        r.resolve_rgba_u8()
        c = r.counters()
    print(name, nl, arg, kw, "samples", c["samples"], "mean", float(np.nanmean(img[..., :3])), flush=True)


for nl in (8, 24, 32, 72, 128):                       # resident: exact and partial widths; resolve at every tile size
    run("cornell", nl, integrator=srt.INTEGRATOR_RESIDENT)
run("default", 32, integrator=srt.INTEGRATOR_RESIDENT)
run("prism", 16, integrator=srt.INTEGRATOR_RESIDENT, math=srt.MATH_EXACT)
run("cornell", 32, integrator=srt.INTEGRATOR_WAVEFRONT, pool_paths=1024)
run("default", 40, integrator=srt.INTEGRATOR_WAVEFRONT, pool_paths=2048, rng=srt.RNG_PHILOX)
run("spheres", 32, 700, integrator=srt.INTEGRATOR_WAVEFRONT, pool_paths=4096)          # BVH + shadow queue + k_shadow
run("spheres", 32, 700, integrator=srt.INTEGRATOR_WAVEFRONT, math=srt.MATH_EXACT, max_bounces=2)   # last-bounce inline path
run("spheres", 32, 300, integrator=srt.INTEGRATOR_RESIDENT)                            # resident + BVH
flat = scenes.preset("cornell", 32)
with srt.Renderer(flat, 64, 48, intended_frames=8) as r:                               # progressive + preview + deterministic
    r.set_deterministic(True)
    r.render_progressive(0, 6, 2, lambda d, t, img: False, preview=True)
    r.render_frames(6, 2)
print("spectrum tools", srt.spectra_resample(np.ones((3, 32), np.float32), 64).shape, srt.spectra_radiance(np.ones((2, 16), np.float32)),
      srt.spectra_normalize(np.ones((2, 8), np.float32)).shape, srt.spectrum_to_rgb(np.ones((4, 128), np.float32)).shape)
print("sanitize_small done")
