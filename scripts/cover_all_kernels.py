"""Developer helper (GPU box): one tiny invocation of every kernel family through the C ABI (both integrators x
every preset x linear / BVH x both math and RNG modes, run-time spectral widths, small pools, progressive,
checkpoint, spectrum tools) -- a fast crash / launch-error screen after kernel changes, and the command to put
under compute-sanitizer where that is allowed (this pool refuses it).  No oracle here; parity lives in tests/."""
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

import spectral_raytracer_b200 as srt  # noqa: E402
from spectral_raytracer_b200 import scenes  # noqa: E402

W, H, FRAMES = 64, 40, 2


def render(name, arg=0, n_lambda=32, **kw):
    flat = scenes.preset(name, n_lambda, arg)
    with srt.Renderer(flat, W, H, intended_frames=8, **kw) as r:
        r.render_frames(0, FRAMES)
        r.render_frames(FRAMES, 1)
        img = r.resolve_rgba_f32()
        u8 = r.resolve_rgba_u8()
        ids, t = r.primary_ids(0)
        acc = r.read_accum()
        assert img.shape == (H, W, 4) and u8.shape == (H, W, 4) and ids.shape == (H, W)
        return acc


def main():
    done = []
    for integ in (srt.INTEGRATOR_WAVEFRONT, srt.INTEGRATOR_RESIDENT):
        for name, arg in (("cornell", 0), ("default", 0), ("prism", 0), ("spheres", 50)):
            for accel in (srt.ACCEL_LINEAR, srt.ACCEL_BVH):
                for math, rng in ((srt.MATH_FAST, 0), (srt.MATH_EXACT, 1)):
                    render(name, arg, integrator=integ, accel=accel, math=math, rng=rng)
                    done.append((integ, name, accel, math, rng))
    # run-time spectral widths (the NL4 = 0 kernels) and the widest one
    for nl in (8, 80, 128):
        for integ in (srt.INTEGRATOR_WAVEFRONT, srt.INTEGRATOR_RESIDENT):
            render("cornell", 0, nl, integrator=integ)
    # a pool smaller than the image (wavefront regeneration) and a BVH scene above the linear-scan limit
    render("cornell", integrator=srt.INTEGRATOR_WAVEFRONT, pool_paths=1024)
    render("spheres", 500, integrator=srt.INTEGRATOR_WAVEFRONT)
    render("spheres", 500, integrator=srt.INTEGRATOR_RESIDENT)
    # progressive protocol, checkpoint, spectrum tooling, colour conversion, arithmetic self test
    flat = scenes.preset("cornell", 32, 0)
    with srt.Renderer(flat, W, H, intended_frames=8) as r:
        seen = []
        r.render_progressive(0, 4, 2, on_update=lambda done_, total, img: seen.append(done_))
        assert seen, "no FrameUpdate arrived"
        with tempfile.TemporaryDirectory() as d:
            p = os.path.join(d, "c.srt")
            r.save_checkpoint(p)
            with srt.Renderer.open_checkpoint(p) as r2:
                assert np.array_equal(r2.read_accum(), r.read_accum(), equal_nan=True)
    s = np.random.default_rng(0).random((5, 32), dtype=np.float32)
    srt.spectrum_to_rgb(s)
    srt.spectra_resample(s, 64)
    srt.spectra_radiance(s)
    srt.spectra_normalize(s)
    assert srt.selftest_arith(1 << 16, 3) == 0
    print(f"cover_all_kernels: {len(done)} render configurations + tools ok")


if __name__ == "__main__":
    main()
