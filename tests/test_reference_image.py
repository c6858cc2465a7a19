"""The one output of the REAL reference that exists: the image its README publishes (example_image.png, README.md:15,
"Render of the example scene over 1000 iterations", 1920x1080 RGBA8).  tests/golden/reference_example_image.npz
holds its 8x8 block means and 4096 sampled pixels (tests/golden/make_example_fixture.py).  It pins the whole
per-pixel path -- ray generation, submit_ray, intersection, hit / miss shaders, get_rgb_early, the running-mean
blend and the RGBA8 export -- for the oracle (CPU, sampled pixels) and for the CUDA path (GPU, whole image).

Same pcg3d keys on both sides, so the images agree far better than two independent 1000-sample estimates would:
what remains is libm-level path re-rolls and the 8-bit truncation.  Known drift: the published image shows a
perfectly sharp mirror; the default scene of this revision has roughness 0.2 (main.rs:1696).  Both variants are
checked: roughness 0 against the whole image, the current default outside the mirror's silhouette."""
import dataclasses
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = np.load(os.path.join(HERE, "golden", "reference_example_image.npz"))
W, H, FRAMES = 1920, 1080, 1000


def _rgba8(rgb_f32):
    """From<CustomImage> for DynamicImage, custom_image.rs:92-101"""
    a = np.nan_to_num(np.asarray(rgb_f32, np.float32), nan=0.0)
    return (np.clip(a, 0.0, 1.0) * np.float32(255.0)).astype(np.uint8)


def _pixel_stats(got_u8, want_u8):
    d = got_u8.astype(np.int32) - want_u8.astype(np.int32)
    return float(np.abs(d).mean()), float(d.mean()), int(np.abs(d).max()), float((np.abs(d) > 4).mean())


@pytest.mark.parametrize("sharp_mirror", [True, False])
def test_oracle_reproduces_the_published_image_at_sampled_pixels(oracle, sharp_mirror):
    """CPU: the oracle's frame loop for 4096 pixels of the 1920x1080 image, 1000 frames each."""
    O = oracle
    sc = O.Scene(32, "default", 1 if sharp_mirror else 0)
    xy, want = FIX["xy"].astype(np.uint32), FIX["rgb"]
    keep = np.ones(len(xy), bool) if sharp_mirror else ~FIX["mirror"][xy[:, 1] // 8, xy[:, 0] // 8]
    got = _rgba8(sc.render_pixels(W, H, xy[keep], FRAMES)[:, :3])
    mean_abs, mean_signed, worst, frac4 = _pixel_stats(got, want[keep])
    assert keep.sum() > 3000
    assert mean_abs <= 1.0, mean_abs          # measured 0.6 levels
    assert abs(mean_signed) <= 0.1, mean_signed
    assert worst <= 20 and frac4 <= 0.01, (worst, frac4)


@pytest.mark.gpu
@pytest.mark.parametrize("sharp_mirror,integrator", [(True, "resident"), (False, "resident"), (False, "wavefront")])
def test_cuda_path_reproduces_the_published_image(sharp_mirror, integrator):
    """GPU: the whole image through the C ABI (host preset -> srt_create -> 1000 frames -> RGBA8)."""
    import spectral_raytracer_b200 as srt
    from spectral_raytracer_b200 import scenes
    flat = scenes.preset("default", 32)
    if sharp_mirror:
        m = flat.materials.copy()
        assert ((m[:, 0] > 0) & (m[:, 1] == np.float32(0.2))).sum() == 1      # the mirror, main.rs:1696
        m[m[:, 0] > 0, 1] = 0.0
        flat = dataclasses.replace(flat, materials=m)
    integ = srt.INTEGRATOR_RESIDENT if integrator == "resident" else srt.INTEGRATOR_WAVEFRONT
    with srt.Renderer(flat, W, H, intended_frames=FRAMES, integrator=integ) as r:
        r.render_frames(0, FRAMES)
        img = r.resolve_rgba_u8()
    assert (img[..., 3] == 255).all()
    blocks = img[..., :3].reshape(135, 8, 240, 8, 3).astype(np.float64).mean(axis=(1, 3))
    d = blocks - FIX["block_means"].astype(np.float64)
    keep_b = np.ones((135, 240), bool) if sharp_mirror else ~FIX["mirror"]
    assert np.abs(d[keep_b]).mean() <= 0.25, np.abs(d[keep_b]).mean()        # measured 0.085 levels
    assert np.abs(d[keep_b]).max() <= 4.0, np.abs(d[keep_b]).max()          # measured 1.4
    assert abs(d[keep_b].mean()) <= 0.05, d[keep_b].mean()                  # no bias: measured -0.012
    xy, want = FIX["xy"].astype(np.int64), FIX["rgb"]
    keep = keep_b[xy[:, 1] // 8, xy[:, 0] // 8]
    mean_abs, mean_signed, worst, frac4 = _pixel_stats(img[xy[keep, 1], xy[keep, 0], :3], want[keep])
    assert mean_abs <= 1.0 and abs(mean_signed) <= 0.1 and worst <= 20 and frac4 <= 0.01, (mean_abs, mean_signed, worst, frac4)
    if not sharp_mirror:
        # and inside the silhouette the current default really differs from the published image (blurred reflection)
        assert np.abs(d[FIX["mirror"]]).mean() > 1.0
