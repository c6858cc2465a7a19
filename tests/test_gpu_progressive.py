"""the caller side of the path -- App::render's per-frame protocol (main.rs:1327-1371:
FrameUpdate + RenderingProgressUpdate per frame, AbortRender polled once per frame, TrueTimeUpdate +
DestroySender at the end) mapped onto srt_render_progressive batches with an overlapped preview readback."""
import numpy as np
import pytest

import spectral_raytracer_b200 as srt
from spectral_raytracer_b200 import scenes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("integrator", [srt.INTEGRATOR_WAVEFRONT, srt.INTEGRATOR_RESIDENT])
def test_progressive_updates_previews_and_result(integrator):
    flat = scenes.preset("cornell", 32)
    w, h, n, per = 160, 120, 14, 4
    seen = []
    with srt.Renderer(flat, w, h, intended_frames=n, integrator=integrator) as r:
        def on_update(done, total, img):
            assert img is not None and img.shape == (h, w, 4) and img.dtype == np.uint8
            seen.append((done, total, img.copy()))
            return False
        assert r.render_progressive(0, n, per, on_update, preview=True) is False
        assert [s[:2] for s in seen] == [(4, n), (8, n), (12, n), (14, n)]   # the last batch is ragged
        assert r.frames_accumulated == n
        final8 = r.resolve_rgba_u8()
        final = r.resolve_rgba_f32()
        # the last preview IS the RGBA8 export of the finished image (same accumulation buffer, same kernel)
        assert np.array_equal(seen[-1][2], final8)
        assert (seen[-1][2][..., 3] == 255).all()
        # earlier previews are images of fewer frames: alpha 255, colour not yet equal to the final one
        assert (seen[0][2][..., 3] == 255).all() and not np.array_equal(seen[0][2], final8)
    # same frames through srt_render_frames: same image up to the order of the f32 atomic adds
    with srt.Renderer(flat, w, h, intended_frames=n, integrator=integrator) as r:
        r.render_frames(0, n)
        assert np.allclose(r.resolve_rgba_f32(), final, rtol=1e-5, atol=1e-7)
    # a preview of k frames equals a fresh render of exactly those k frames (RGBA8 truncation: +-1 level)
    with srt.Renderer(flat, w, h, intended_frames=n, integrator=integrator) as r:
        r.render_frames(0, 8)
        ref8 = r.resolve_rgba_u8().astype(np.int16)
    assert np.abs(seen[1][2].astype(np.int16) - ref8).max() <= 1


@pytest.mark.parametrize("integrator", [srt.INTEGRATOR_WAVEFRONT, srt.INTEGRATOR_RESIDENT])
def test_progressive_abort_takes_effect_one_batch_later(integrator):
    flat = scenes.preset("default", 32)
    w, h, n, per = 96, 72, 40, 2
    calls = []
    with srt.Renderer(flat, w, h, intended_frames=n, integrator=integrator) as r:
        def on_update(done, total, img):
            assert img is None       # no preview requested
            calls.append(done)
            return done >= 6         # AbortRender in the third update
        aborted = r.render_progressive(0, n, per, on_update, preview=False)
        assert aborted is True
        # update k's callback runs while batch k+1 is on the GPU: that batch completes and is reported, then it stops
        assert calls == [2, 4, 6, 8]
        assert r.frames_accumulated == 8
        img = r.resolve_rgba_f32()
        assert np.isfinite(img).all() and img[..., :3].mean() > 0
        # the context stays usable: continue where it stopped
        assert r.render_progressive(8, n - 8, 16, None) is False
        assert r.frames_accumulated == n


def test_progressive_without_callback_and_abort_flag():
    flat = scenes.preset("cornell", 32)
    with srt.Renderer(flat, 64, 48, intended_frames=8) as r:
        assert r.render_progressive(0, 8, 3) is False
        assert r.frames_accumulated == 8
        r.abort()                    # srt_abort before the call: nothing is rendered
        assert r.render_progressive(8, 8, 3) is True
        assert r.frames_accumulated == 8
        assert r.render_progressive(8, 8, 0) is False     # frames_per_update 0 = every frame, flag was consumed
        assert r.frames_accumulated == 16


def test_host_render_action_protocol():
    """srt_host::render pushes what App::render pushes (main.rs:1343-1348, :1366-1370)."""
    F, P, T, D = (scenes.ACTION_FRAME_UPDATE, scenes.ACTION_PROGRESS_UPDATE, scenes.ACTION_TRUE_TIME_UPDATE,
                  scenes.ACTION_DESTROY_SENDER)
    n = 6
    kinds, vals, img, acc, done = scenes.render_protocol("default", 80, 60, n, frames_per_update=1)
    assert done and acc == n
    assert kinds == [F, P] * n + [T, D]
    assert np.allclose([vals[2 * k + 1] for k in range(n)], [(k + 1) / n for k in range(n)])   # (frame_number + 1) / N
    assert vals[2 * n] > 0.0                                                                    # TrueTimeUpdate
    assert (img[..., 3] == 255).all() and img[..., :3].max() > 0
    # AbortRender while the second update is pushed: two more frames at most, then the closing actions
    kinds, vals, img, acc, done = scenes.render_protocol("default", 80, 60, 50, frames_per_update=1, abort_at_update=1)
    assert not done and 2 <= acc <= 6
    assert kinds[-2:] == [T, D] and kinds[:-2] == [F, P] * acc
