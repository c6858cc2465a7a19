"""checkpoint / resume of the accumulation buffer with the scene (the reference keeps the
image in memory only and lists scene saving as a TODO, main.rs:73)."""
import dataclasses
import os

import numpy as np
import pytest

import spectral_raytracer_b200 as srt
from spectral_raytracer_b200 import scenes

pytestmark = pytest.mark.gpu


def test_interrupted_render_resumed_from_file_equals_uninterrupted_render(tmp_path):
    """Sample-exact mode, one frame per pixel at a time: frames [0, 5) then save, reopen FROM THE FILE ALONE,
    frames [5, 12) -- the buffer equals the uninterrupted render's bit for bit (per pixel the terms of one frame
    are added in path order, and frames are added in the same order)."""
    flat = scenes.preset("cornell", 32)
    w, h, n = 96, 64, 12
    path = tmp_path / "cornell.srtckpt"
    with srt.Renderer(flat, w, h, intended_frames=n, math=srt.MATH_EXACT) as r:
        for f in range(5):
            r.render_frames(f, 1)
        r.save_checkpoint(path)
        assert not os.path.exists(str(path) + ".tmp")
    assert os.path.getsize(path) > w * h * 32 * 4
    with srt.Renderer.open_checkpoint(path) as r2:
        assert (r2.width, r2.height, r2.n_lambda) == (w, h, 32)
        assert r2.frames_accumulated == 5
        for f in range(5, n):
            r2.render_frames(f, 1)
        resumed = r2.read_accum()
        img = r2.resolve_rgba_f32()
    with srt.Renderer(flat, w, h, intended_frames=n, math=srt.MATH_EXACT) as r3:
        for f in range(n):
            r3.render_frames(f, 1)
        assert np.array_equal(r3.read_accum(), resumed, equal_nan=True)
        assert np.array_equal(r3.resolve_rgba_f32(), img, equal_nan=True)


def test_load_into_matching_context_and_rejections(tmp_path):
    flat = scenes.preset("default", 32)
    w, h = 80, 60
    path = tmp_path / "a.srtckpt"
    with srt.Renderer(flat, w, h, intended_frames=16) as r:
        r.render_frames(0, 6)
        acc = r.read_accum()
        r.save_checkpoint(path)
    # same scene, other backend knobs (wavefront integrator, BVH): loads -- they do not change the image
    with srt.Renderer(flat, w, h, intended_frames=16, integrator=srt.INTEGRATOR_WAVEFRONT, accel=srt.ACCEL_BVH) as r:
        r.load_checkpoint(path)
        assert r.frames_accumulated == 6
        assert np.array_equal(r.read_accum(), acc, equal_nan=True)
    # different render constants / scene / size: rejected
    for kw, sc, size in (({"intended_frames": 17}, flat, (w, h)), ({"max_bounces": 5, "intended_frames": 16}, flat, (w, h)),
                         ({"intended_frames": 16}, scenes.preset("cornell", 32), (w, h)),
                         ({"intended_frames": 16}, flat, (w, h + 1))):
        with srt.Renderer(sc, *size, **kw) as r:
            with pytest.raises(srt.SrtError) as e:
                r.load_checkpoint(path)
            assert e.value.code == srt.native.SRT_ERR_INVALID_ARGUMENT
            assert r.frames_accumulated == 0
    moved = dataclasses.replace(flat, lights=flat.lights.copy())
    moved.lights[0, 0] += 0.5
    with srt.Renderer(moved, w, h, intended_frames=16) as r:
        with pytest.raises(srt.SrtError):
            r.load_checkpoint(path)


def test_corrupt_and_truncated_files_are_rejected(tmp_path):
    flat = scenes.preset("cornell", 32)
    path = tmp_path / "c.srtckpt"
    with srt.Renderer(flat, 32, 24, intended_frames=4) as r:
        r.render_frames(0, 2)
        r.save_checkpoint(path)
        blob = bytearray(open(path, "rb").read())
        cases = {"payload bit flip": bytes(blob[:-100]) + bytes([blob[-100] ^ 0x10]) + bytes(blob[-99:]),
                 "scene bit flip": bytes(blob[:200]) + bytes([blob[200] ^ 0x01]) + bytes(blob[201:]),
                 "truncated": bytes(blob[:len(blob) // 2]), "bad magic": b"NOTACKPT" + bytes(blob[8:]), "empty": b""}
        for name, data in cases.items():
            bad = tmp_path / "bad.srtckpt"
            open(bad, "wb").write(data)
            with pytest.raises(srt.SrtError):
                r.load_checkpoint(bad)
            with pytest.raises(srt.SrtError):
                srt.Renderer.open_checkpoint(bad)
        assert r.frames_accumulated == 2            # a failed load leaves the context untouched
        with pytest.raises(srt.SrtError):
            r.load_checkpoint(tmp_path / "missing.srtckpt")
