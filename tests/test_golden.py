"""Committed fixtures (tests/golden/oracle_golden.npz, made by tests/golden/make_golden.py): the oracle must
keep reproducing them (CPU), and the CUDA path must match them without the oracle in the loop (GPU)."""
import os

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_golden.npz"))


def test_oracle_reproduces_golden(oracle):
    O = oracle
    for k, raw in zip(G["pcg3d_keys"], G["pcg3d_raw"]):
        assert O.pcg3d(*[int(x) for x in k])[0] == tuple(int(x) for x in raw)
    assert np.array_equal(np.array([O.hammersley(n, 64) for n in range(64)], np.float32), G["hammersley_64"])
    for nl in (8, 32, 80, 128):
        got = np.stack([O.get_rgb_early(s) for s in G[f"rgb_spectra_{nl}"]])
        assert np.array_equal(got, G[f"rgb_values_{nl}"])
    for name in ("cornell", "default"):
        sc = O.Scene(32, name)
        ids, t, band = sc.primary(64, 36, 0, 1)
        assert np.array_equal(ids, G[f"{name}_ids"]) and np.array_equal(t, G[f"{name}_t"])
        O.set_modes(O.MATH_CANONICAL, O.RNG_PCG3D)
        _, spec = sc.render(32, 18, 1, first_frame=3, intended_frames=8, spectral=True, threads=2)
        assert np.array_equal(spec.astype(np.float32), G[f"{name}_frame3_spectra"], equal_nan=True)
        O.set_modes(O.MATH_NATIVE, O.RNG_PCG3D)
        assert np.array_equal(sc.render(48, 27, 16, threads=2), G[f"{name}_rgba_16spp"], equal_nan=True)


@pytest.mark.gpu
@pytest.mark.parametrize("integrator", [0, 1])
def test_cuda_path_matches_golden(integrator):
    import spectral_raytracer_b200 as srt
    from spectral_raytracer_b200 import scenes
    for nl in (8, 32, 80, 128):
        assert np.array_equal(srt.spectrum_to_rgb(G[f"rgb_spectra_{nl}"]), G[f"rgb_values_{nl}"])
    for name in ("cornell", "default"):
        flat = scenes.preset(name, 32)
        with srt.Renderer(flat, 64, 36, intended_frames=1, integrator=integrator) as r:
            ids, t = r.primary_ids(0)
        assert np.array_equal(ids, G[f"{name}_ids"]) and np.array_equal(t, G[f"{name}_t"])
        with srt.Renderer(flat, 32, 18, intended_frames=8, math=srt.MATH_EXACT, integrator=integrator) as r:
            r.render_frames(3, 1)
            got = r.read_accum()
        want = G[f"{name}_frame3_spectra"]
        assert np.array_equal(np.isnan(got), np.isnan(want))
        tol = 2e-5 * np.maximum(np.abs(want), np.nanmax(np.abs(want)) * 1e-6)
        assert (np.abs(got - want) > tol).sum() == 0
        # converged image, production math: statistical agreement with the 16-spp oracle image
        with srt.Renderer(flat, 48, 27, intended_frames=16, integrator=integrator) as r:
            r.render_frames(0, 16)
            img = r.resolve_rgba_f32()
        ref = G[f"{name}_rgba_16spp"]
        ok = np.isfinite(ref[..., 0]) & np.isfinite(img[..., 0])
        assert abs(img[..., :3][ok].mean() - ref[..., :3][ok].mean()) / ref[..., :3][ok].mean() < 0.03


@pytest.mark.gpu
def test_two_shards_on_one_gpu_equal_one_render():
    """Frame sharding + aliasing of the accumulation buffer as a torch tensor, emulated with two contexts on
    one device: shard A + shard B (summed through the aliased tensors) == one context rendering all frames."""
    import torch

    import spectral_raytracer_b200 as srt
    from spectral_raytracer_b200 import scenes
    from spectral_raytracer_b200.distributed import accum_as_tensor, frame_shard
    flat = scenes.preset("cornell", 32)
    w, h, n = 96, 54, 7
    with srt.Renderer(flat, w, h, intended_frames=n, integrator=1) as whole, \
            srt.Renderer(flat, w, h, intended_frames=n, integrator=1) as a, \
            srt.Renderer(flat, w, h, intended_frames=n, integrator=0) as b:
        whole.render_frames(0, n)
        for rank, r in enumerate((a, b)):
            first, count = frame_shard(0, n, rank, 2)
            r.render_frames(first, count)
        ta, tb = accum_as_tensor(a), accum_as_tensor(b)
        torch.cuda.synchronize()
        ta += tb  # what dist.reduce(SUM) does across ranks
        torch.cuda.synchronize()
        a.frames_accumulated = n
        assert np.allclose(a.read_accum(), whole.read_accum(), rtol=1e-5, atol=1e-6)
        assert np.allclose(a.resolve_rgba_f32(), whole.resolve_rgba_f32(), rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_srt_reduce_nccl_two_devices():
    """srt_reduce (libsrt_nccl.so): one process, one context per device, NCCL sum onto the first."""
    import spectral_raytracer_b200 as srt
    from spectral_raytracer_b200 import scenes
    from spectral_raytracer_b200.distributed import frame_shard
    if srt.native.lib().srt_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    flat = scenes.preset("cornell", 32)
    w, h, n = 160, 90, 6
    with srt.Renderer(flat, w, h, intended_frames=n, device=0) as whole, \
            srt.Renderer(flat, w, h, intended_frames=n, device=0) as a, \
            srt.Renderer(flat, w, h, intended_frames=n, device=1) as b:
        whole.render_frames(0, n)
        for rank, r in enumerate((a, b)):
            first, count = frame_shard(0, n, rank, 2)
            r.render_frames(first, count)
        srt.reduce_contexts([a, b])
        assert a.frames_accumulated == n and b.frames_accumulated == 0
        assert not b.read_accum().any()          # the non-root context starts its next shard from an empty image
        assert np.allclose(a.resolve_rgba_f32(), whole.resolve_rgba_f32(), rtol=1e-5, atol=1e-6)
        # a second round (cached communicators): frames n .. 2n, every frame counted once
        whole.render_frames(n, n)
        for rank, r in enumerate((a, b)):
            first, count = frame_shard(n, n, rank, 2)
            r.render_frames(first, count)
        ms = srt.reduce_contexts([a, b])
        assert ms > 0.0
        assert a.frames_accumulated == 2 * n and b.frames_accumulated == 0
        assert np.allclose(a.resolve_rgba_f32(), whole.resolve_rgba_f32(), rtol=1e-5, atol=1e-6)
