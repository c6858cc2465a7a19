"""Parity where it is benchmarked (round-2 additions):

  * gate 3 for real: the converged 1024-spp Cornell box, PRODUCTION math, both integrators, against the committed
    oracle image (tests/golden/make_converged_cornell.py) with ABSOLUTE thresholds;
  * BASELINE config C4 at its own size: 10 000 spheres through the BVH against the reference's linear scan
    (shader.rs:468-495), ids / t bit-exact and per-sample spectra in exact-math mode;
  * the resident integrator at every legal spectral width (spectrum.rs:37-38);
  * srt_abort from a second thread in the middle of a wavefront call; deterministic accumulation.
"""
import os
import threading
import time

import numpy as np
import pytest

import spectral_raytracer_b200 as srt
from spectral_raytracer_b200 import scenes
from helpers import flat_from_oracle, rel_rmse

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SAMPLE_RTOL = 2e-5  # see test_gpu_parity.py


# --------------------------------------------------------------------------- gate 3, absolute thresholds
# Stated thresholds (SURVEY.md 8d gate 3) for 480x270, 1024 spp, production math vs the oracle's CPU render with the
# reference's pcg3d keys:
#   relative RMSE of linear RGB  <= 0.03   (two INDEPENDENT 1024-spp estimates of this scene differ by ~0.12; the
#                                            same keys with a different libm re-roll ~1.5 % of the paths)
#   |mean(a) / mean(b) - 1|      <= 0.002  (the bias check that catches a wrong self-intersection rate)
#   self-hit rate                within 1 % relative of the oracle's (12.5 % of all Cornell hits are rounding-level
#                                            self-intersections that are part of the reference's image)
CONVERGED_REL_RMSE, CONVERGED_MEAN, CONVERGED_SELF_HIT = 0.03, 0.002, 0.01


@pytest.mark.parametrize("integrator", [srt.INTEGRATOR_WAVEFRONT, srt.INTEGRATOR_RESIDENT])
def test_converged_cornell_1024spp_production_math(integrator):
    g = np.load(os.path.join(GOLDEN, "converged_cornell_480x270_1024spp.npz"))
    w, h, spp = int(g["width"]), int(g["height"]), int(g["spp"])
    want = g["rgb"]
    oc = dict(zip([str(k) for k in g["counter_names"]], [int(v) for v in g["counter_values"]]))
    flat = scenes.preset("cornell", int(g["n_lambda"]))
    with srt.Renderer(flat, w, h, intended_frames=spp, max_bounces=int(g["max_bounces"]), integrator=integrator) as r:
        r.render_frames(0, spp)
        got = r.resolve_rgba_f32()[..., :3]
        c = r.counters()
    assert c["samples"] == w * h * spp == oc["samples"]
    assert np.isfinite(got).all() and np.isfinite(want).all()
    stats = {"rel_rmse": rel_rmse(got, want), "mean_ratio": float(got.mean(dtype=np.float64) / want.mean(dtype=np.float64)),
             "self_hit_rate": c["self_hits"] / c["hits"], "oracle_self_hit_rate": oc["self_hits"] / oc["hits"]}
    print("converged gate:", stats)
    assert stats["rel_rmse"] <= CONVERGED_REL_RMSE, stats
    assert abs(stats["mean_ratio"] - 1.0) <= CONVERGED_MEAN, stats
    assert abs(stats["self_hit_rate"] / stats["oracle_self_hit_rate"] - 1.0) <= CONVERGED_SELF_HIT, stats
    # the path-length statistics behind the throughput numbers are the reference's too
    assert abs(c["hits"] / oc["hits"] - 1.0) <= 0.002
    assert abs(c["rays_continuation"] / oc["rays_continuation"] - 1.0) <= 0.002


# --------------------------------------------------------------------------- C4 at its own size
@pytest.fixture(scope="module")
def spheres_10k(oracle):
    return oracle.Scene(32, "spheres", 10000)


# The same gate for BASELINE's config C4 in the mode it is benchmarked in: production math, BVH, queued shadow rays
# (k_shade -> k_shadow), against the oracle's render through the reference's linear scan over all 10 001 objects
# (tests/golden/make_converged_spheres.py: 160x90, 256 spp, pcg3d keys).
SPHERES_REL_RMSE, SPHERES_MEAN = 0.03, 0.002


@pytest.mark.parametrize("integrator", [srt.INTEGRATOR_WAVEFRONT, srt.INTEGRATOR_RESIDENT])
def test_converged_10k_spheres_256spp_production_math(oracle, spheres_10k, integrator):
    g = np.load(os.path.join(GOLDEN, "converged_spheres10k_160x90_256spp.npz"))
    w, h, spp = int(g["width"]), int(g["height"]), int(g["spp"])
    want = g["rgb"]
    oc = dict(zip([str(k) for k in g["counter_names"]], [int(v) for v in g["counter_values"]]))
    assert int(g["n_spheres"]) == 10000 and int(g["n_lambda"]) == 32
    with srt.Renderer(flat_from_oracle(spheres_10k), w, h, intended_frames=spp, max_bounces=int(g["max_bounces"]), integrator=integrator) as r:
        r.render_frames(0, spp)
        got = r.resolve_rgba_f32()[..., :3]
        c = r.counters()
    assert c["samples"] == w * h * spp == oc["samples"]
    assert np.isfinite(got).all() and np.isfinite(want).all()
    stats = {"rel_rmse": rel_rmse(got, want), "mean_ratio": float(got.mean(dtype=np.float64) / want.mean(dtype=np.float64))}
    print("converged gate, 10 000 spheres:", stats)
    assert stats["rel_rmse"] <= SPHERES_REL_RMSE, stats          # measured 0.010 (exact math: 0.005)
    assert abs(stats["mean_ratio"] - 1.0) <= SPHERES_MEAN, stats  # measured 1.00004
    # Half of this scene's materials are metals: the oracle's recursion also traces -- and counts -- the subtree of a
    # specular child that its parent then discards (shader.rs:407), the loop here stops at the discard.  So only the
    # primaries are equal, everything else is bounded by the oracle's count (see test_continuation_rays_match_oracle_loop).
    assert c["rays_primary"] == oc["rays_primary"]
    for k in ("rays_continuation", "hits", "self_hits", "spec_hits", "spec_dropped"):
        assert (0.4 if k == "spec_dropped" else 0.8) * oc[k] <= c[k] <= oc[k], (k, c[k], oc[k])   # measured 0.54 / 0.86 .. 0.96
    assert c["rays_shadow"] + c["shadow_skipped"] <= oc["rays_shadow"]


def test_bvh_10k_spheres_primary_ids_and_t_bit_exact(oracle, spheres_10k):
    """10 001 objects: depth of the tree, the 64-entry traversal stack, leaves of up to 4 primitives and the padded
    distance culling, against the linear scan over all objects with the stable lowest-index tie-break."""
    sc = spheres_10k
    w, h = 192, 108
    with srt.Renderer(flat_from_oracle(sc), w, h, intended_frames=64) as r:
        for frame, n in ((0, 1), (37, 64)):
            want_ids, want_t, _ = sc.primary(w, h, frame=frame, intended_frames=n)
            if n != 64:
                with srt.Renderer(flat_from_oracle(sc), w, h, intended_frames=n) as r1:
                    ids, t = r1.primary_ids(frame)
            else:
                ids, t = r.primary_ids(frame)
            assert np.array_equal(ids, want_ids)
            assert np.array_equal(t, want_t)
            assert len(np.unique(ids)) > 1000      # the camera really sees thousands of different spheres


@pytest.mark.parametrize("integrator", [srt.INTEGRATOR_WAVEFRONT, srt.INTEGRATOR_RESIDENT])
def test_bvh_10k_spheres_per_sample_spectra(oracle, spheres_10k, integrator):
    """Whole paths (closest hits, shadow rays through the BVH's any-hit traversal, specular gates) sample for sample."""
    O = oracle
    sc = spheres_10k
    w, h, N = 96, 54, 8
    O.set_modes(O.MATH_CANONICAL, O.RNG_PCG3D)
    with srt.Renderer(flat_from_oracle(sc), w, h, intended_frames=N, math=srt.MATH_EXACT, integrator=integrator) as r:
        for frame in (0, 3):
            O.counters_reset()
            _, want = sc.render(w, h, 1, first_frame=frame, intended_frames=N, spectral=True, threads=0)
            oc = O.counters()
            r.clear()
            r.reset_counters()
            r.render_frames(frame, 1)
            got = r.read_accum()
            gc = r.counters()
            assert gc["samples"] == oc["samples"] and gc["rays_primary"] == oc["rays_primary"]
            assert np.array_equal(np.isnan(got), np.isnan(want))
            tol = SAMPLE_RTOL * np.maximum(np.abs(want), np.nanmax(np.abs(want)) * 1e-6)
            bad = (np.abs(got - want) > tol) & ~np.isnan(want)
            assert bad.sum() == 0, f"{bad.sum()} of {bad.size} spectral samples differ"


# --------------------------------------------------------------------------- every legal spectral width
@pytest.mark.parametrize("integrator", [srt.INTEGRATOR_WAVEFRONT, srt.INTEGRATOR_RESIDENT])
@pytest.mark.parametrize("name,n_lambda", [("cornell", 8), ("cornell", 16), ("cornell", 24), ("cornell", 40), ("cornell", 64),
                                           ("cornell", 72), ("cornell", 120), ("cornell", 128), ("default", 16), ("default", 56),
                                           ("default", 128), ("prism", 8), ("prism", 64)])
def test_every_spectral_width_both_integrators(oracle, name, n_lambda, integrator):
    """Spectrum::new allows every multiple of 8 up to 128 (spectrum.rs:37-38; UI guard main.rs:662-691).  The resident
    integrator has compile-time loops for 8/16/32/64/128 and guarded loops for the widths in between."""
    O = oracle
    w, h, N = 64, 48, 4
    sc = O.Scene(n_lambda, name)
    O.set_modes(O.MATH_CANONICAL, O.RNG_PCG3D)
    with srt.Renderer(flat_from_oracle(sc), w, h, intended_frames=N, math=srt.MATH_EXACT, integrator=integrator) as r:
        _, want = sc.render(w, h, 1, first_frame=1, intended_frames=N, spectral=True, threads=4)
        r.render_frames(1, 1)
        got = r.read_accum()
        rgb = r.resolve_rgba_f32()
    assert got.shape == (h, w, n_lambda)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    tol = SAMPLE_RTOL * np.maximum(np.abs(want), np.nanmax(np.abs(want)) * 1e-6)
    assert ((np.abs(got - want) > tol) & ~np.isnan(want)).sum() == 0
    assert np.isfinite(rgb[..., :3]).all() or name != "cornell"


def test_resident_many_objects_occupancy_sized_grid():
    """64 staged primitives need 7 KB more shared memory per block than the presets: the persistent grid is sized by
    the occupancy calculator, the image must equal the wavefront's."""
    rng = np.random.default_rng(3)
    n = 64
    objs = np.zeros((n, 23), np.float32)
    c = rng.uniform(-2, 2, (n, 3)).astype(np.float32) + np.array([0, 0, 4], np.float32)
    rad = rng.uniform(0.1, 0.4, n).astype(np.float32)
    objs[:, 0:3], objs[:, 3:6], objs[:, 6] = c - rad[:, None], c + rad[:, None], srt.native.SPHERE
    objs[:, 22] = rng.integers(0, 2, n)
    spectra = np.stack([np.full(32, 0.7, np.float32), np.linspace(0.2, 0.9, 32, dtype=np.float32), np.full(32, 5.0, np.float32)])
    flat = srt.FlatScene(32, np.array([0, 0, -2, 0, 0, 1, 0, 1, 0, 60], np.float32), objs,
                         np.array([[0, 0, 0, 0, 1, 0], [0.5, 0.1, 1, 0, 1, 0]], np.float32),
                         np.array([[0, 3, 2, 2], [2, 1, 0, 2]], np.float32), spectra)
    out = []
    for integ in (srt.INTEGRATOR_RESIDENT, srt.INTEGRATOR_WAVEFRONT):
        with srt.Renderer(flat, 96, 64, intended_frames=4, math=srt.MATH_EXACT, integrator=integ) as r:
            r.render_frames(2, 1)
            out.append((r.read_accum(), r.counters()))
    (a, ca), (b, cb) = out
    for k in ("hits", "self_hits", "rays_shadow", "lit", "spec_hits"):
        assert ca[k] == cb[k]
    assert np.allclose(a, b, rtol=2e-5, atol=1e-7, equal_nan=True)


# --------------------------------------------------------------------------- abort in the middle of a call
def test_wavefront_abort_from_second_thread_leaves_whole_frames():
    """srt_abort() while srt_render_frames is running (wavefront: the host drives the iterations): the call stops at
    a frame boundary -- the buffer holds exactly srt_frames_accumulated whole frames, nothing of a partial one."""
    flat = scenes.preset("cornell", 32)
    w, h, n = 1920, 1080, 400          # ~0.9 s of wavefront rendering
    with srt.Renderer(flat, w, h, intended_frames=n, integrator=srt.INTEGRATOR_WAVEFRONT) as r:
        r.render_frames(n, 1)          # (first-launch costs out of the way)
        r.clear()
        r.reset_counters()
        result = {}

        def work():
            try:
                r.render_frames(0, n)
                result["rc"] = "completed"
            except srt.SrtError as e:
                result["rc"] = e.code

        t = threading.Thread(target=work)
        t.start()
        time.sleep(0.15)
        r.abort()
        t.join(60)
        assert not t.is_alive()
        assert result["rc"] == srt.native.SRT_ERR_ABORTED, result   # (400 frames of 1920x1080 take well over 0.15 s)
        done = r.frames_accumulated
        assert 0 < done < n
        c = r.counters()
        assert c["samples"] == done * w * h          # whole frames, and every started path was traced to its end
        img = r.resolve_rgba_f32()
        # the same frames rendered without interruption: the same image up to the order of the f32 adds
        with srt.Renderer(flat, w, h, intended_frames=n, integrator=srt.INTEGRATOR_WAVEFRONT) as ref:
            ref.render_frames(0, done)
            want = ref.resolve_rgba_f32()
        assert np.allclose(img, want, rtol=1e-4, atol=1e-6)
        # the context stays usable
        r.render_frames(done, 2)
        assert r.frames_accumulated == done + 2


def test_progressive_abort_inside_batch_delivers_the_partial_update():
    flat = scenes.preset("cornell", 32)
    w, h, n, per = 1920, 1080, 400, 200
    with srt.Renderer(flat, w, h, intended_frames=n, integrator=srt.INTEGRATOR_WAVEFRONT) as r:
        r.render_frames(n, 1)
        r.clear()
        r.reset_counters()
        seen = []
        timer = threading.Timer(0.15, r.abort)
        timer.start()
        aborted = r.render_progressive(0, n, per, lambda done, total, img: seen.append(done) or False)
        timer.cancel()
        assert aborted is True
        assert seen and seen[-1] == r.frames_accumulated and 0 < seen[-1] < n
        assert r.counters()["samples"] == r.frames_accumulated * w * h


# --------------------------------------------------------------------------- deterministic accumulation
@pytest.mark.parametrize("integrator", [srt.INTEGRATOR_WAVEFRONT, srt.INTEGRATOR_RESIDENT])
def test_deterministic_mode_is_bit_reproducible(integrator):
    flat = scenes.preset("cornell", 32)
    w, h, n = 320, 180, 12
    bufs = []
    for split in ((12,), (5, 7), (1,) * 12):
        with srt.Renderer(flat, w, h, intended_frames=n, integrator=integrator) as r:
            r.set_deterministic(True)
            first = 0
            for k in split:
                r.render_frames(first, k)
                first += k
            assert r.frames_accumulated == n
            bufs.append(r.read_accum())
    assert np.array_equal(bufs[0], bufs[1]) and np.array_equal(bufs[0], bufs[2])
    # and it is the same estimate as the default mode, up to the order of the adds
    with srt.Renderer(flat, w, h, intended_frames=n, integrator=integrator) as r:
        r.render_frames(0, n)
        assert np.allclose(r.read_accum(), bufs[0], rtol=1e-4, atol=1e-6)


# --------------------------------------------------------------------------- BVH wavefront: queued shadow rays
@pytest.mark.parametrize("max_bounces", [1, 6])          # 1: every hit is a path's last bounce (its throughput stays in its own pool)
@pytest.mark.parametrize("math", [srt.MATH_EXACT, srt.MATH_FAST])
def test_shadow_kernel_equals_in_place_shadow_rays(oracle, monkeypatch, math, max_bounces):
    """Large BVH scenes trace their shadow rays in a kernel of their own (k_shade queues them per light, k_shadow
    traces the queues in light order); the rays, their outcome and the terms are the same as with the rays traced in
    place -- only the f32 association of several lights of one hit differs (per light instead of per pair)."""
    sc = oracle.Scene(32, "spheres", 700)
    flat = flat_from_oracle(sc)
    w, h, n = 160, 90, 3
    out = []
    for flag in ("0", "1"):
        monkeypatch.setenv("SRT_SHADOW_KERNEL", flag)
        with srt.Renderer(flat, w, h, intended_frames=8, math=math, integrator=srt.INTEGRATOR_WAVEFRONT, max_bounces=max_bounces) as r:
            for f in range(n):           # one frame per call: a pixel's terms are added in path order
                r.render_frames(f, 1)
            out.append((r.read_accum(), r.counters()))
    (a, ca), (b, cb) = out
    for k in ("samples", "rays_primary", "rays_continuation", "rays_shadow", "shadow_skipped", "hits", "self_hits", "lit", "spec_hits",
              "spec_dropped", "misses"):
        assert ca[k] == cb[k], (k, ca[k], cb[k])
    assert cb["kernel_launches"] > ca["kernel_launches"]          # (the queued rays really went through k_shadow)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    assert np.allclose(a, b, rtol=1e-5, atol=1e-9, equal_nan=True)
    # deterministic: the same render twice gives the same bits
    monkeypatch.setenv("SRT_SHADOW_KERNEL", "1")
    with srt.Renderer(flat, w, h, intended_frames=8, math=math, integrator=srt.INTEGRATOR_WAVEFRONT, max_bounces=max_bounces) as r:
        for f in range(n):
            r.render_frames(f, 1)
        assert np.array_equal(r.read_accum(), b, equal_nan=True)


# --------------------------------------------------------------------------- C4 at BASELINE's resolution: properties
def test_c4_full_hd_properties(oracle, spheres_10k):
    """10 000 spheres at 1920x1080 (the size the bench's `configs.C4` runs at), production math, the integrator AUTO
    picks (wavefront: BVH, shadow-ray queues, k_shadow): the image and every event counter are independent of how the
    frame range is split into calls and of the size of the path pool (64 Ki paths: hundreds of refill iterations per
    frame instead of one), and the deterministic mode is bit-reproducible."""
    flat = flat_from_oracle(spheres_10k)
    w, h, n = 1920, 1080, 4
    keys = ("samples", "rays_primary", "rays_continuation", "rays_shadow", "shadow_skipped", "hits", "self_hits", "misses", "lit",
            "spec_hits", "spec_dropped")
    runs = []
    for pool, split in ((0, (4,)), (0, (1, 3)), (1 << 16, (2, 2))):
        with srt.Renderer(flat, w, h, intended_frames=1024, pool_paths=pool) as r:
            first = 0
            for k in split:
                r.render_frames(first, k)
                first += k
            assert r.frames_accumulated == n
            runs.append((r.resolve_rgba_f32(), r.counters(), r.read_accum()))
    img0, c0, acc0 = runs[0]
    assert c0["samples"] == n * w * h == c0["rays_primary"]
    assert np.isfinite(img0).all() and 0.01 < img0[..., :3].mean() < 2.0
    for img, c, acc in runs[1:]:
        for k in keys:
            assert c[k] == c0[k], (k, c[k], c0[k])
        # the same non-negative radiance terms added in another order: compared where nothing cancels, in the spectra
        # (the XYZ -> RGB matrix has negative entries; saturated colours amplify an ulp of the spectrum in one channel)
        assert np.allclose(acc, acc0, rtol=1e-5, atol=1e-7 * float(np.nanmax(acc0)), equal_nan=True)
        assert np.abs(img - img0).max() <= 1e-4 * max(1.0, float(img0.max()))
    bits = []
    for _ in range(2):
        with srt.Renderer(flat, w, h, intended_frames=1024) as r:
            r.set_deterministic(True)
            r.render_frames(0, 2)
            bits.append(r.read_accum())
    assert np.array_equal(bits[0], bits[1], equal_nan=True)


# --------------------------------------------------------------------------- resident kernel, pair mode
@pytest.mark.parametrize("max_bounces", [1, 2, 3, 30])
@pytest.mark.parametrize("w,h", [(5, 3), (64, 48)])
def test_pair_mode_edges_sample_exact(oracle, w, h, max_bounces):
    """One-light diffuse scenes run the resident kernel's pair mode (a lane's shadow ray and its next path ray share a
    scan; the last shadow ray of a path rides along with the next sample's primary ray).  Paths of 1..3 bounces -- every
    hit the path's last -- and images smaller than a warp, sample for sample and event for event against the oracle."""
    O = oracle
    sc = O.Scene(32, "cornell")
    O.set_modes(O.MATH_CANONICAL, O.RNG_PCG3D)
    with srt.Renderer(flat_from_oracle(sc), w, h, intended_frames=8, math=srt.MATH_EXACT, integrator=srt.INTEGRATOR_RESIDENT,
                      max_bounces=max_bounces) as r:
        for frame in (0, 5):
            O.counters_reset()
            _, want = sc.render(w, h, 1, first_frame=frame, intended_frames=8, spectral=True, threads=4, max_bounces=max_bounces)
            oc = O.counters()
            r.clear()
            r.reset_counters()
            r.render_frames(frame, 1)
            got = r.read_accum()
            gc = r.counters()
            for k in ("samples", "rays_primary", "rays_continuation", "hits", "self_hits"):
                assert gc[k] == oc[k], (k, gc[k], oc[k])
            assert gc["rays_shadow"] + gc["shadow_skipped"] == oc["rays_shadow"]
            assert np.array_equal(np.isnan(got), np.isnan(want))
            tol = SAMPLE_RTOL * np.maximum(np.abs(want), np.nanmax(np.abs(want)) * 1e-6)
            assert ((np.abs(got - want) > tol) & ~np.isnan(want)).sum() == 0


def test_pair_mode_equals_pass_loop(monkeypatch):
    """The same render through the pass-loop kernel (SRT_RESIDENT_PAIR=0, read by srt_create) and the pair-mode kernel:
    the same paths -- every event counter equal -- and the same radiance terms, added in another order."""
    flat = scenes.preset("cornell", 32)
    w, h, n = 640, 360, 6
    out = []
    for flag in ("0", "1"):
        monkeypatch.setenv("SRT_RESIDENT_PAIR", flag)
        with srt.Renderer(flat, w, h, intended_frames=64, integrator=srt.INTEGRATOR_RESIDENT) as r:
            r.render_frames(3, n)
            out.append((r.read_accum(), r.counters()))
    (a, ca), (b, cb) = out
    for k in ("samples", "rays_primary", "rays_continuation", "rays_shadow", "shadow_skipped", "hits", "self_hits", "misses", "lit"):
        assert ca[k] == cb[k], (k, ca[k], cb[k])
    assert np.allclose(a, b, rtol=1e-5, atol=1e-7 * float(np.nanmax(a)), equal_nan=True)
