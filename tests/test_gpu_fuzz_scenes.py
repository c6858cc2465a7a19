"""Randomised-scene parity: the presets exercise a handful of object layouts, these scenes are drawn at random --
overlapping primitives of all three kinds, cameras inside objects, 0..5 lights (light-group boundaries),
metallic / rough / glass materials, several spectral widths -- and every one must reproduce the oracle sample
for sample (SRT_MATH_EXACT vs the oracle's canonical-libm mode), with bit-exact primary-hit ids and distances,
through both integrators and both acceleration structures.  Run on the B200 box with `pytest -m gpu`."""
import numpy as np
import pytest

import spectral_raytracer_b200 as srt
from helpers import flat_from_oracle

pytestmark = pytest.mark.gpu

SAMPLE_RTOL = 2e-5  # as in test_gpu_parity.py: loop vs. nested summation order


def random_scene(O, seed: int, n_lambda: int, glass: bool):
    rng = np.random.default_rng(seed)
    sc = O.Scene(n_lambda)
    v = rng.normal(size=3)
    eye = v / np.linalg.norm(v) * rng.uniform(2.2, 3.5)  # looking at the cluster of objects from any side
    target = rng.uniform(-0.4, 0.4, 3)
    if seed % 5 == 0:
        eye = rng.uniform(-0.3, 0.3, 3)  # camera inside the cluster of objects
        target = eye + rng.uniform(-1.0, 1.0, 3)
    sc.set_camera(eye, target - eye, (0.0, 1.0, 0.0), float(rng.uniform(35.0, 95.0)))
    refl = [sc.add_spectrum(rng.uniform(0.0, 1.0, n_lambda).astype(np.float32)) for _ in range(4)]
    emit = [sc.add_spectrum((rng.uniform(0.0, 1.0, n_lambda) * rng.choice([0.3, 3.0, 40.0])).astype(np.float32))
            for _ in range(3)]
    mats = []
    for _ in range(5):
        mats.append(sc.add_material(float(rng.choice([0.0, 0.0, 0.0, 0.4, 1.0])), float(rng.choice([0.0, 0.0005, 0.15, 0.8])),
                                    int(rng.choice(refl))))
    if glass:
        mats.append(sc.add_glass(int(rng.choice(refl)), float(rng.uniform(1.1, 1.6)), float(rng.uniform(0.0, 9000.0))))
    n_obj = int(rng.integers(1, 11))
    for _ in range(n_obj):
        c = rng.uniform(-1.2, 1.2, 3)
        m = int(rng.choice(mats))
        kind = int(rng.integers(0, 3))
        if kind == 0:
            sc.add_box(c, rng.uniform(0.1, 1.5, 3), m)
        elif kind == 1:
            sc.add_sphere(c, float(rng.uniform(0.1, 0.8)), m)
        else:
            sc.add_rotated_box(c, rng.uniform(0.1, 1.5, 3), rng.uniform(-3.2, 3.2, 3), m)
    if seed % 2 == 0:  # a floor and a back wall, so that paths get long
        sc.add_box((0.0, -2.0, 0.0), (12.0, 0.2, 12.0), mats[0])
        sc.add_box((0.0, 0.0, 3.0), (12.0, 12.0, 0.2), mats[1])
    for _ in range(int(rng.choice([0, 1, 1, 2, 3, 5]))):
        sc.add_light(rng.uniform(-2.5, 2.5, 3), int(rng.choice(emit)))
    return sc


def _one_frame(r, frame):
    r.clear()
    r.render_frames(frame, 1)
    return r.read_accum()


@pytest.mark.parametrize("seed", range(48))
def test_random_scene_matches_oracle_sample_for_sample(oracle, seed):
    O = oracle
    n_lambda = (8, 32, 24, 16, 64, 128, 40, 32, 72, 32, 120, 56)[seed % 12]  # compile-time and guarded-loop widths of the resident kernel
    glass = seed % 4 == 1
    rng_mode = 1 if seed % 7 == 3 else 0
    w, h, N = 64, 48, 8
    sc = random_scene(O, 1000 + seed, n_lambda, glass)
    flat = flat_from_oracle(sc)
    O.set_modes(O.MATH_CANONICAL, rng_mode, (5, seed))
    try:
        want_ids, want_t, _ = sc.primary(w, h, frame=0, intended_frames=1)
        frames = (0, 3)
        want = [sc.render(w, h, 1, first_frame=f, intended_frames=N, spectral=True, threads=4)[1] for f in frames]
    finally:
        O.set_modes(O.MATH_NATIVE, O.RNG_PCG3D)
    for integrator in (srt.INTEGRATOR_WAVEFRONT, srt.INTEGRATOR_RESIDENT):
        for accel in (srt.ACCEL_LINEAR, srt.ACCEL_BVH):
            with srt.Renderer(flat, w, h, intended_frames=1, math=srt.MATH_EXACT, accel=accel, integrator=integrator) as r:
                ids, t = r.primary_ids(0)
            assert np.array_equal(ids, want_ids), (seed, integrator, accel, int((ids != want_ids).sum()))
            assert np.array_equal(t, want_t, equal_nan=True), (seed, integrator, accel)
            with srt.Renderer(flat, w, h, intended_frames=N, math=srt.MATH_EXACT, rng=rng_mode, philox_seed=(5, seed),
                              accel=accel, integrator=integrator, pool_paths=2048) as r:
                for f, wnt in zip(frames, want):
                    got = _one_frame(r, f)
                    assert np.array_equal(np.isnan(got), np.isnan(wnt)), (seed, integrator, accel, f)
                    both_nan = np.isnan(got) & np.isnan(wnt)
                    tol = SAMPLE_RTOL * np.maximum(np.abs(wnt), max(np.nanmax(np.abs(wnt)), 1e-30) * 1e-6)
                    bad = (np.abs(got - wnt) > tol) & ~both_nan
                    assert bad.sum() == 0, (seed, integrator, accel, f, int(bad.sum()), bad.size)


@pytest.mark.parametrize("order", ["plain-rot-plain", "rot-plain-plain", "plain-plain-rot", "sphere-first"])
def test_ties_go_to_the_lowest_original_index(oracle, order):
    """submit_ray sorts the candidates by t with a STABLE sort and takes the first (shader.rs:481-483): coincident
    surfaces of different primitives -- two identical plain boxes, a rotated box with the identity rotation on top
    of them, spheres inside each other touching a box face's plane -- must report the primitive that comes first in
    the caller's list, whatever kind it is (the device scans the kinds in its own order).  Ids and distances bit-exact
    against the oracle through both acceleration structures, per-sample spectra through both integrators."""
    O = oracle
    sc = O.Scene(32)
    sc.set_camera((0.0, 0.25, -3.0), (0.0, -0.05, 1.0), (0.0, 1.0, 0.0), 50.0)
    r = [sc.add_spectrum(np.full(32, v, np.float32)) for v in (0.2, 0.5, 0.9)]
    e = sc.add_spectrum(np.full(32, 5.0, np.float32))
    m = [sc.add_material(0.0, 0.0, s) for s in r]
    c, l = (0.0, 0.0, 0.0), (1.0, 1.0, 1.0)
    add = {"plain": lambda mat: sc.add_box(c, l, mat), "rot": lambda mat: sc.add_rotated_box(c, l, (0.0, 0.0, 0.0), mat),
           "sphere": lambda mat: sc.add_sphere((0.0, 0.0, 0.0), 0.5, mat)}
    kinds = {"plain-rot-plain": ("plain", "rot", "plain"), "rot-plain-plain": ("rot", "plain", "plain"),
             "plain-plain-rot": ("plain", "plain", "rot"), "sphere-first": ("sphere", "sphere", "plain")}[order]
    for k, mat in zip(kinds, m):
        add[k](mat)
    sc.add_box((0.0, -0.6, 0.0), (8.0, 0.2, 8.0), m[1])  # floor touching the boxes' bottom face (y = -0.5)
    sc.add_light((1.5, 2.0, -2.0), e)
    flat = flat_from_oracle(sc)
    w, h, N = 96, 64, 4
    O.set_modes(O.MATH_CANONICAL, O.RNG_PCG3D)
    try:
        want_ids, want_t, _ = sc.primary(w, h, frame=0, intended_frames=1)
        want = sc.render(w, h, 1, first_frame=1, intended_frames=N, spectral=True, threads=4)[1]
    finally:
        O.set_modes(O.MATH_NATIVE, O.RNG_PCG3D)
    assert len(np.unique(want_ids[want_ids >= 0])) >= 2  # the coincident stack and the floor are both in view
    for accel in (srt.ACCEL_LINEAR, srt.ACCEL_BVH):
        with srt.Renderer(flat, w, h, intended_frames=1, math=srt.MATH_EXACT, accel=accel) as rr:
            ids, t = rr.primary_ids(0)
        assert np.array_equal(ids, want_ids), (order, accel, int((ids != want_ids).sum()))
        assert np.array_equal(t, want_t, equal_nan=True)
        for integrator in (srt.INTEGRATOR_WAVEFRONT, srt.INTEGRATOR_RESIDENT):
            with srt.Renderer(flat, w, h, intended_frames=N, math=srt.MATH_EXACT, accel=accel, integrator=integrator) as rr:
                got = _one_frame(rr, 1)
            assert np.array_equal(np.isnan(got), np.isnan(want))
            tol = SAMPLE_RTOL * np.maximum(np.abs(want), max(np.nanmax(np.abs(want)), 1e-30) * 1e-6)
            bad = (np.abs(got - want) > tol) & ~(np.isnan(got) & np.isnan(want))
            assert bad.sum() == 0, (order, accel, integrator, int(bad.sum()))


def random_box_room(O, seed: int, n_lambda: int):
    """Diffuse boxes only (plain and rotated, overlapping, some of them walls) and exactly ONE light: the scenes the
    resident kernel runs in pair mode.  Emission up to 1e20 and reflectances up to 1 so that both accumulate variants
    (tame scene or not) and the last-bounce hand-over are exercised."""
    rng = np.random.default_rng(seed)
    sc = O.Scene(n_lambda)
    v = rng.normal(size=3)
    eye = v / np.linalg.norm(v) * rng.uniform(1.5, 3.0)
    if seed % 3 == 0:
        eye = rng.uniform(-0.3, 0.3, 3)  # inside the cluster
    sc.set_camera(eye, rng.uniform(-0.4, 0.4, 3) - eye, (0.0, 1.0, 0.0), float(rng.uniform(35.0, 95.0)))
    refl = [sc.add_spectrum(rng.uniform(0.0, 1.0, n_lambda).astype(np.float32)) for _ in range(4)]
    mats = [sc.add_material(0.0, float(rng.choice([0.0, 0.3])), int(rng.choice(refl))) for _ in range(4)]
    for _ in range(int(rng.integers(1, 9))):
        c, l, m = rng.uniform(-1.2, 1.2, 3), rng.uniform(0.1, 1.5, 3), int(rng.choice(mats))
        if rng.integers(0, 2):
            sc.add_box(c, l, m)
        else:
            sc.add_rotated_box(c, l, rng.uniform(-3.2, 3.2, 3), m)
    if seed % 2 == 0:  # a closed room: nothing escapes, every path runs into the bounce limit
        for c, l in (((0, -2.5, 0), (6, 0.2, 6)), ((0, 2.5, 0), (6, 0.2, 6)), ((-2.5, 0, 0), (0.2, 6, 6)), ((2.5, 0, 0), (0.2, 6, 6)),
                     ((0, 0, -2.5), (6, 6, 0.2)), ((0, 0, 2.5), (6, 6, 0.2))):
            sc.add_box(c, l, int(rng.choice(mats)))
    scale = float(rng.choice([0.5, 30.0, 1e20]))
    sc.add_light(rng.uniform(-2.0, 2.0, 3), sc.add_spectrum((rng.uniform(0.0, 1.0, n_lambda) * scale).astype(np.float32)))
    return sc


@pytest.mark.parametrize("seed", range(16))
def test_random_one_light_box_rooms_pair_mode(oracle, seed):
    O = oracle
    n_lambda = (32, 8, 24, 64, 32, 128, 16, 40)[seed % 8]
    max_bounces = (30, 4, 1, 7)[seed % 4]
    rng_mode = 1 if seed % 5 == 2 else 0
    w, h, N = 64, 48, 8
    sc = random_box_room(O, 7000 + seed, n_lambda)
    flat = flat_from_oracle(sc)
    O.set_modes(O.MATH_CANONICAL, rng_mode, (9, seed))
    try:
        frames = (0, 6)
        want = [sc.render(w, h, 1, first_frame=f, intended_frames=N, spectral=True, threads=4, max_bounces=max_bounces)[1] for f in frames]
    finally:
        O.set_modes(O.MATH_NATIVE, O.RNG_PCG3D)
    with srt.Renderer(flat, w, h, intended_frames=N, math=srt.MATH_EXACT, rng=rng_mode, philox_seed=(9, seed), max_bounces=max_bounces,
                      integrator=srt.INTEGRATOR_RESIDENT) as r:
        for f, wnt in zip(frames, want):
            got = _one_frame(r, f)
            assert np.array_equal(np.isnan(got), np.isnan(wnt)), (seed, f)
            both_nan = np.isnan(got) & np.isnan(wnt)
            tol = SAMPLE_RTOL * np.maximum(np.abs(wnt), max(np.nanmax(np.abs(wnt)), 1e-30) * 1e-6)
            bad = (np.abs(got - wnt) > tol) & ~both_nan
            assert bad.sum() == 0, (seed, f, int(bad.sum()), bad.size)
