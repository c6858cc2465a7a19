"""Pins the CPU oracle (oracle/oracle.cpp) against every known-answer test the reference holds for
this path (spectrum.rs:772-886, the Hammersley doc comment shader.rs:667-669) and against the
derived vectors of SURVEY.md appendix A.  CPU only."""
import math

import numpy as np
import pytest


def test_wavelength_to_xyz_reference_kat(oracle):
    # test_wavelength_to_XYZ, spectrum.rs:777-804 (tolerance F32_DELTA = 1e-5, shader.rs:7)
    O = oracle
    assert tuple(O.wavelength_to_xyz(379.0)) == (0.0, 0.0, 0.0)
    assert tuple(O.wavelength_to_xyz(781.0)) == (0.0, 0.0, 0.0)
    assert np.allclose(O.wavelength_to_xyz(750.0), (0.000251, 0.000098, 0.0), atol=1e-5)
    assert np.allclose(O.wavelength_to_xyz(702.5), (0.008091, 0.0031415, 0.0), atol=1e-5)
    assert np.allclose(O.wavelength_to_xyz(776.0), (0.0000434, 0.000017, 0.0), atol=1e-5)
    # the reference's lerp has its weights swapped (spectrum.rs:673-680); the test above only passes
    # through its tolerance.  The restatement must reproduce the swapped value, not the intended one.
    assert np.allclose(O.wavelength_to_xyz(776.0), (3.5599962e-05, 1.3999985e-05, 0.0), rtol=1e-6)
    assert tuple(O.wavelength_to_xyz(750.0)) == (np.float32(0.000251), np.float32(0.000098), 0.0)


def test_xyz_to_rgb_matrix_reference_kat(oracle):
    # test_spectrum_to_rgb part (a), spectrum.rs:809-815: D65 white -> (100, 100, 100) +- 0.01
    rgb = oracle.xyz_to_rgb((95.047, 100.0, 108.883))
    assert np.allclose(rgb, (100.0, 100.0, 100.0), atol=0.01)


def test_black_body_reference_kat(oracle):
    # test_black_body_calculation, spectrum.rs:832-869 (relative 1e-4)
    O = oracle
    assert math.isclose(O.black_body(500.0, 5000.0), 12107.190590398, rel_tol=1e-4)
    assert math.isclose(O.black_body(500.0, 1000.0), 1.2134e-6, rel_tol=1e-4)
    assert math.isclose(O.black_body(700.0, 2000.0), 24.390318624, rel_tol=1e-4)
    assert O.black_body(400.0, 500.0) < 1e-10
    # illegal parameters panic in the reference (spectrum.rs:871-885); the oracle signals with NaN
    assert math.isnan(O.black_body(500.0, 0.0)) and math.isnan(O.black_body(-1.0, 300.0))


def test_hammersley_doc_sequence(oracle):
    # shader.rs:667-669
    want = [(0.05, 0.5), (0.15, 0.25), (0.25, 0.75), (0.35, 0.125), (0.45, 0.625), (0.55, 0.375), (0.65, 0.875),
            (0.75, 0.0625), (0.85, 0.5625), (0.95, 0.3125)]
    for n, (x, y) in enumerate(want):
        gx, gy = oracle.hammersley(n, 10)
        assert abs(gx - x) < 1e-6 and gy == y
    assert oracle.hammersley(0, 1) == (0.5, 0.5)  # the jitter-free grid of gate 1


def test_pcg3d_vectors(oracle):
    # SURVEY.md appendix A (float32 emulation of shader.rs:685-705 done during the survey)
    vec = {(0, 0, 0): (2611992518, 2833812075, 1058359340), (0, 0, 30): (3889496412, 197479266, 763805037),
           (1, 2, 3): (4204755366, 1223881804, 1500469937), (959, 539, 31): (3016734871, 459016973, 1192424373),
           (1919, 1079, 1053): (1554658416, 3082447657, 2603914651)}
    for k, raw in vec.items():
        got_raw, f = oracle.pcg3d(*k)
        assert got_raw == raw
        assert np.array_equal(f, (np.array(raw, np.uint32).astype(np.float32) * np.float32(2.0 ** -32)))
        assert (f >= 0).all() and (f <= 1).all()


def test_rgb_loop_drops_samples(oracle):
    # the f32-accumulated `while wavelength <= max` loop of get_rgb_early (spectrum.rs:244-249)
    want = {8: 7, 16: 15, 24: 24, 32: 32, 40: 40, 48: 48, 56: 56, 64: 64, 72: 72, 80: 79, 88: 87, 96: 96, 128: 128}
    for n, c in want.items():
        assert oracle.rgb_loop_count(n) == c


def test_get_rgb_early_vectors(oracle):
    O = oracle
    assert np.allclose(O.get_rgb_early(O.spectrum(O.SPEC_FLAT, 32, 1.0)), (0.31335697, 0.26905552, 0.25139898), rtol=2e-6)
    assert np.allclose(O.get_rgb_early(O.spectrum(O.SPEC_FLAT, 32, 0.7)), (0.21935, 0.18833868, 0.17597929), rtol=2e-6)
    sun = O.spectrum(O.SPEC_TEMPERATURE, 32, 6500.0, 1.0)
    assert np.allclose(sun[[0, 31]], (44516.914, 25657.072), rtol=1e-6)
    assert np.allclose(O.get_rgb_early(sun), (12037.75, 11844.13, 12004.447), rtol=1e-5)
    assert np.allclose(O.get_rgb_early(O.spectrum(O.SPEC_TEMPERATURE, 64, 6500.0, 1.0)), (12604.883, 11952.894, 12413.98),
                       rtol=1e-5)


def test_rgba8_export_rule(oracle):
    # custom_image.rs:92-101: clamp(0,1) * 255, truncating cast, NaN -> 0
    v = np.array([-1.0, 0.0, 0.5, 0.999, 1.0, 7.0, np.nan, 1.0 / 255.0, 254.9 / 255.0], np.float32)
    assert oracle.to_rgba8(v).tolist() == [0, 0, 127, 254, 255, 255, 0, 1, 254]


def test_euler_rotation_is_rz_ry_rx(oracle):
    r = oracle.euler_rotation(0.0, 1.0, 0.0)  # the Cornell box's right front box (main.rs:1612)
    c, s = np.float32(np.cos(np.float32(1.0))), np.float32(np.sin(np.float32(1.0)))
    assert np.allclose(r, [[c, 0, s], [0, 1, 0], [-s, 0, c]], atol=1e-7)
    r = oracle.euler_rotation(0.3, -0.2, 0.7)
    assert np.allclose(r @ r.T, np.eye(3), atol=1e-6) and abs(np.linalg.det(r) - 1) < 1e-6


def test_sampling_directions(oracle):
    O = oracle
    rng = np.random.default_rng(3)
    for _ in range(50):
        n = rng.normal(size=3).astype(np.float32)
        n /= np.linalg.norm(n)
        d = O.cosine_direction(float(rng.random()), float(rng.random()), n)
        assert abs(np.linalg.norm(d) - 1) < 1e-5 and np.dot(d, n) >= -1e-6
        c = O.cone_direction(n, 0.2, float(rng.random()), float(rng.random()))
        assert np.dot(c, n) >= math.cos(0.2 * 0.2 * math.pi / 2) - 1e-5
    # rx = 0 -> straight along the normal; hash floats can be exactly 1.0 -> theta = pi/2 is reachable
    assert np.allclose(O.cosine_direction(0.0, 0.3, (0, 0, 1)), (0, 0, 1), atol=1e-6)
    assert abs(np.dot(O.cosine_direction(1.0, 0.3, (0, 0, 1)), (0, 0, 1))) < 1e-6


def test_oracle_modes_change_only_transcendentals(oracle):
    O = oracle
    sc = O.Scene(32, "cornell")
    a = sc.render(48, 32, 2, threads=2)
    O.set_modes(O.MATH_CANONICAL, O.RNG_PCG3D)
    b = sc.render(48, 32, 2, threads=2)
    O.set_modes(O.MATH_NATIVE, O.RNG_PCG3D)
    c = sc.render(48, 32, 2, threads=2)
    assert np.array_equal(a, c)
    # primary hits and first-bounce direct light do not involve sin/cos/asin: most pixels identical
    assert (np.abs(a - b).max(axis=2) < 1e-6).mean() > 0.3
    assert abs(a.mean() - b.mean()) / a.mean() < 0.05


def test_oracle_scene_statistics(oracle):
    """Event rates of the Cornell box (derived in f64 there; the f32 oracle supersedes)."""
    O = oracle
    sc = O.Scene(32, "cornell")
    O.counters_reset()
    sc.render(96, 54, 4, threads=0)
    c = O.counters()
    s = c["samples"]
    assert s == 96 * 54 * 4 == c["rays_primary"]
    rays = (c["rays_primary"] + c["rays_continuation"] + c["rays_shadow"]) / s
    assert 11.0 < rays < 17.0
    assert 5.0 < c["hits"] / s < 8.0
    assert 0.05 < c["self_hits"] / c["hits"] < 0.25
    assert c["spec_hits"] == 0 and c["rays_shadow"] == c["hits"]


# --------------------------------------------------------------------------- spectrum tooling (8f row f3)
def test_spectrum_resample_restatement(oracle):
    """Spectrum::resample, spectrum.rs:285-323: identity at equal size, end points kept, flat stays flat, and the
    reductions the reference panics on (second trip of the down-sampling loop, spectrum.rs:298; assert of
    linear_interpolate_halved, spectrum.rs:616)."""
    O = oracle
    s = O.spectrum(O.SPEC_TEMPERATURE, 32, 6500.0, 1.0)
    assert np.array_equal(O.spectrum_resample(s, 32), s)
    up = O.spectrum_resample(s, 64)
    assert up[0] == s[0] and up[-1] == s[-1] and np.all(np.diff(up[:8]) > 0)
    # up-sampling: out[i] = I[floor(x)] * (1 - frac) + I[floor(x) + 1] * frac, x = i / (new - 1) * (old - 1)
    x = np.float32(5) / np.float32(63) * np.float32(31)
    lo = int(np.floor(x)); fr = np.float32(x - np.trunc(x))
    assert up[5] == np.float32(s[lo] * (np.float32(1) - fr)) + np.float32(s[lo + 1] * fr)
    flat = O.spectrum(O.SPEC_FLAT, 128, 0.25)
    for n in (32, 64, 96):
        assert np.allclose(O.spectrum_resample(flat, n), 0.25, rtol=1e-6)
    ok = {(a, b): O.spectrum_resample(O.spectrum(O.SPEC_FLAT, a, 1.0), b) is not None
          for a in range(8, 129, 8) for b in range(8, 129, 8)}
    assert all(ok[(a, b)] for a in range(8, 129, 8) for b in range(a, 129, 8))      # up-sampling never panics
    assert ok[(128, 32)] and ok[(32, 8)] and ok[(64, 16)]                           # one collapse, then interpolate
    assert not ok[(128, 24)] and not ok[(128, 8)] and not ok[(64, 8)]               # second trip of the loop


def test_spectrum_radiance_and_normalize_restatement(oracle):
    O = oracle
    flat = O.spectrum(O.SPEC_FLAT, 32, 1.0)
    step = np.float32(400.0) / np.float32(31)
    acc = np.float32(0)
    for _ in range(32):
        acc = np.float32(acc + step)
    assert O.spectrum_radiance(flat) == acc                                          # get_radiance, spectrum.rs:357-362
    for kind, a0, a1 in ((O.SPEC_TEMPERATURE, 6500.0, 1.0), (O.SPEC_TEMPERATURE, 2000.0, 1.0), (O.SPEC_FLAT, 0.7, 0.0)):
        s = O.spectrum(kind, 32, a0, a1)
        n = O.spectrum_normalize(s)
        assert abs(float(O.get_rgb_early(n).max()) - 1.0) < 1e-5                     # spectrum.rs:364-368
        assert np.allclose(n / n[0], s / s[0], rtol=1e-5)                            # the shape is kept


# --------------------------------------------------------------------------- the "tight" CPU build
@pytest.mark.parametrize("name,arg", [("cornell", 0), ("default", 0), ("spheres", 30), ("prism", 0)])
def test_tight_cpu_build_is_bit_identical_to_the_faithful_one(oracle, name, arg):
    """liboracle_tight.so removes the reference's avoidable cost items (128-wide spectra, per-sample colour
    weights, heap vector + sort per ray, double slab test) and nothing else: same images, same spectra, same
    event counters, bit for bit."""
    O = oracle
    w, h, frames = 64, 48, 3
    out = []
    for tight in (False, True):
        sc = O.Scene(32, name, arg, tight=tight)
        img, spec = sc.render(w, h, frames, intended_frames=8, spectral=True, threads=2)
        ids, t, _ = sc.primary(w, h, 0, 8)
        out.append((img, spec, ids, t))
    for a, b in zip(*out):
        assert np.array_equal(a, b, equal_nan=True)
    with pytest.raises(ValueError):
        O.Scene(64, "cornell", tight=True)      # the tight build stores spectra 32 wide


def test_tight_cpu_build_is_faster(oracle):
    import time
    O = oracle
    dt = []
    for tight in (False, True):
        sc = O.Scene(32, "cornell", tight=tight)
        t0 = time.perf_counter()
        sc.render(160, 120, 2, intended_frames=8, threads=2)
        dt.append(time.perf_counter() - t0)
    assert dt[1] < dt[0]
