"""the spectrum tooling on the input side of the path -- Spectrum::resample /
get_radiance / normalize (spectrum.rs:285-374) on the device, bit-exact against the oracle's restatement."""
import numpy as np
import pytest

import spectral_raytracer_b200 as srt
from spectral_raytracer_b200 import scenes

pytestmark = pytest.mark.gpu

SIZES = list(range(8, 129, 8))


def _batch(O, n, rng):
    """preset builders + arbitrary custom curves (positive, negative, tiny, huge)"""
    rows = [O.spectrum(O.SPEC_TEMPERATURE, n, 6500.0, 1.0), O.spectrum(O.SPEC_TEMPERATURE, n, 2000.0, 1e-3),
            O.spectrum(O.SPEC_FLAT, n, 0.7), O.spectrum(O.SPEC_RED, n, 1.0), O.spectrum(O.SPEC_GREEN, n, 0.9),
            O.spectrum(O.SPEC_BLUE, n, 1.0)]
    rows += [rng.standard_normal(n).astype(np.float32) * s for s in (1.0, 1e-20, 1e20)]
    rows += [np.eye(n, dtype=np.float32)[k] for k in (0, n // 2, n - 1)]
    return np.stack(rows).astype(np.float32)


def test_resample_every_size_pair_matches_the_oracle_or_is_rejected_where_the_reference_panics(oracle):
    O = oracle
    rng = np.random.default_rng(7)
    panics = 0
    for n_old in SIZES:
        batch = _batch(O, n_old, rng)
        for n_new in SIZES:
            want = [O.spectrum_resample(row, n_new) for row in batch]
            if want[0] is None:                      # the reference panics for this reduction
                panics += 1
                with pytest.raises(srt.SrtError) as e:
                    srt.spectra_resample(batch, n_new)
                assert e.value.code == srt.native.SRT_ERR_UNSUPPORTED
                continue
            got = srt.spectra_resample(batch, n_new)
            assert got.shape == (len(batch), n_new)
            assert np.array_equal(got, np.stack(want), equal_nan=True), (n_old, n_new)
    assert panics > 0                                 # e.g. 128 -> 24, 64 -> 8


def test_resample_known_cases(oracle):
    O = oracle
    s = O.spectrum(O.SPEC_TEMPERATURE, 32, 6500.0, 1.0)
    assert np.array_equal(srt.spectra_resample(s, 32)[0], s)                  # same size: untouched
    up = srt.spectra_resample(s, 64)[0]
    assert up[0] == s[0] and up[-1] == s[-1]                                 # end points are kept
    down = srt.spectra_resample(s, 8)[0]                                     # collapse to 16, then interpolate to 8
    assert down[0] == s[0] and np.all(np.isfinite(down))
    flat = O.spectrum(O.SPEC_FLAT, 128, 0.25)
    for n in (32, 64, 96, 128):                                              # a flat spectrum stays flat
        assert np.allclose(srt.spectra_resample(flat, n)[0], 0.25, rtol=1e-6)
    with pytest.raises(srt.SrtError) as e:
        srt.spectra_resample(np.zeros((1, 12), np.float32), 8)               # Spectrum::new asserts n % 8 == 0
    assert e.value.code == srt.native.SRT_ERR_SPECTRUM_SAMPLES


@pytest.mark.parametrize("n", [8, 16, 32, 64, 80, 128])
def test_radiance_and_normalize_match_the_oracle(oracle, n):
    O = oracle
    rng = np.random.default_rng(n)
    batch = _batch(O, n, rng)
    rad = srt.spectra_radiance(batch)
    assert np.array_equal(rad, np.array([O.spectrum_radiance(r) for r in batch], np.float32), equal_nan=True)
    nrm = srt.spectra_normalize(batch)
    want = np.stack([O.spectrum_normalize(r) for r in batch])
    assert np.array_equal(nrm, want, equal_nan=True)
    # definition (spectrum.rs:364-368): the normalised spectrum's largest RGB component is 1
    rgb = srt.spectrum_to_rgb(nrm[:2])
    assert np.allclose(rgb.max(axis=1), 1.0, atol=1e-5)


def test_host_mirror_methods_run_on_the_device(oracle):
    O = oracle
    s = O.spectrum(O.SPEC_TEMPERATURE, 32, 6500.0, 1.0)
    assert np.array_equal(scenes.host_spectrum_tool("resample", s, 64), O.spectrum_resample(s, 64))
    assert scenes.host_spectrum_tool("radiance", s) == O.spectrum_radiance(s)
    assert np.array_equal(scenes.host_spectrum_tool("normalize", s), O.spectrum_normalize(s))
    with pytest.raises(ValueError):
        scenes.host_spectrum_tool("resample", O.spectrum(O.SPEC_FLAT, 128, 1.0), 24)   # the reference panics here
