"""Regenerates tests/golden/*.npz from the CPU oracle (oracle/oracle.cpp).

The reference itself (Rust) cannot run in this image, so these fixtures are NOT outputs of the reference
binary; they freeze the oracle -- which tests/test_oracle_kat.py pins against the reference's own known-answer
tests -- so that (a) a silent change of the oracle is caught on CPU and (b) the CUDA path can be checked
against committed numbers without executing the oracle.  Run from the repo root: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    O.build()
    out = {}
    # hashing / jitter
    keys = [(0, 0, 0), (0, 0, 30), (1, 2, 3), (959, 539, 31), (1919, 1079, 1053), (7, 11, 4096)]
    out["pcg3d_keys"] = np.array(keys, np.uint32)
    out["pcg3d_raw"] = np.array([O.pcg3d(*k)[0] for k in keys], np.uint32)
    out["hammersley_64"] = np.array([O.hammersley(n, 64) for n in range(64)], np.float32)
    # spectrum -> RGB on fixed spectra (gate 2)
    for nl in (8, 32, 80, 128):
        S = np.stack([O.spectrum(O.SPEC_FLAT, nl, 1.0), O.spectrum(O.SPEC_FLAT, nl, 0.7), O.spectrum(O.SPEC_RED, nl, 1.0),
                      O.spectrum(O.SPEC_GREEN, nl, 1.0), O.spectrum(O.SPEC_BLUE, nl, 1.0),
                      O.spectrum(O.SPEC_TEMPERATURE, nl, 6500.0, 1.0), O.spectrum(O.SPEC_TEMPERATURE, nl, 2000.0, 1.0)])
        out[f"rgb_spectra_{nl}"] = S
        out[f"rgb_values_{nl}"] = np.stack([O.get_rgb_early(s) for s in S])
    # primary-hit ids on the jitter-free grid (gate 1) and per-sample spectra (canonical libm)
    for name in ("cornell", "default"):
        sc = O.Scene(32, name)
        ids, t, band = sc.primary(64, 36, 0, 1)
        out[f"{name}_ids"], out[f"{name}_t"], out[f"{name}_band"] = ids, t, band
        O.set_modes(O.MATH_CANONICAL, O.RNG_PCG3D)
        _, spec = sc.render(32, 18, 1, first_frame=3, intended_frames=8, spectral=True, threads=2)
        out[f"{name}_frame3_spectra"] = spec.astype(np.float32)
        O.set_modes(O.MATH_NATIVE, O.RNG_PCG3D)
        out[f"{name}_rgba_16spp"] = sc.render(48, 27, 16, threads=2)
    np.savez_compressed(os.path.join(HERE, "oracle_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "oracle_golden.npz"), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
