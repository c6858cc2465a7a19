"""Regenerates tests/golden/converged_cornell_480x270_1024spp.npz: BASELINE.json's headline scene (Cornell box,
main.rs:1538-1635) rendered by the CPU oracle (oracle/oracle.cpp: the C++ restatement of the reference, platform
libm, the reference's pcg3d keys) at 480x270, 1024 frames -- the "converged image at a high sample count" the north
star's third correctness gate compares against (SURVEY.md 8d gate 3).  About 133 M samples: a few minutes of CPU.

Stored: the linear f32 RGB image (before clamping), and the oracle's event counters of that very run (hits, self-hits,
rays by class) for the self-hit-rate check.  Run from the repo root: python tests/golden/make_converged_cornell.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
W, H, SPP, N_LAMBDA, BOUNCES = 480, 270, 1024, 32, 30


def main():
    O.build()
    O.set_modes(O.MATH_NATIVE, O.RNG_PCG3D)
    sc = O.Scene(N_LAMBDA, "cornell")
    O.counters_reset()
    t0 = time.time()
    img = sc.render(W, H, SPP, first_frame=0, intended_frames=SPP, max_bounces=BOUNCES, threads=0)
    dt = time.time() - t0
    c = O.counters()
    out = {"rgb": img[..., :3].astype(np.float32), "width": W, "height": H, "spp": SPP, "n_lambda": N_LAMBDA,
           "max_bounces": BOUNCES, "counter_names": np.array(O.COUNTER_NAMES), "counter_values": np.array([c[k] for k in O.COUNTER_NAMES], np.uint64)}
    path = os.path.join(HERE, "converged_cornell_480x270_1024spp.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} in {dt:.0f} s: mean rgb {img[..., :3].reshape(-1, 3).mean(0)}, "
          f"self-hit rate {c['self_hits'] / c['hits']:.5f}, {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
