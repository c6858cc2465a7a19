"""Regenerates tests/golden/converged_spheres10k_160x90_256spp.npz: BASELINE.json's config C4 -- the floor plus 10 000
random spheres generated with the reference's own random_pcg3d -- rendered by the CPU oracle with the
reference's acceleration structure, i.e. the linear scan over all 10 001 objects for every ray (shader.rs:468-495),
platform libm, pcg3d keys, 160x90, 256 frames.  The GPU path renders the same scene through its BVH, the shadow-ray
queues and k_shadow in production math; tests/test_gpu_scale_parity.py compares the two with absolute thresholds.
About 1.4e11 slab tests: a minute or two of CPU.  Run from the repo root: python tests/golden/make_converged_spheres.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
W, H, SPP, N_LAMBDA, BOUNCES, N_SPHERES = 160, 90, 256, 32, 30, 10000


def main():
    O.build()
    O.set_modes(O.MATH_NATIVE, O.RNG_PCG3D)
    sc = O.Scene(N_LAMBDA, "spheres", N_SPHERES)
    O.counters_reset()
    t0 = time.time()
    img = sc.render(W, H, SPP, first_frame=0, intended_frames=SPP, max_bounces=BOUNCES, threads=0)
    dt = time.time() - t0
    c = O.counters()
    path = os.path.join(HERE, "converged_spheres10k_160x90_256spp.npz")
    np.savez_compressed(path, rgb=img[..., :3].astype(np.float32), width=W, height=H, spp=SPP, n_lambda=N_LAMBDA, max_bounces=BOUNCES,
                        n_spheres=N_SPHERES, counter_names=np.array(O.COUNTER_NAMES),
                        counter_values=np.array([c[k] for k in O.COUNTER_NAMES], np.uint64))
    print(f"wrote {path} in {dt:.0f} s: mean rgb {img[..., :3].reshape(-1, 3).mean(0)}, self-hit rate {c['self_hits'] / c['hits']:.5f}, "
          f"rays/sample {(c['rays_primary'] + c['rays_continuation'] + c['rays_shadow']) / c['samples']:.3f}, {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
