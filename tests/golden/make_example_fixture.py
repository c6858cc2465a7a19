"""Builds tests/golden/reference_example_image.npz from the image the reference's README publishes
(/root/reference/example_image.png, README.md:15: "Render of the example scene over 1000 iterations", 1920x1080
RGBA8 straight out of the Rust program's From<CustomImage> for DynamicImage).  It is the only output of the real
reference that exists here (no Rust toolchain), so it is what pins the whole per-pixel path -- shader.rs,
the per-sample part of spectrum.rs, custom_image.rs, apply_shader2 / render -- and not just spectrum.rs.

Stored (small; the PNG itself is 1.9 MB and stays where it is):
  block_means  (135, 240, 3) f16   mean RGB level of every 8x8 pixel block
  xy           (4096, 2) u16       pseudo-random pixel positions (fixed seed) + a stratified grid
  rgb          (4096, 3) u8        the image's RGB at those positions
  mirror       (135, 240) bool     8x8 blocks inside the silhouette of the mirror (see below)

One known difference between the published image and the current source: the image shows a perfectly sharp
reflection in the mirror (left of the frame), while the default scene of this revision gives the mirror
roughness 0.2 (main.rs:1696); the image predates that value.  With roughness 0 the current code reproduces the
image everywhere; with 0.2 everywhere but inside the mirror's silhouette.  `mirror` marks that region (blocks
where two renders of this repo with roughness 0 and 0.2 differ visibly; generated on the GPU box by
tests/golden/make_example_mirror_mask.py and merged here).

Run here (needs /root/reference):  python tests/golden/make_example_fixture.py
"""
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/example_image.png"


def main():
    img = np.asarray(Image.open(SRC))
    assert img.shape == (1080, 1920, 4) and img.dtype == np.uint8 and (img[..., 3] == 255).all()
    rgb = img[..., :3]
    block_means = rgb.reshape(135, 8, 240, 8, 3).astype(np.float64).mean(axis=(1, 3)).astype(np.float16)
    rng = np.random.default_rng(20251018)
    n_rand = 4096 - 32 * 18
    xs = rng.integers(0, 1920, n_rand)
    ys = rng.integers(0, 1080, n_rand)
    gx, gy = np.meshgrid(np.arange(30, 1920, 60), np.arange(30, 1080, 60))
    xy = np.concatenate([np.stack([xs, ys], 1), np.stack([gx.ravel(), gy.ravel()], 1)]).astype(np.uint16)
    assert xy.shape == (4096, 2)
    out = os.path.join(HERE, "reference_example_image.npz")
    mirror = np.zeros((135, 240), bool)
    mask_file = os.path.join(HERE, "example_mirror_mask.npy")
    if os.path.exists(mask_file):
        mirror = np.load(mask_file)
    np.savez_compressed(out, block_means=block_means, xy=xy, rgb=rgb[xy[:, 1], xy[:, 0]], mirror=mirror,
                        source=np.array("happy737/spectral-raytracer example_image.png (README.md:15), 1920x1080, 1000 iterations"))
    print(out, os.path.getsize(out), "bytes; mirror blocks:", int(mirror.sum()))


if __name__ == "__main__":
    sys.exit(main())
