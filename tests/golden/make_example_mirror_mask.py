"""GPU box: blocks where the default scene with mirror roughness 0 and 0.2 differ (the mirror's silhouette)."""
import sys, os, dataclasses
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import spectral_raytracer_b200 as srt
from spectral_raytracer_b200 import scenes
flat = scenes.preset("default", 32)
m = flat.materials.copy(); m[m[:, 0] > 0, 1] = 0.0
imgs = []
for f in (flat, dataclasses.replace(flat, materials=m)):
    with srt.Renderer(f, 1920, 1080, intended_frames=1000) as r:
        r.render_frames(0, 1000)
        imgs.append(r.resolve_rgba_u8()[..., :3].astype(np.float64))
bm = lambda a: a.reshape(135, 8, 240, 8, 3).mean(axis=(1, 3))
d = np.abs(bm(imgs[0]) - bm(imgs[1])).max(axis=2)
mask = d > 0.75
# dilate by one block
pad = np.pad(mask, 1)
mask = pad[1:-1, 1:-1] | pad[:-2, 1:-1] | pad[2:, 1:-1] | pad[1:-1, :-2] | pad[1:-1, 2:]
np.save(os.path.join(os.path.dirname(os.path.abspath(__file__)), "example_mirror_mask.npy"), mask)
print("mirror blocks", int(mask.sum()), "of", mask.size)
