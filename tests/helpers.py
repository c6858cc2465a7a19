"""Shared test helpers: hand the oracle's scene, bit for bit, to the CUDA path."""
import numpy as np

import spectral_raytracer_b200 as srt


def flat_from_oracle(oscene) -> srt.FlatScene:
    e = oscene.export()
    flat = srt.FlatScene.from_tables(e["n_lambda"], e["camera"], e["objects"], e["materials"], e["lights"])
    flat.materials[:, 3:6] = e["materials_ext"]  # dispersion extension: transmissive, ior_a, ior_b
    return flat


def rel_rmse(a, b):
    """sqrt(mean((a-b)^2)) / mean(b) on linear f32 RGB (SURVEY.md 8d gate 3)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.sqrt(np.mean((a - b) ** 2)) / np.mean(b))
