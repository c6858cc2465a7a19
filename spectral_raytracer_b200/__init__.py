"""B200-native spectral render backend: a drop-in for the per-pixel render path of
happy737/spectral-raytracer (shader.rs, the per-sample part of spectrum.rs,
custom_image.rs and App::apply_shader2 / App::render in main.rs).

The product is libsrt.so (CUDA kernels + the C ABI of include/srt.h) plus the C++
host mirror of the reference's scene types under host/.  This Python package is
the ctypes plumbing tests and bench.py use; it contains no rendering code.
"""
from . import _native as native
from ._native import (ACCEL_AUTO, ACCEL_BVH, ACCEL_LINEAR, INTEGRATOR_AUTO, INTEGRATOR_RESIDENT, INTEGRATOR_WAVEFRONT, MATH_EXACT,
                      MATH_FAST, RNG_PCG3D_REFERENCE, RNG_PHILOX, SrtError)
from .renderer import (FlatScene, Renderer, reduce_contexts, selftest_arith, spectra_normalize, spectra_radiance,
                       spectra_resample, spectrum_to_rgb)

__all__ = ["native", "FlatScene", "Renderer", "spectrum_to_rgb", "selftest_arith", "spectra_resample", "spectra_radiance", "spectra_normalize", "reduce_contexts", "SrtError", "ACCEL_AUTO", "ACCEL_BVH",
           "ACCEL_LINEAR", "INTEGRATOR_AUTO", "INTEGRATOR_RESIDENT", "INTEGRATOR_WAVEFRONT", "MATH_EXACT", "MATH_FAST",
           "RNG_PCG3D_REFERENCE", "RNG_PHILOX"]
