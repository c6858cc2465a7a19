"""ctypes binding of libsrt.so (the C ABI declared in include/srt.h).

This is plumbing only: every call goes straight to the CUDA library.  There is no
Python or CPU implementation of the render path behind it -- if libsrt.so has not
been built (`python -c "import __graft_entry__ as g; g.build()"`) importing the
symbols fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SRT_LIB_PATH", os.path.join(_HERE, "libsrt.so"))  # override: kernel A/B experiments

# every symbol include/srt.h declares that lives in libsrt.so
EXPORTS = (
    "srt_abi_version", "srt_launch_param_bytes", "srt_selftest_arith", "srt_spectra_resample", "srt_spectra_radiance", "srt_spectra_normalize", "srt_device_count", "srt_create", "srt_destroy", "srt_last_error",
    "srt_render_frames", "srt_render_progressive", "srt_checkpoint_save", "srt_checkpoint_load", "srt_checkpoint_open", "srt_get_params", "srt_abort", "srt_clear", "srt_frames_accumulated", "srt_set_frames_accumulated",
    "srt_accum_device_ptr", "srt_stream", "srt_device", "srt_read_accum", "srt_write_accum",
    "srt_resolve_rgba_f32", "srt_resolve_rgba_u8", "srt_resolve_rgba_f32_device", "srt_primary_ids",
    "srt_spectrum_to_rgb", "srt_get_counters", "srt_reset_counters", "srt_last_render_stats",
    "srt_set_profiling", "srt_last_stage_times", "srt_set_deterministic",
)

SRT_OK = 0
(SRT_ERR_INVALID_ARGUMENT, SRT_ERR_SPECTRUM_SAMPLES, SRT_ERR_CAMERA_COLLINEAR, SRT_ERR_CUDA, SRT_ERR_UNSUPPORTED,
 SRT_ERR_ABORTED) = range(1, 7)
PLAIN_BOX, SPHERE, ROTATED_BOX = 0, 1, 2
RNG_PCG3D_REFERENCE, RNG_PHILOX = 0, 1
MATH_FAST, MATH_EXACT = 0, 1
ACCEL_AUTO, ACCEL_LINEAR, ACCEL_BVH = 0, 1, 2
INTEGRATOR_WAVEFRONT, INTEGRATOR_RESIDENT, INTEGRATOR_AUTO = 0, 1, 2


class SrtObject(C.Structure):
    _fields_ = [("min", C.c_float * 3), ("max", C.c_float * 3), ("kind", C.c_uint32), ("center", C.c_float * 3),
                ("dims", C.c_float * 3), ("rot", C.c_float * 9), ("material", C.c_uint32)]


class SrtMaterial(C.Structure):
    _fields_ = [("metallicness", C.c_float), ("roughness", C.c_float), ("reflectance", C.c_uint32),
                ("transmissive", C.c_uint32), ("ior_a", C.c_float), ("ior_b", C.c_float)]


class SrtLight(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("spectrum", C.c_uint32)]


class SrtCamera(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("direction", C.c_float * 3), ("up", C.c_float * 3),
                ("fov_y_deg", C.c_float)]


class SrtParams(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("n_lambda", C.c_uint32),
                ("lambda_min", C.c_float), ("lambda_max", C.c_float), ("max_bounces", C.c_uint32),
                ("intended_frames", C.c_uint32), ("rng_mode", C.c_uint32), ("math_mode", C.c_uint32),
                ("accel", C.c_uint32), ("integrator", C.c_uint32), ("device", C.c_int32),
                ("pool_paths", C.c_uint32), ("philox_seed_lo", C.c_uint32), ("philox_seed_hi", C.c_uint32)]


class SrtCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "samples", "rays_primary", "rays_continuation", "rays_shadow", "hits", "self_hits", "misses", "lit",
        "spec_hits", "spec_dropped", "iterations", "kernel_launches", "shadow_skipped")]


class SrtError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"srt error {code}: {message}")
        self.code = code
        self.message = message


_lib = None


# srt_progress_fn
PROGRESS_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint8))


def lib() -> C.CDLL:
    """Load libsrt.so and declare the prototypes.  Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA backend has not been built and there is no CPU fallback. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` at the repo root.")
    L = C.CDLL(LIB_PATH)
    vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
    fp = C.POINTER(C.c_float)
    L.srt_abi_version.restype = u32
    L.srt_launch_param_bytes.restype = u32
    L.srt_render_progressive.argtypes = [vp, u32, u32, u32, C.c_int, PROGRESS_FN, vp]
    L.srt_selftest_arith.argtypes = [u64, u32, C.POINTER(u64)]
    L.srt_get_params.argtypes = [vp, C.POINTER(SrtParams)]
    L.srt_checkpoint_save.argtypes = [vp, C.c_char_p]
    L.srt_checkpoint_load.argtypes = [vp, C.c_char_p]
    L.srt_checkpoint_open.argtypes = [C.c_char_p, C.c_int32, C.POINTER(vp)]
    L.srt_spectra_resample.argtypes = [fp, u32, u32, u32, fp]
    L.srt_spectra_radiance.argtypes = [fp, u32, u32, C.c_float, C.c_float, fp]
    L.srt_spectra_normalize.argtypes = [fp, u32, u32, C.c_float, C.c_float, fp]
    L.srt_device_count.restype = C.c_int
    L.srt_create.argtypes = [C.POINTER(SrtParams), C.POINTER(SrtCamera), C.POINTER(SrtObject), u32,
                             C.POINTER(SrtMaterial), u32, C.POINTER(SrtLight), u32, fp, u32, C.POINTER(vp)]
    L.srt_destroy.argtypes = [vp]
    L.srt_destroy.restype = None
    L.srt_last_error.argtypes = [vp]
    L.srt_last_error.restype = C.c_char_p
    L.srt_render_frames.argtypes = [vp, u32, u32]
    L.srt_abort.argtypes = [vp]
    L.srt_clear.argtypes = [vp]
    L.srt_frames_accumulated.argtypes = [vp]
    L.srt_frames_accumulated.restype = u64
    L.srt_set_frames_accumulated.argtypes = [vp, u64]
    L.srt_accum_device_ptr.argtypes = [vp, C.POINTER(C.c_size_t)]
    L.srt_accum_device_ptr.restype = vp
    L.srt_device.argtypes = [vp]
    L.srt_stream.argtypes = [vp]
    L.srt_stream.restype = vp
    L.srt_read_accum.argtypes = [vp, fp]
    L.srt_write_accum.argtypes = [vp, fp, u64]
    L.srt_resolve_rgba_f32.argtypes = [vp, fp]
    L.srt_resolve_rgba_u8.argtypes = [vp, C.POINTER(C.c_uint8)]
    L.srt_resolve_rgba_f32_device.argtypes = [vp, vp]
    L.srt_primary_ids.argtypes = [vp, u32, C.POINTER(C.c_int32), fp]
    L.srt_spectrum_to_rgb.argtypes = [fp, u32, u32, C.c_float, C.c_float, fp]
    L.srt_get_counters.argtypes = [vp, C.POINTER(SrtCounters)]
    L.srt_reset_counters.argtypes = [vp]
    L.srt_last_render_stats.argtypes = [vp, fp, C.POINTER(u64)]
    L.srt_set_profiling.argtypes = [vp, C.c_int]
    L.srt_set_deterministic.argtypes = [vp, C.c_int]
    L.srt_last_stage_times.argtypes = [vp, fp, C.POINTER(u64)]
    _lib = L
    return L


def check(rc: int, ctx=None):
    if rc != SRT_OK:
        msg = lib().srt_last_error(ctx)
        raise SrtError(rc, msg.decode() if msg else "")
