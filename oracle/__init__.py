"""ctypes binding of the CPU oracle (oracle.cpp).

TEST INFRASTRUCTURE ONLY: import this from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never from the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

COUNTER_NAMES = (
    "samples", "rays_primary", "rays_continuation", "rays_shadow", "slab_tests", "shape_sphere",
    "shape_plain", "shape_rotated", "hits", "self_hits", "misses", "lit", "spec_hits", "spec_dropped",
)


_TIGHT_LIB_PATH = os.path.join(_HERE, "liboracle_tight.so")


def build(force: bool = False) -> str:
    """Compile oracle.cpp -> liboracle.so (faithful cost structure) and liboracle_tight.so (-DORACLE_TIGHT: same
    arithmetic, the reference's avoidable cost items removed) with the committed Makefile."""
    src = os.path.join(_HERE, "oracle.cpp")
    stale = [p for p in (_LIB_PATH, _TIGHT_LIB_PATH) if not os.path.exists(p) or os.path.getmtime(p) < os.path.getmtime(src)]
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "all"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_libs = {}


def lib(tight: bool = False):
    if tight not in _libs:
        path = _TIGHT_LIB_PATH if tight else _LIB_PATH
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        fp = C.POINTER(C.c_float)
        L.orc_scene_new.restype = C.c_void_p
        L.orc_scene_new.argtypes = [C.c_uint32]
        L.orc_scene_free.argtypes = [C.c_void_p]
        L.orc_scene_preset.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32]
        L.orc_scene_set_camera.argtypes = [C.c_void_p, fp, fp, fp, C.c_float]
        L.orc_scene_add_spectrum.argtypes = [C.c_void_p, fp]
        L.orc_scene_add_spectrum.restype = C.c_uint32
        L.orc_scene_add_material.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_uint32]
        L.orc_scene_add_material.restype = C.c_uint32
        L.orc_scene_add_light.argtypes = [C.c_void_p, fp, C.c_uint32]
        L.orc_scene_add_sphere.argtypes = [C.c_void_p, fp, C.c_float, C.c_uint32]
        L.orc_scene_add_box.argtypes = [C.c_void_p, fp, fp, C.c_uint32]
        L.orc_scene_add_rotated_box.argtypes = [C.c_void_p, fp, fp, fp, C.c_uint32]
        u32p = C.POINTER(C.c_uint32)
        L.orc_scene_counts.argtypes = [C.c_void_p, u32p, u32p, u32p, u32p]
        L.orc_scene_counts.restype = C.c_uint32
        L.orc_scene_add_glass.argtypes = [C.c_void_p, C.c_uint32, C.c_float, C.c_float]
        L.orc_scene_add_glass.restype = C.c_uint32
        L.orc_scene_export_materials_ext.argtypes = [C.c_void_p, fp]
        for n in ("objects", "materials", "lights", "camera"):
            getattr(L, "orc_scene_export_" + n).argtypes = [C.c_void_p, fp]
        L.orc_hammersley.argtypes = [C.c_uint32, C.c_uint32, fp]
        L.orc_pcg3d.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, u32p, fp]
        L.orc_wavelength_to_xyz.argtypes = [C.c_float, fp]
        L.orc_black_body.argtypes = [C.c_double, C.c_double]
        L.orc_black_body.restype = C.c_double
        L.orc_xyz_to_rgb.argtypes = [fp, fp]
        L.orc_get_rgb_early.argtypes = [fp, C.c_uint32, C.c_float, C.c_float, fp]
        L.orc_rgb_loop_count.argtypes = [C.c_uint32, C.c_float, C.c_float]
        L.orc_rgb_loop_count.restype = C.c_uint32
        L.orc_spectrum_build.argtypes = [C.c_uint32, C.c_uint32, C.c_float, C.c_float, fp]
        L.orc_spectrum_resample.argtypes = [fp, C.c_uint32, C.c_uint32, fp]
        L.orc_spectrum_radiance.argtypes = [fp, C.c_uint32, C.c_float, C.c_float]
        L.orc_spectrum_radiance.restype = C.c_float
        L.orc_spectrum_normalize.argtypes = [fp, C.c_uint32, C.c_float, C.c_float, fp]
        L.orc_euler_rotation.argtypes = [C.c_float, C.c_float, C.c_float, fp]
        L.orc_cosine_direction.argtypes = [C.c_float, C.c_float, fp, fp]
        L.orc_cone_direction.argtypes = [fp, C.c_float, C.c_float, C.c_float, fp]
        L.orc_to_rgba8.argtypes = [fp, C.c_size_t, C.POINTER(C.c_uint8)]
        L.orc_render.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                 C.c_uint32, fp, C.POINTER(C.c_double)]
        L.orc_sample.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                 C.c_uint32, fp, fp, u32p]
        L.orc_primary.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_int32), fp,
                                  C.POINTER(C.c_uint8)]
        L.orc_counters_get.argtypes = [C.POINTER(C.c_uint64)]
        L.orc_render_pixels.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32, fp,
                                        C.c_int]
        L.orc_set_modes.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_uint32]
        L.orc_hardware_threads.restype = C.c_uint
        assert bool(L.orc_is_tight()) == tight
        _libs[tight] = L
    return _libs[tight]


def _fp(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _f3(v):
    return _fp(np.ascontiguousarray(v, dtype=np.float32))


class Scene:
    """Opaque oracle scene (RaytracingUniforms + the spectra/material tables)."""

    def __init__(self, n_lambda: int = 32, preset: str | None = None, arg: int = 0, tight: bool = False):
        self.n_lambda = n_lambda
        self._L = lib(tight)
        self._h = self._L.orc_scene_new(n_lambda)
        if not self._h:
            raise ValueError("illegal number of spectral samples (multiple of 8, <= 128)")
        if preset is not None:
            if self._L.orc_scene_preset(self._h, preset.encode(), arg) != 0:
                raise ValueError(f"unknown preset {preset}")

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_scene_free(self._h)
            self._h = None

    # ---- building
    def set_camera(self, pos, direction, up, fov_y_deg):
        self._L.orc_scene_set_camera(self._h, _f3(pos), _f3(direction), _f3(up), fov_y_deg)

    def add_spectrum(self, values) -> int:
        v = np.ascontiguousarray(values, dtype=np.float32)
        assert v.shape == (self.n_lambda,)
        return self._L.orc_scene_add_spectrum(self._h, _fp(v))

    def add_material(self, metallicness, roughness, spectrum_id) -> int:
        return self._L.orc_scene_add_material(self._h, metallicness, roughness, spectrum_id)

    def add_glass(self, spectrum_id, ior_a, ior_b) -> int:
        """extension: dispersive dielectric, n(lambda) = ior_a + ior_b / lambda_nm^2"""
        return self._L.orc_scene_add_glass(self._h, spectrum_id, ior_a, ior_b)

    def add_light(self, pos, spectrum_id):
        self._L.orc_scene_add_light(self._h, _f3(pos), spectrum_id)

    def add_sphere(self, center, radius, material):
        self._L.orc_scene_add_sphere(self._h, _f3(center), radius, material)

    def add_box(self, center, lengths, material):
        self._L.orc_scene_add_box(self._h, _f3(center), _f3(lengths), material)

    def add_rotated_box(self, center, lengths, euler, material):
        self._L.orc_scene_add_rotated_box(self._h, _f3(center), _f3(lengths), _f3(euler), material)

    # ---- export
    def export(self) -> dict:
        n = [C.c_uint32() for _ in range(4)]
        nl = self._L.orc_scene_counts(self._h, *[C.byref(x) for x in n])
        n_obj, n_mat, n_light, n_spec = [x.value for x in n]
        obj = np.zeros((n_obj, 26), np.float32)
        mat = np.zeros((n_mat, 2 + nl), np.float32)
        lig = np.zeros((n_light, 3 + nl), np.float32)
        cam = np.zeros(10, np.float32)
        if n_obj:
            self._L.orc_scene_export_objects(self._h, _fp(obj))
        if n_mat:
            self._L.orc_scene_export_materials(self._h, _fp(mat))
        if n_light:
            self._L.orc_scene_export_lights(self._h, _fp(lig))
        self._L.orc_scene_export_camera(self._h, _fp(cam))
        ext = np.zeros((n_mat, 3), np.float32)
        if n_mat:
            self._L.orc_scene_export_materials_ext(self._h, _fp(ext))
        return {"n_lambda": nl, "objects": obj, "materials": mat, "lights": lig, "camera": cam, "materials_ext": ext}

    # ---- rendering
    def render(self, w, h, n_frames, *, first_frame=0, intended_frames=None, max_bounces=30, threads=0,
               img=None, spectral=False):
        """App::render for frames [first_frame, first_frame+n_frames).  Returns the
        RGBA f32 running-mean image (h, w, 4) and, if spectral, the per-pixel f64 sum of spectra."""
        if intended_frames is None:
            intended_frames = first_frame + n_frames
        if img is None:
            img = np.zeros((h, w, 4), np.float32)
        spec = np.zeros((h, w, self.n_lambda), np.float64) if spectral else None
        rc = self._L.orc_render(self._h, w, h, max_bounces, first_frame, n_frames, intended_frames, threads, _fp(img),
                              spec.ctypes.data_as(C.POINTER(C.c_double)) if spectral else None)
        assert rc == 0
        return (img, spec) if spectral else img

    def sample(self, w, h, x, y, frame, intended_frames, max_bounces=30):
        spec = np.zeros(self.n_lambda, np.float32)
        rgb = np.zeros(3, np.float32)
        depth = C.c_uint32()
        self._L.orc_sample(self._h, w, h, max_bounces, x, y, frame, intended_frames, _fp(spec), _fp(rgb), C.byref(depth))
        return spec, rgb, depth.value

    def render_pixels(self, w, h, xy, n_frames, max_bounces=30, threads=0):
        """The frame loop for a subset of the pixels of a w x h image: (n, 4) f32 running means."""
        xy = np.ascontiguousarray(xy, np.uint32)
        out = np.zeros((xy.shape[0], 4), np.float32)
        self._L.orc_render_pixels(self._h, w, h, max_bounces, n_frames, xy.ctypes.data_as(C.POINTER(C.c_uint32)), xy.shape[0],
                                _fp(out), threads)
        return out

    def primary(self, w, h, frame=0, intended_frames=1):
        ids = np.zeros((h, w), np.int32)
        t = np.zeros((h, w), np.float32)
        band = np.zeros((h, w), np.uint8)
        self._L.orc_primary(self._h, w, h, frame, intended_frames, ids.ctypes.data_as(C.POINTER(C.c_int32)), _fp(t),
                          band.ctypes.data_as(C.POINTER(C.c_uint8)))
        return ids, t, band


MATH_NATIVE, MATH_CANONICAL = 0, 1
RNG_PCG3D, RNG_PHILOX = 0, 1


def set_modes(math_mode=MATH_NATIVE, rng_mode=RNG_PCG3D, philox_key=(0, 0)):
    """Process-wide oracle modes (see oracle.cpp, g_math_mode): math 0 = platform f32 libm
    (what the reference calls), 1 = correctly rounded; rng 0 = pcg3d (reference), 1 = Philox."""
    lib().orc_set_modes(math_mode, rng_mode, philox_key[0], philox_key[1])
    if True in _libs:  # the tight build has its own copy of the process-wide modes
        _libs[True].orc_set_modes(math_mode, rng_mode, philox_key[0], philox_key[1])


# ---- known-answer helpers
def hammersley(n, N):
    o = np.zeros(2, np.float32)
    lib().orc_hammersley(n, N, _fp(o))
    return float(o[0]), float(o[1])


def pcg3d(x, y, z):
    raw = (C.c_uint32 * 3)()
    f = np.zeros(3, np.float32)
    lib().orc_pcg3d(x, y, z, raw, _fp(f))
    return tuple(int(v) for v in raw), f


def wavelength_to_xyz(w):
    o = np.zeros(3, np.float32)
    lib().orc_wavelength_to_xyz(w, _fp(o))
    return o


def black_body(wavelength_nm, temperature_k):
    return lib().orc_black_body(wavelength_nm, temperature_k)


def xyz_to_rgb(xyz):
    o = np.zeros(3, np.float32)
    lib().orc_xyz_to_rgb(_f3(xyz), _fp(o))
    return o


def get_rgb_early(intensities, lo=380.0, hi=780.0):
    v = np.ascontiguousarray(intensities, dtype=np.float32)
    o = np.zeros(3, np.float32)
    lib().orc_get_rgb_early(_fp(v), v.shape[0], lo, hi, _fp(o))
    return o


def spectrum_resample(intensities, n_new):
    """Spectrum::resample (spectrum.rs:285-323); None where the reference panics."""
    v = np.ascontiguousarray(intensities, dtype=np.float32)
    o = np.zeros(n_new, np.float32)
    return None if lib().orc_spectrum_resample(_fp(v), v.shape[0], n_new, _fp(o)) else o


def spectrum_radiance(intensities, lo=380.0, hi=780.0):
    v = np.ascontiguousarray(intensities, dtype=np.float32)
    return np.float32(lib().orc_spectrum_radiance(_fp(v), v.shape[0], lo, hi))


def spectrum_normalize(intensities, lo=380.0, hi=780.0):
    v = np.ascontiguousarray(intensities, dtype=np.float32)
    o = np.zeros_like(v)
    lib().orc_spectrum_normalize(_fp(v), v.shape[0], lo, hi, _fp(o))
    return o


def rgb_loop_count(n, lo=380.0, hi=780.0):
    return lib().orc_rgb_loop_count(n, lo, hi)


SPEC_TEMPERATURE, SPEC_FLAT, SPEC_RED, SPEC_GREEN, SPEC_BLUE, SPEC_SUN = range(6)


def spectrum(kind, n, arg0=1.0, arg1=1.0):
    o = np.zeros(n, np.float32)
    assert lib().orc_spectrum_build(kind, n, arg0, arg1, _fp(o)) == 0
    return o


def euler_rotation(roll, pitch, yaw):
    o = np.zeros(9, np.float32)
    lib().orc_euler_rotation(roll, pitch, yaw, _fp(o))
    return o.reshape(3, 3)


def cosine_direction(rx, ry, normal):
    o = np.zeros(3, np.float32)
    lib().orc_cosine_direction(rx, ry, _f3(normal), _fp(o))
    return o


def cone_direction(direction, roughness, rx, ry):
    o = np.zeros(3, np.float32)
    lib().orc_cone_direction(_f3(direction), roughness, rx, ry, _fp(o))
    return o


def to_rgba8(data):
    d = np.ascontiguousarray(data, dtype=np.float32)
    o = np.zeros(d.size, np.uint8)
    lib().orc_to_rgba8(_fp(d.reshape(-1)), d.size, o.ctypes.data_as(C.POINTER(C.c_uint8)))
    return o.reshape(d.shape)


def counters_reset():
    lib().orc_counters_reset()


def counters():
    buf = (C.c_uint64 * (14 + 129))()
    lib().orc_counters_get(buf)
    d = {k: int(buf[i]) for i, k in enumerate(COUNTER_NAMES)}
    d["depth_hist"] = [int(buf[14 + i]) for i in range(129)]
    return d


def hardware_threads() -> int:
    return int(lib().orc_hardware_threads())
