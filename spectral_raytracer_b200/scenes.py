"""Scene presets, built by the C++ host mirror (host/srt_host.hpp -> libsrt_host.so): the
README default scene (main.rs:1638-1758), the Cornell box (main.rs:1538-1635), the 10k random-sphere
scene and the prism extension of BASELINE.json's configs.  Python only moves the arrays."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _native as N
from .renderer import FlatScene

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "libsrt_host.so")

HOST_EXPORTS = ("srth_last_error", "srth_scene_preset", "srth_scene_free", "srth_scene_counts", "srth_scene_copy",
                "srth_spectrum", "srth_black_body", "srth_to_rgba8", "srth_save_png", "srth_dispatch_render", "srth_render_protocol", "srth_spectrum_tool")

_lib = None


def host_lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    N.lib()  # libsrt.so first (libsrt_host.so links against it)
    if not os.path.exists(HOST_LIB_PATH):
        raise ImportError(f"{HOST_LIB_PATH} is missing: run __graft_entry__.build()")
    L = C.CDLL(HOST_LIB_PATH)
    u32 = C.c_uint32
    fp = C.POINTER(C.c_float)
    L.srth_last_error.restype = C.c_char_p
    L.srth_scene_preset.argtypes = [C.c_char_p, u32, u32]
    L.srth_scene_preset.restype = C.c_void_p
    L.srth_scene_free.argtypes = [C.c_void_p]
    L.srth_scene_counts.argtypes = [C.c_void_p] + [C.POINTER(u32)] * 5
    L.srth_scene_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, fp, C.c_void_p, fp]
    L.srth_spectrum.argtypes = [u32, u32, C.c_float, C.c_float, fp]
    L.srth_black_body.argtypes = [C.c_double, C.c_double]
    L.srth_black_body.restype = C.c_double
    L.srth_to_rgba8.argtypes = [fp, C.c_size_t, C.POINTER(C.c_uint8)]
    L.srth_save_png.argtypes = [C.c_char_p, u32, u32, fp]
    L.srth_dispatch_render.argtypes = [C.c_char_p, u32, u32, u32, u32, u32, u32, u32, u32, C.c_int32, u32, u32, fp,
                                       C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(N.SrtCounters)]
    L.srth_spectrum_tool.argtypes = [u32, fp, u32, u32, fp]
    L.srth_render_protocol.argtypes = [C.c_char_p, u32, u32, u32, u32, u32, u32, u32, C.c_int32, C.POINTER(C.c_int32), fp, u32,
                                       C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.POINTER(C.c_int32)]
    _lib = L
    return L


def preset(name: str, n_lambda: int = 32, arg: int = 0) -> FlatScene:
    """'default' | 'cornell' | 'spheres' (arg = sphere count) | 'prism' -> flattened RaytracingUniforms."""
    L = host_lib()
    h = L.srth_scene_preset(name.encode(), n_lambda, arg)
    if not h:
        raise ValueError(L.srth_last_error().decode())
    try:
        n = [C.c_uint32() for _ in range(5)]
        L.srth_scene_counts(h, *[C.byref(x) for x in n])
        nl, n_obj, n_mat, n_light, n_spec = [x.value for x in n]
        objs = (N.SrtObject * max(1, n_obj))()
        mats = (N.SrtMaterial * max(1, n_mat))()
        ligs = (N.SrtLight * max(1, n_light))()
        spectra = np.zeros((n_spec, nl), np.float32)
        cam = N.SrtCamera()
        mm = np.zeros(2, np.float32)
        L.srth_scene_copy(h, objs, mats, ligs, spectra.ctypes.data_as(C.POINTER(C.c_float)), C.byref(cam),
                          mm.ctypes.data_as(C.POINTER(C.c_float)))
    finally:
        L.srth_scene_free(h)
    raw = np.frombuffer(objs, dtype=np.uint32, count=n_obj * 23).reshape(n_obj, 23).copy()
    objects = raw.view(np.float32).copy()
    objects[:, 6] = raw[:, 6].astype(np.float32)
    objects[:, 22] = raw[:, 22].astype(np.float32)
    materials = np.array([[m.metallicness, m.roughness, m.reflectance, m.transmissive, m.ior_a, m.ior_b]
                          for m in mats[:n_mat]], np.float32).reshape(n_mat, 6)
    lights = np.array([[*l.position, l.spectrum] for l in ligs[:n_light]], np.float32).reshape(n_light, 4)
    camera = np.array([*cam.position, *cam.direction, *cam.up, cam.fov_y_deg], np.float32)
    return FlatScene(nl, camera, objects, materials, lights, spectra, float(mm[0]), float(mm[1]),
                     meta={"preset": name, "arg": arg})


def dispatch_render(preset_name: str, width: int, height: int, iterations: int, *, arg: int = 0, n_lambda: int = 32,
                    bounces: int = 30, rng: int = N.RNG_PCG3D_REFERENCE, math: int = N.MATH_FAST, device: int = -1,
                    first_frame: int = 0, n_frames: int = 0):
    """The headless sibling of App::dispatch_render (host/srt_host.hpp: dispatch_render_headless): one call
    from scene description to the CustomImage data on the host.  Returns (image, info)."""
    L = host_lib()
    img = np.empty((height, width, 4), np.float32)
    secs = C.c_double()
    launches = C.c_uint64()
    ctr = N.SrtCounters()
    rc = L.srth_dispatch_render(preset_name.encode(), arg, width, height, n_lambda, iterations, bounces, rng, math,
                                device, first_frame, n_frames, img.ctypes.data_as(C.POINTER(C.c_float)),
                                C.byref(secs), C.byref(launches), C.byref(ctr))
    if rc != N.SRT_OK:
        raise N.SrtError(rc, L.srth_last_error().decode())
    info = {"device_seconds": secs.value, "kernel_launches": int(launches.value)}
    info.update({n: int(getattr(ctr, n)) for n, _ in N.SrtCounters._fields_})
    return img, info


ACTION_FRAME_UPDATE, ACTION_PROGRESS_UPDATE, ACTION_TRUE_TIME_UPDATE, ACTION_DESTROY_SENDER = range(4)


def render_protocol(name: str, width: int, height: int, iterations: int, frames_per_update: int = 1, abort_at_update: int = -1,
                    n_lambda: int = 32, bounces: int = 30, arg: int = 0):
    """srt_host::render (the App::render sibling, main.rs:1327-1371) on a preset: returns the action list as
    (kinds, values), the image of the last FrameUpdate (H, W, 4 uint8), frames accumulated, completed flag."""
    L = host_lib()
    cap = 2 * (iterations // max(1, frames_per_update) + 2) + 4
    kinds = (C.c_int32 * cap)()
    vals = (C.c_float * cap)()
    img = np.zeros((height, width, 4), np.uint8)
    acc = C.c_uint64(0)
    done = C.c_int32(0)
    n = L.srth_render_protocol(name.encode(), arg, width, height, n_lambda, iterations, bounces, frames_per_update,
                               abort_at_update, kinds, vals, cap, img.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(acc),
                               C.byref(done))
    if n < 0:
        raise N.SrtError(-n, L.srth_last_error().decode())
    n = min(n, cap)
    return list(kinds[:n]), list(vals[:n]), img, int(acc.value), bool(done.value)


def host_spectrum_tool(op: str, intensities, n_new: int = 0):
    """Spectrum::resample / get_radiance / normalize as methods of the C++ host mirror (device-backed)."""
    L = host_lib()
    v = np.ascontiguousarray(intensities, np.float32)
    code = {"resample": 0, "radiance": 1, "normalize": 2}[op]
    out = np.zeros(max(n_new, v.shape[0], 1), np.float32)
    fp = C.POINTER(C.c_float)
    if L.srth_spectrum_tool(code, v.ctypes.data_as(fp), v.shape[0], n_new, out.ctypes.data_as(fp)) != 0:
        raise ValueError(L.srth_last_error().decode())
    return out[:n_new] if code == 0 else (out[0] if code == 1 else out[:v.shape[0]])


def save_png(path: str, image: np.ndarray) -> None:
    """DynamicImage::from(CustomImage).save(path) (main.rs:2325-2326) by the C++ host mirror: image = (H, W, 4) f32."""
    img = np.ascontiguousarray(image, np.float32)
    h, w = img.shape[:2]
    if host_lib().srth_save_png(str(path).encode(), w, h, img.ctypes.data_as(C.POINTER(C.c_float))) != 0:
        raise OSError(f"cannot write {path}")
