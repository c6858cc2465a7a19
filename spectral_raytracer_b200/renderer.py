"""Thin object wrapper over the C ABI (include/srt.h) for tests, bench.py and the
multi-GPU driver.  Host buffers in, host buffers out; all rendering happens in
libsrt.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _native as N


@dataclass
class FlatScene:
    """The flattened RaytracingUniforms (shader.rs:32-41) the C ABI takes.

    objects   (n, 26) f32: min3 max3 kind center3 dims3 rot9(row-major) material  [last 3 columns unused]
    materials (n, 5)  f32: metallicness roughness reflectance-spectrum-index transmissive ior_a ior_b -> see from_tables
    """
    n_lambda: int
    camera: np.ndarray                      # (10,) position3 direction3 up3 fov_y_deg
    objects: np.ndarray                     # (n_obj, 23) min3 max3 kind center3 dims3 rot9 material
    materials: np.ndarray                   # (n_mat, 6) metallicness roughness spectrum transmissive ior_a ior_b
    lights: np.ndarray                      # (n_light, 4) position3 spectrum
    spectra: np.ndarray                     # (n_spec, n_lambda)
    lambda_min: float = 380.0               # spectrum.rs:5
    lambda_max: float = 780.0               # spectrum.rs:6
    meta: dict = field(default_factory=dict)

    @staticmethod
    def from_tables(n_lambda, camera, objects26, material_table, light_table, **meta) -> "FlatScene":
        """Build from per-object / per-material / per-light tables that carry their spectra inline:
        objects26 (n,26) as above with the material index in column 22; material_table (n, 2+n_lambda) =
        metallicness, roughness, reflectance (already min1-clamped, spectrum.rs:486-494); light_table
        (n, 3+n_lambda) = position, emission."""
        objects26 = np.asarray(objects26, np.float32).reshape(-1, 26)
        material_table = np.asarray(material_table, np.float32).reshape(-1, 2 + n_lambda)
        light_table = np.asarray(light_table, np.float32).reshape(-1, 3 + n_lambda)
        n_mat, n_light = material_table.shape[0], light_table.shape[0]
        spectra = np.concatenate([material_table[:, 2:], light_table[:, 3:]], axis=0).astype(np.float32)
        materials = np.zeros((n_mat, 6), np.float32)
        materials[:, 0:2] = material_table[:, 0:2]
        materials[:, 2] = np.arange(n_mat)
        materials[:, 4] = 1.0
        lights = np.zeros((n_light, 4), np.float32)
        lights[:, 0:3] = light_table[:, 0:3]
        lights[:, 3] = n_mat + np.arange(n_light)
        return FlatScene(n_lambda, np.asarray(camera, np.float32).copy(), objects26[:, :23].copy(), materials, lights,
                         spectra, meta=dict(meta))


class Renderer:
    """One srt_ctx.  Mirrors the life cycle of App::dispatch_render + App::render
    (main.rs:1376-1427, :1327-1371): create (validate, upload) -> render_frames -> resolve."""

    def __init__(self, scene: FlatScene, width: int, height: int, *, max_bounces: int = 30,
                 intended_frames: int = 100, rng: int = N.RNG_PCG3D_REFERENCE, math: int = N.MATH_FAST,
                 accel: int = N.ACCEL_AUTO, integrator: int = N.INTEGRATOR_AUTO, device: int = -1,
                 pool_paths: int = 0, philox_seed=(0, 0)):
        L = N.lib()
        self._L = L
        self.scene = scene
        self.width, self.height, self.n_lambda = width, height, scene.n_lambda
        p = N.SrtParams(width, height, scene.n_lambda, scene.lambda_min, scene.lambda_max, max_bounces,
                        intended_frames, rng, math, accel, integrator, device, pool_paths,
                        philox_seed[0], philox_seed[1])
        cam = N.SrtCamera()
        c = np.asarray(scene.camera, np.float32)
        cam.position[:] = c[0:3].tolist()
        cam.direction[:] = c[3:6].tolist()
        cam.up[:] = c[6:9].tolist()
        cam.fov_y_deg = float(c[9])
        n_obj, n_mat, n_light = len(scene.objects), len(scene.materials), len(scene.lights)
        objs = (N.SrtObject * max(1, n_obj))()
        if n_obj:
            # one bulk copy: the (n, 23) f32/u32 rows have exactly the layout of srt_object
            raw = np.zeros((n_obj, 23), np.uint32)
            o = np.asarray(scene.objects, np.float32)
            raw[:, 0:6] = o[:, 0:6].view(np.uint32)
            raw[:, 6] = o[:, 6].astype(np.uint32)
            raw[:, 7:22] = o[:, 7:22].view(np.uint32)
            raw[:, 22] = o[:, 22].astype(np.uint32)
            assert C.sizeof(N.SrtObject) == 23 * 4
            C.memmove(objs, raw.ctypes.data, raw.nbytes)
        mats = (N.SrtMaterial * max(1, n_mat))()
        for i, m in enumerate(np.asarray(scene.materials, np.float32)):
            mats[i] = N.SrtMaterial(float(m[0]), float(m[1]), int(m[2]), int(m[3]), float(m[4]), float(m[5]))
        ligs = (N.SrtLight * max(1, n_light))()
        for i, l in enumerate(np.asarray(scene.lights, np.float32)):
            ligs[i].position[:] = l[0:3].tolist()
            ligs[i].spectrum = int(l[3])
        spectra = np.ascontiguousarray(scene.spectra, np.float32)
        h = C.c_void_p()
        rc = L.srt_create(C.byref(p), C.byref(cam), objs, n_obj, mats, n_mat, ligs, n_light,
                          spectra.ctypes.data_as(C.POINTER(C.c_float)), spectra.shape[0], C.byref(h))
        if rc != N.SRT_OK:
            msg = L.srt_last_error(None)
            raise N.SrtError(rc, msg.decode() if msg else "")
        self._h = h

    @classmethod
    def open_checkpoint(cls, path: str, device: int = -1) -> "Renderer":
        """srt_checkpoint_open: a context rebuilt from a checkpoint file alone (scene + accumulated frames)."""
        L = N.lib()
        h = C.c_void_p()
        rc = L.srt_checkpoint_open(str(path).encode(), device, C.byref(h))
        if rc != N.SRT_OK:
            msg = L.srt_last_error(None)
            raise N.SrtError(rc, msg.decode() if msg else "")
        r = cls.__new__(cls)
        r._L, r._h, r.scene = L, h, None
        p = N.SrtParams()
        N.check(L.srt_get_params(h, C.byref(p)), h)
        r.width, r.height, r.n_lambda = int(p.width), int(p.height), int(p.n_lambda)
        return r

    def save_checkpoint(self, path: str):
        N.check(self._L.srt_checkpoint_save(self._h, str(path).encode()), self._h)

    def load_checkpoint(self, path: str):
        N.check(self._L.srt_checkpoint_load(self._h, str(path).encode()), self._h)

    # ---- life cycle
    def close(self):
        if getattr(self, "_h", None):
            self._L.srt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- rendering
    def render_frames(self, first_frame: int, n_frames: int):
        N.check(self._L.srt_render_frames(self._h, first_frame, n_frames), self._h)

    def render_progressive(self, first_frame: int, n_frames: int, frames_per_update: int = 1, on_update=None,
                           preview: bool = False) -> bool:
        """srt_render_progressive (App::render's FrameUpdate / RenderingProgressUpdate / AbortRender protocol,
        main.rs:1338-1357).  on_update(frames_done, frames_total, rgba8 or None) -> truthy to abort; the image is
        an (H, W, 4) uint8 view valid during the callback only.  Returns True when the render was aborted."""
        errors = []

        def _cb(_user, done, total, img):
            try:
                if on_update is None:
                    return 0
                arr = None
                if img:
                    arr = np.ctypeslib.as_array(img, shape=(self.height, self.width, 4))
                return 1 if on_update(int(done), int(total), arr) else 0
            except Exception as e:  # never unwind through the C frames
                errors.append(e)
                return 1

        cb = N.PROGRESS_FN(_cb)
        rc = self._L.srt_render_progressive(self._h, first_frame, n_frames, frames_per_update, 1 if preview else 0, cb, None)
        if errors:
            raise errors[0]
        if rc == N.SRT_ERR_ABORTED:
            return True
        N.check(rc, self._h)
        return False

    def clear(self):
        N.check(self._L.srt_clear(self._h), self._h)

    def abort(self):
        N.check(self._L.srt_abort(self._h), self._h)

    @property
    def frames_accumulated(self) -> int:
        return int(self._L.srt_frames_accumulated(self._h))

    @frames_accumulated.setter
    def frames_accumulated(self, n: int):
        N.check(self._L.srt_set_frames_accumulated(self._h, n), self._h)

    def read_accum(self) -> np.ndarray:
        out = np.empty((self.height, self.width, self.n_lambda), np.float32)
        N.check(self._L.srt_read_accum(self._h, out.ctypes.data_as(C.POINTER(C.c_float))), self._h)
        return out

    def write_accum(self, accum: np.ndarray, n_frames: int):
        a = np.ascontiguousarray(accum, np.float32)
        assert a.size == self.height * self.width * self.n_lambda
        N.check(self._L.srt_write_accum(self._h, a.ctypes.data_as(C.POINTER(C.c_float)), n_frames), self._h)

    def accum_device_ptr(self):
        n = C.c_size_t()
        p = self._L.srt_accum_device_ptr(self._h, C.byref(n))
        return int(p), int(n.value)

    def stream(self) -> int:
        return int(self._L.srt_stream(self._h) or 0)

    def resolve_rgba_f32(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.height, self.width, 4), np.float32)
        N.check(self._L.srt_resolve_rgba_f32(self._h, out.ctypes.data_as(C.POINTER(C.c_float))), self._h)
        return out

    def resolve_rgba_u8(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.height, self.width, 4), np.uint8)
        N.check(self._L.srt_resolve_rgba_u8(self._h, out.ctypes.data_as(C.POINTER(C.c_uint8))), self._h)
        return out

    def primary_ids(self, frame: int = 0, with_t: bool = True):
        ids = np.empty((self.height, self.width), np.int32)
        t = np.empty((self.height, self.width), np.float32) if with_t else None
        N.check(self._L.srt_primary_ids(self._h, frame, ids.ctypes.data_as(C.POINTER(C.c_int32)),
                                        t.ctypes.data_as(C.POINTER(C.c_float)) if with_t else None), self._h)
        return (ids, t) if with_t else ids

    def counters(self) -> dict:
        c = N.SrtCounters()
        N.check(self._L.srt_get_counters(self._h, C.byref(c)), self._h)
        return {n: int(getattr(c, n)) for n, _ in N.SrtCounters._fields_}

    def reset_counters(self):
        N.check(self._L.srt_reset_counters(self._h), self._h)

    def last_render_stats(self):
        ms = C.c_float()
        k = C.c_uint64()
        N.check(self._L.srt_last_render_stats(self._h, C.byref(ms), C.byref(k)), self._h)
        return float(ms.value), int(k.value)


    def set_deterministic(self, on: bool):
        """srt_set_deterministic: one launch per frame, bit-reproducible accumulation."""
        N.check(self._L.srt_set_deterministic(self._h, int(on)), self._h)

    def set_profiling(self, on: bool):
        N.check(self._L.srt_set_profiling(self._h, int(on)), self._h)

    def last_stage_times(self):
        """(ms[3], launches[3]) for generate / extend / shade of the last render_frames call."""
        ms = (C.c_float * 3)()
        k = (C.c_uint64 * 3)()
        N.check(self._L.srt_last_stage_times(self._h, ms, k), self._h)
        return [float(x) for x in ms], [int(x) for x in k]


_nccl_lib = None


def nccl_lib():
    """libsrt_nccl.so (srt_reduce and friends), loaded on first use."""
    global _nccl_lib
    if _nccl_lib is None:
        import os
        N.lib()
        L = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libsrt_nccl.so"))
        L.srt_reduce.argtypes = [C.POINTER(C.c_void_p), C.c_uint32]
        L.srt_reduce_last_error.restype = C.c_char_p
        L.srt_reduce_last_ms.restype = C.c_float
        L.srt_reduce_shutdown.restype = None
        _nccl_lib = L
    return _nccl_lib


def reduce_contexts(renderers) -> float:
    """srt_reduce (libsrt_nccl.so): sum the accumulation buffers of contexts living on DIFFERENT devices of this
    process into renderers[0] with NCCL; renderers[0] then holds the whole render, the others are cleared.
    Returns the device time of the reduce in ms."""
    L = nccl_lib()
    arr = (C.c_void_p * len(renderers))(*[r._h for r in renderers])
    rc = L.srt_reduce(arr, len(renderers))
    if rc != N.SRT_OK:
        raise N.SrtError(rc, L.srt_reduce_last_error().decode())
    return float(L.srt_reduce_last_ms())


def spectrum_to_rgb(spectra: np.ndarray, lambda_min: float = 380.0, lambda_max: float = 780.0) -> np.ndarray:
    """Spectrum::get_rgb_early (spectrum.rs:238-261) for a batch of spectra, on the GPU."""
    s = np.ascontiguousarray(spectra, np.float32)
    if s.ndim == 1:
        s = s[None, :]
    out = np.empty((s.shape[0], 3), np.float32)
    rc = N.lib().srt_spectrum_to_rgb(s.ctypes.data_as(C.POINTER(C.c_float)), s.shape[0], s.shape[1], lambda_min,
                                     lambda_max, out.ctypes.data_as(C.POINTER(C.c_float)))
    N.check(rc, None)
    return out


def selftest_arith(n: int = 1 << 26, seed: int = 1) -> int:
    """srt_selftest_arith: number of results of the kernels' batched exact reciprocal / quotient helpers that
    differ from the IEEE operations on n pseudo-random operand sets (0 = bit-identical)."""
    bad = C.c_uint64(0)
    N.check(N.lib().srt_selftest_arith(n, seed, C.byref(bad)), None)
    return int(bad.value)


def _rows(spectra):
    s = np.ascontiguousarray(spectra, np.float32)
    return s[None, :] if s.ndim == 1 else s


def spectra_resample(spectra: np.ndarray, n_new: int) -> np.ndarray:
    """Spectrum::resample (spectrum.rs:285-323) for a batch of spectra, on the GPU."""
    s = _rows(spectra)
    out = np.empty((s.shape[0], n_new), np.float32)
    N.check(N.lib().srt_spectra_resample(s.ctypes.data_as(C.POINTER(C.c_float)), s.shape[0], s.shape[1], n_new,
                                         out.ctypes.data_as(C.POINTER(C.c_float))), None)
    return out


def spectra_radiance(spectra: np.ndarray, lambda_min: float = 380.0, lambda_max: float = 780.0) -> np.ndarray:
    """Spectrum::get_radiance (spectrum.rs:357-362) for a batch of spectra, on the GPU."""
    s = _rows(spectra)
    out = np.empty(s.shape[0], np.float32)
    N.check(N.lib().srt_spectra_radiance(s.ctypes.data_as(C.POINTER(C.c_float)), s.shape[0], s.shape[1], lambda_min, lambda_max,
                                         out.ctypes.data_as(C.POINTER(C.c_float))), None)
    return out


def spectra_normalize(spectra: np.ndarray, lambda_min: float = 380.0, lambda_max: float = 780.0) -> np.ndarray:
    """Spectrum::normalize (spectrum.rs:369-374) for a batch of spectra, on the GPU."""
    s = _rows(spectra)
    out = np.empty_like(s)
    N.check(N.lib().srt_spectra_normalize(s.ctypes.data_as(C.POINTER(C.c_float)), s.shape[0], s.shape[1], lambda_min, lambda_max,
                                          out.ctypes.data_as(C.POINTER(C.c_float))), None)
    return out
