"""CPU-only checks of the host logic and the C-ABI library: every symbol include/srt.h declares is
exported, the C++ host mirror's presets flatten to exactly the oracle's scenes, and the library
refuses to work (loudly) without a GPU instead of falling back."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import spectral_raytracer_b200 as srt
from spectral_raytracer_b200 import scenes
from helpers import flat_from_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "srt.h")).read()
    declared = set(re.findall(r"\b(srt_[a-z0-9_]+)\s*\(", header))
    in_nccl = {n for n in declared if n.startswith("srt_reduce")}  # live in libsrt_nccl.so (single-process multi-device NCCL)
    assert in_nccl == {"srt_reduce", "srt_reduce_shutdown", "srt_reduce_last_ms", "srt_reduce_last_error"}
    declared -= in_nccl
    lib = srt.native.lib()
    assert declared == set(srt.native.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.srt_abi_version() == 1
    host = scenes.host_lib()
    for name in scenes.HOST_EXPORTS:
        assert hasattr(host, name), name
    nccl = C.CDLL(os.path.join(ROOT, "spectral_raytracer_b200", "libsrt_nccl.so"))
    for name in in_nccl:
        assert hasattr(nccl, name), name


def test_abi_struct_sizes_match_header():
    # no padding: plain 4-byte fields
    assert C.sizeof(srt.native.SrtObject) == 23 * 4
    assert C.sizeof(srt.native.SrtMaterial) == 6 * 4
    assert C.sizeof(srt.native.SrtLight) == 4 * 4
    assert C.sizeof(srt.native.SrtCamera) == 10 * 4
    assert C.sizeof(srt.native.SrtParams) == 15 * 4
    assert C.sizeof(srt.native.SrtCounters) == 13 * 8


@pytest.mark.parametrize("name,arg,n_lambda", [("cornell", 0, 32), ("default", 0, 32), ("spheres", 500, 32),
                                               ("cornell", 0, 64), ("default", 0, 8), ("prism", 0, 32)])
def test_host_presets_equal_oracle_scenes(oracle, name, arg, n_lambda):
    """dispatch_render's uniform assembly in the C++ host mirror vs the oracle's restatement of the same
    reference code (main.rs:1389-1404, :1538-1758; shader.rs:108-166): bit-identical inputs."""
    want = flat_from_oracle(oracle.Scene(n_lambda, name, arg))
    got = scenes.preset(name, n_lambda, arg)
    assert got.n_lambda == want.n_lambda
    assert np.array_equal(got.camera, want.camera)
    assert np.array_equal(got.objects[:, :7].view(np.uint32), want.objects[:, :7].view(np.uint32))
    rotated = want.objects[:, 6] == 2  # the RotatedBox payload only exists for that kind (shader.rs:171)
    assert np.array_equal(got.objects[rotated, 7:22].view(np.uint32), want.objects[rotated, 7:22].view(np.uint32))
    # material / spectrum numbering may differ; compare what each object / light resolves to
    def obj_material(f, i):
        m = f.materials[int(f.objects[i, 22])]
        return m[0], m[1], f.spectra[int(m[2])], tuple(m[3:6]) if m[3] else (0,)
    for i in range(len(want.objects)):
        a, b = obj_material(got, i), obj_material(want, i)
        assert a[0] == b[0] and a[1] == b[1] and np.array_equal(a[2], b[2]) and a[3] == b[3]
    assert len(got.lights) == len(want.lights)
    for a, b in zip(got.lights, want.lights):
        assert np.array_equal(a[:3], b[:3])
        assert np.array_equal(got.spectra[int(a[3])], want.spectra[int(b[3])])


def test_host_spectrum_constructors_equal_oracle(oracle):
    L = scenes.host_lib()
    for n in (8, 32, 128):
        for kind, a0, a1 in [(0, 6500.0, 1.0), (0, 2000.0, 0.5), (1, 0.7, 0), (2, 1.0, 0), (3, 0.9, 0), (4, 1.0, 0),
                             (5, 1e-4, 0)]:
            out = np.zeros(n, np.float32)
            assert L.srth_spectrum(kind, n, a0, a1, out.ctypes.data_as(C.POINTER(C.c_float))) == 0
            assert np.array_equal(out, oracle.spectrum(kind, n, a0, a1))
    # spectrum.rs:832-869
    assert abs(L.srth_black_body(500.0, 5000.0) - 12107.190590398) / 12107.19 < 1e-4
    assert np.isnan(L.srth_black_body(500.0, -1.0))  # the reference panics (spectrum.rs:583-584)
    out = np.zeros(12, np.float32)
    assert L.srth_spectrum(1, 12, 1.0, 0, out.ctypes.data_as(C.POINTER(C.c_float))) == -1  # spectrum.rs:37


def test_prism_preset_is_cornell_plus_glass():
    f = scenes.preset("prism", 32)
    c = scenes.preset("cornell", 32)
    assert len(f.objects) == len(c.objects) + 1
    glass = f.materials[int(f.objects[-1, 22])]
    assert glass[3] == 1 and glass[4] == np.float32(1.30) and glass[5] == 6000.0


def test_no_gpu_means_loud_failure_not_fallback():
    lib = srt.native.lib()
    if lib.srt_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(srt.SrtError) as e:
        srt.Renderer(scenes.preset("cornell"), 16, 16)
    assert e.value.code == srt.native.SRT_ERR_CUDA and "no CPU fallback" in e.value.message
    with pytest.raises(srt.SrtError):
        srt.spectrum_to_rgb(np.ones(32, np.float32))
    with pytest.raises(srt.SrtError):
        scenes.dispatch_render("cornell", 16, 16, 1)


def test_validation_happens_before_the_device_is_touched():
    """The reference panics on these (main.rs:1407-1412, spectrum.rs:37-38); srt_create returns a status."""
    import dataclasses
    flat = scenes.preset("cornell")
    cam = flat.camera.copy()
    cam[6:9] = 2.0 * cam[3:6]
    with pytest.raises(srt.SrtError) as e:
        srt.Renderer(dataclasses.replace(flat, camera=cam), 8, 8)
    assert e.value.code == srt.native.SRT_ERR_CAMERA_COLLINEAR
    with pytest.raises(srt.SrtError) as e:
        srt.Renderer(dataclasses.replace(flat, n_lambda=20, spectra=flat.spectra[:, :20]), 8, 8)
    assert e.value.code == srt.native.SRT_ERR_SPECTRUM_SAMPLES
    with pytest.raises(srt.SrtError) as e:
        srt.Renderer(flat, 8, 8, max_bounces=200)
    assert e.value.code == srt.native.SRT_ERR_UNSUPPORTED


def test_header_is_plain_c11_and_layout_is_pinned(tmp_path):
    """include/srt.h must compile as C (the Rust -sys crate / cgo / JNI side sees C, not C++), and the struct layout
    the mirrors rely on is asserted from C as well (the C++ side asserts the same numbers in srt_api.cu)."""
    import shutil
    import subprocess
    src = os.path.join(ROOT, "tests", "abi", "srt_h_c11.c")
    cc = shutil.which("gcc") or shutil.which("cc")
    assert cc, "no C compiler"
    out = tmp_path / "srt_h_c11.o"
    subprocess.check_call([cc, "-std=c11", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-c", src, "-o", str(out)])


def _checkpoint_header(**over):
    """A CkptHeader (srt_api.cu) with consistent sizes for a 4x2 image of 8 wavelengths and an empty scene."""
    import struct
    f = dict(magic=b"SRTCKPT1", version=1, header_bytes=96, width=4, height=2, n_lambda=8, n_objects=0, n_materials=0,
             n_lights=0, n_spectra=0, sizeof_params=60, sizeof_camera=40, sizeof_object=92, sizeof_material=24, sizeof_light=16,
             frames=1, scene_hash=0, payload_hash=0, payload_floats=4 * 2 * 8)
    f.update(over)
    return struct.pack("<8s15IQQQQ", f["magic"], f["version"], f["header_bytes"], f["width"], f["height"], f["n_lambda"],
                       f["n_objects"], f["n_materials"], f["n_lights"], f["n_spectra"], f["sizeof_params"], f["sizeof_camera"],
                       f["sizeof_object"], f["sizeof_material"], f["sizeof_light"], 0, f["frames"], f["scene_hash"],
                       f["payload_hash"], f["payload_floats"])


@pytest.mark.parametrize("case", ["truncated", "huge_spectra", "huge_image", "wrapping_image", "bad_lambda", "garbage"])
def test_checkpoint_open_rejects_hostile_files_without_allocating(tmp_path, case):
    """A truncated or hostile checkpoint must come back as an error code: sizes are validated against the file before
    anything is allocated from them, and nothing unwinds across the C boundary (no GPU needed: all of this happens
    before a device is touched)."""
    lib = srt.native.lib()
    body = b"\0" * (60 + 40)
    if case == "truncated":
        data = _checkpoint_header() + body            # payload missing
    elif case == "huge_spectra":
        data = _checkpoint_header(n_spectra=1 << 24, n_lambda=128, payload_floats=4 * 2 * 128) + body
    elif case == "huge_image":
        data = _checkpoint_header(width=1 << 20, height=1 << 20, payload_floats=(1 << 40) * 8) + body
    elif case == "wrapping_image":
        data = _checkpoint_header(width=0xFFFFFFFF, height=0xFFFFFFFF, payload_floats=((0xFFFFFFFF * 0xFFFFFFFF) * 8) & (2**64 - 1)) + body
    elif case == "bad_lambda":
        data = _checkpoint_header(n_lambda=12, payload_floats=4 * 2 * 12) + body + b"\0" * (4 * 2 * 12 * 4)
    else:
        data = os.urandom(4096)
    path = tmp_path / "bad.ckpt"
    path.write_bytes(data)
    h = C.c_void_p()
    rc = lib.srt_checkpoint_open(str(path).encode(), -1, C.byref(h))
    assert rc in (srt.native.SRT_ERR_INVALID_ARGUMENT, srt.native.SRT_ERR_UNSUPPORTED) and not h.value
    assert b"checkpoint" in lib.srt_last_error(None)


def test_rust_binding_matches_header():
    """ffi/src/lib.rs (the -sys side of the boundary; it cannot be compiled here -- no Rust toolchain) must declare
    exactly the functions of include/srt.h, and mirror every struct field for field in the header's order."""
    header = open(os.path.join(ROOT, "include", "srt.h")).read()
    rust = open(os.path.join(ROOT, "ffi", "src", "lib.rs")).read()
    declared = set(re.findall(r"\b(srt_[a-z0-9_]+)\s*\(", header))
    bound = set(re.findall(r"pub fn (srt_[a-z0-9_]+)\s*\(", rust))
    assert declared == bound, (sorted(declared - bound), sorted(bound - declared))
    for name in ("srt_object", "srt_material", "srt_light", "srt_camera", "srt_params", "srt_counters"):
        c_body = re.search(r"typedef struct " + name + r" \{(.*?)\} " + name + ";", header, re.S).group(1)
        c_body = re.sub(r"/\*.*?\*/", "", c_body, flags=re.S)
        c_fields = []
        for decl in c_body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            ctype, names = decl.split(None, 1)
            for n in names.split(","):
                m = re.match(r"\s*([a-z0-9_]+)(?:\[(\d+)\])?", n)
                c_fields.append((m.group(1), ctype, int(m.group(2) or 1)))
        r_body = re.search(r"pub struct " + name + r" \{(.*?)\n    \}", rust, re.S).group(1)
        r_fields = []
        for m in re.finditer(r"pub ([a-z0-9_]+): (\[(\w+); (\d+)\]|\w+)", r_body):
            r_fields.append((m.group(1), m.group(3) or m.group(2), int(m.group(4) or 1)))
        tmap = {"float": "f32", "uint32_t": "u32", "int32_t": "i32", "uint64_t": "u64"}
        assert [(n, tmap[t], k) for n, t, k in c_fields] == r_fields, name
    # the status codes and enums carry the header's values
    for cname, value in re.findall(r"(SRT_[A-Z0-9_]+) = (\d+)", header):
        m = re.search(r"pub const " + cname + r": \w+ = (\d+);", rust)
        assert m and int(m.group(1)) == int(value), cname


@pytest.mark.parametrize("w,h", [(1, 1), (7, 5), (400, 300), (1920, 35)])   # (the last one: more than one 64 KB deflate block)
def test_host_png_export_round_trips(tmp_path, oracle, w, h):
    """The PNG the headless entry saves (no image library: stored-deflate zlib stream, CRC-32, Adler-32) read back by
    an independent PNG decoder equals the reference's RGBA8 conversion (custom_image.rs:92-101: clamp, * 255, truncate,
    NaN -> 0) of the same CustomImage data."""
    from PIL import Image
    rng = np.random.default_rng(w * 1000 + h)
    img = rng.uniform(-0.2, 1.3, (h, w, 4)).astype(np.float32)
    img[..., 3] = 1.0
    img.reshape(-1)[::97] = np.nan
    path = tmp_path / "x.png"
    scenes.save_png(path, img)
    im = Image.open(path)
    assert im.mode == "RGBA" and im.size == (w, h)
    assert np.array_equal(np.asarray(im), oracle.to_rgba8(img))
