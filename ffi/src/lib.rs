//! Bindings for `libsrt.so`, the B200 spectral render backend (C ABI: `include/srt.h`).
//!
//! * [`sys`] -- the raw `extern "C"` block and `#[repr(C)]` mirrors, one to one with the header.
//! * [`FlatScene`], [`Renderer`] -- a thin safe wrapper: scene in, `CustomImage.data` out.
//! * [`headless`] -- `dispatch_render_headless`, the sibling of `App::dispatch_render` (main.rs:1376-1427).
//!
//! Conventions of the C side (see the header): every call returns a status (`0 == SRT_OK`), nothing unwinds across
//! the boundary, a context is driven by one thread at a time (`srt_abort` excepted), and there is no CPU fallback.
#![allow(non_camel_case_types)]

pub mod headless;

pub mod sys {
    use std::os::raw::{c_char, c_int, c_void};

    pub const SRT_ABI_VERSION: u32 = 1;

    // srt_status
    pub const SRT_OK: c_int = 0;
    pub const SRT_ERR_INVALID_ARGUMENT: c_int = 1;
    pub const SRT_ERR_SPECTRUM_SAMPLES: c_int = 2;
    pub const SRT_ERR_CAMERA_COLLINEAR: c_int = 3;
    pub const SRT_ERR_CUDA: c_int = 4;
    pub const SRT_ERR_UNSUPPORTED: c_int = 5;
    pub const SRT_ERR_ABORTED: c_int = 6;

    // AABBType, shader.rs:168-172
    pub const SRT_PLAIN_BOX: u32 = 0;
    pub const SRT_SPHERE: u32 = 1;
    pub const SRT_ROTATED_BOX: u32 = 2;

    pub const SRT_RNG_PCG3D_REFERENCE: u32 = 0;
    pub const SRT_RNG_PHILOX: u32 = 1;
    pub const SRT_MATH_FAST: u32 = 0;
    pub const SRT_MATH_EXACT: u32 = 1;
    pub const SRT_ACCEL_AUTO: u32 = 0;
    pub const SRT_ACCEL_LINEAR: u32 = 1;
    pub const SRT_ACCEL_BVH: u32 = 2;
    pub const SRT_INTEGRATOR_WAVEFRONT: u32 = 0;
    pub const SRT_INTEGRATOR_RESIDENT: u32 = 1;
    pub const SRT_INTEGRATOR_AUTO: u32 = 2;

    /// Aabb, shader.rs:99-104 (92 bytes)
    #[repr(C)]
    #[derive(Clone, Copy, Debug, Default)]
    pub struct srt_object {
        pub min: [f32; 3],
        pub max: [f32; 3],
        pub kind: u32,
        pub center: [f32; 3],
        pub dims: [f32; 3],
        pub rot: [f32; 9],
        pub material: u32,
    }

    /// Material, shader.rs:253-258 (+ the dispersion extension) (24 bytes)
    #[repr(C)]
    #[derive(Clone, Copy, Debug, Default)]
    pub struct srt_material {
        pub metallicness: f32,
        pub roughness: f32,
        pub reflectance: u32,
        pub transmissive: u32,
        pub ior_a: f32,
        pub ior_b: f32,
    }

    /// Light, shader.rs:192-195 (16 bytes)
    #[repr(C)]
    #[derive(Clone, Copy, Debug, Default)]
    pub struct srt_light {
        pub position: [f32; 3],
        pub spectrum: u32,
    }

    /// Camera, shader.rs:213-218 (40 bytes)
    #[repr(C)]
    #[derive(Clone, Copy, Debug, Default)]
    pub struct srt_camera {
        pub position: [f32; 3],
        pub direction: [f32; 3],
        pub up: [f32; 3],
        pub fov_y_deg: f32,
    }

    /// RaytracingUniforms minus the scene vectors + image size + backend knobs (60 bytes)
    #[repr(C)]
    #[derive(Clone, Copy, Debug, Default)]
    pub struct srt_params {
        pub width: u32,
        pub height: u32,
        pub n_lambda: u32,
        pub lambda_min: f32,
        pub lambda_max: f32,
        pub max_bounces: u32,
        pub intended_frames: u32,
        pub rng_mode: u32,
        pub math_mode: u32,
        pub accel: u32,
        pub integrator: u32,
        pub device: i32,
        pub pool_paths: u32,
        pub philox_seed_lo: u32,
        pub philox_seed_hi: u32,
    }

    /// Event counters (104 bytes)
    #[repr(C)]
    #[derive(Clone, Copy, Debug, Default)]
    pub struct srt_counters {
        pub samples: u64,
        pub rays_primary: u64,
        pub rays_continuation: u64,
        pub rays_shadow: u64,
        pub hits: u64,
        pub self_hits: u64,
        pub misses: u64,
        pub lit: u64,
        pub spec_hits: u64,
        pub spec_dropped: u64,
        pub iterations: u64,
        pub kernel_launches: u64,
        pub shadow_skipped: u64,
    }

    #[repr(C)]
    pub struct srt_ctx {
        _private: [u8; 0],
    }

    pub type srt_progress_fn =
        Option<unsafe extern "C" fn(user: *mut c_void, frames_done: u32, frames_total: u32, rgba8: *const u8) -> c_int>;

    // layout pins (the C side asserts the same numbers: spectral_raytracer_b200/csrc/srt_api.cu, tests/abi/srt_h_c11.c)
    const _: () = assert!(std::mem::size_of::<srt_object>() == 92);
    const _: () = assert!(std::mem::size_of::<srt_material>() == 24);
    const _: () = assert!(std::mem::size_of::<srt_light>() == 16);
    const _: () = assert!(std::mem::size_of::<srt_camera>() == 40);
    const _: () = assert!(std::mem::size_of::<srt_params>() == 60);
    const _: () = assert!(std::mem::size_of::<srt_counters>() == 104);

    extern "C" {
        pub fn srt_abi_version() -> u32;
        pub fn srt_device_count() -> c_int;
        pub fn srt_create(
            params: *const srt_params,
            camera: *const srt_camera,
            objects: *const srt_object,
            n_objects: u32,
            materials: *const srt_material,
            n_materials: u32,
            lights: *const srt_light,
            n_lights: u32,
            spectra: *const f32,
            n_spectra: u32,
            out: *mut *mut srt_ctx,
        ) -> c_int;
        pub fn srt_destroy(ctx: *mut srt_ctx);
        pub fn srt_last_error(ctx: *const srt_ctx) -> *const c_char;
        pub fn srt_render_frames(ctx: *mut srt_ctx, first_frame: u32, n_frames: u32) -> c_int;
        pub fn srt_render_progressive(
            ctx: *mut srt_ctx,
            first_frame: u32,
            n_frames: u32,
            frames_per_update: u32,
            want_preview: c_int,
            callback: srt_progress_fn,
            user: *mut c_void,
        ) -> c_int;
        pub fn srt_abort(ctx: *mut srt_ctx) -> c_int;
        pub fn srt_set_deterministic(ctx: *mut srt_ctx, on: c_int) -> c_int;
        pub fn srt_clear(ctx: *mut srt_ctx) -> c_int;
        pub fn srt_frames_accumulated(ctx: *const srt_ctx) -> u64;
        pub fn srt_set_frames_accumulated(ctx: *mut srt_ctx, n_frames: u64) -> c_int;
        pub fn srt_accum_device_ptr(ctx: *mut srt_ctx, n_floats: *mut usize) -> *mut c_void;
        pub fn srt_device(ctx: *const srt_ctx) -> c_int;
        pub fn srt_stream(ctx: *mut srt_ctx) -> *mut c_void;
        pub fn srt_read_accum(ctx: *mut srt_ctx, out: *mut f32) -> c_int;
        pub fn srt_write_accum(ctx: *mut srt_ctx, data: *const f32, n_frames: u64) -> c_int;
        pub fn srt_get_params(ctx: *const srt_ctx, out: *mut srt_params) -> c_int;
        pub fn srt_checkpoint_save(ctx: *mut srt_ctx, path: *const c_char) -> c_int;
        pub fn srt_checkpoint_load(ctx: *mut srt_ctx, path: *const c_char) -> c_int;
        pub fn srt_checkpoint_open(path: *const c_char, device: i32, out: *mut *mut srt_ctx) -> c_int;
        pub fn srt_resolve_rgba_f32(ctx: *mut srt_ctx, out: *mut f32) -> c_int;
        pub fn srt_resolve_rgba_u8(ctx: *mut srt_ctx, out: *mut u8) -> c_int;
        pub fn srt_resolve_rgba_f32_device(ctx: *mut srt_ctx, d_out: *mut f32) -> c_int;
        pub fn srt_primary_ids(ctx: *mut srt_ctx, frame: u32, ids: *mut i32, t: *mut f32) -> c_int;
        pub fn srt_spectrum_to_rgb(spectra: *const f32, n: u32, n_lambda: u32, lambda_min: f32, lambda_max: f32, rgb: *mut f32) -> c_int;
        pub fn srt_spectra_resample(input: *const f32, n: u32, n_old: u32, n_new: u32, out: *mut f32) -> c_int;
        pub fn srt_spectra_radiance(input: *const f32, n: u32, n_lambda: u32, lambda_min: f32, lambda_max: f32, radiance: *mut f32) -> c_int;
        pub fn srt_spectra_normalize(input: *const f32, n: u32, n_lambda: u32, lambda_min: f32, lambda_max: f32, out: *mut f32) -> c_int;
        pub fn srt_selftest_arith(n: u64, seed: u32, mismatches: *mut u64) -> c_int;
        pub fn srt_get_counters(ctx: *mut srt_ctx, out: *mut srt_counters) -> c_int;
        pub fn srt_reset_counters(ctx: *mut srt_ctx) -> c_int;
        pub fn srt_last_render_stats(ctx: *mut srt_ctx, device_ms: *mut f32, kernel_launches: *mut u64) -> c_int;
        pub fn srt_launch_param_bytes() -> u32;
        pub fn srt_set_profiling(ctx: *mut srt_ctx, on: c_int) -> c_int;
        pub fn srt_last_stage_times(ctx: *mut srt_ctx, ms: *mut f32, launches: *mut u64) -> c_int;
    }

    // libsrt_nccl.so (single-process multi-device NCCL reduce of the accumulation buffers)
    #[cfg(feature = "nccl")]
    extern "C" {
        pub fn srt_reduce(ctxs: *const *mut srt_ctx, n: u32) -> c_int;
        pub fn srt_reduce_shutdown();
        pub fn srt_reduce_last_ms() -> f32;
        pub fn srt_reduce_last_error() -> *const c_char;
    }
}

use std::ffi::CStr;
use std::os::raw::{c_int, c_void};

/// An error of the C side: status code + `srt_last_error` text.
#[derive(Debug, Clone)]
pub struct SrtError {
    pub code: c_int,
    pub message: String,
}

impl std::fmt::Display for SrtError {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "srt error {}: {}", self.code, self.message)
    }
}

impl std::error::Error for SrtError {}

fn last_error(ctx: *const sys::srt_ctx, code: c_int) -> SrtError {
    let p = unsafe { sys::srt_last_error(ctx) };
    let message = if p.is_null() { String::new() } else { unsafe { CStr::from_ptr(p) }.to_string_lossy().into_owned() };
    SrtError { code, message }
}

/// The flattened `RaytracingUniforms` (shader.rs:32-41): what `srt_create` takes.  `spectra` holds `n_lambda`
/// floats per row; materials and lights index its rows.
#[derive(Clone, Debug, Default)]
pub struct FlatScene {
    pub n_lambda: u32,
    pub camera: sys::srt_camera,
    pub objects: Vec<sys::srt_object>,
    pub materials: Vec<sys::srt_material>,
    pub lights: Vec<sys::srt_light>,
    pub spectra: Vec<f32>,
}

impl FlatScene {
    /// Appends one spectrum (the first `n_lambda` intensities) and returns its row index.
    pub fn push_spectrum(&mut self, intensities: &[f32]) -> u32 {
        let n = self.n_lambda as usize;
        assert!(intensities.len() >= n, "spectrum shorter than n_lambda");
        self.spectra.extend_from_slice(&intensities[..n]);
        (self.spectra.len() / n - 1) as u32
    }
}

/// One `srt_ctx`: create (validate + upload) -> render_frames -> resolve, the life cycle of
/// `App::dispatch_render` + `App::render` (main.rs:1376-1427, :1327-1371).
pub struct Renderer {
    ctx: *mut sys::srt_ctx,
    pub width: u32,
    pub height: u32,
    pub n_lambda: u32,
}

// A context may move between threads; it is not Sync (one driving thread at a time, srt_abort excepted).
unsafe impl Send for Renderer {}

impl Renderer {
    pub fn new(scene: &FlatScene, params: &sys::srt_params) -> Result<Renderer, SrtError> {
        let n = params.n_lambda as usize;
        let n_spectra = if n == 0 { 0 } else { scene.spectra.len() / n };
        let mut ctx: *mut sys::srt_ctx = std::ptr::null_mut();
        let rc = unsafe {
            sys::srt_create(
                params,
                &scene.camera,
                scene.objects.as_ptr(),
                scene.objects.len() as u32,
                scene.materials.as_ptr(),
                scene.materials.len() as u32,
                scene.lights.as_ptr(),
                scene.lights.len() as u32,
                scene.spectra.as_ptr(),
                n_spectra as u32,
                &mut ctx,
            )
        };
        if rc != sys::SRT_OK {
            return Err(last_error(std::ptr::null(), rc));
        }
        Ok(Renderer { ctx, width: params.width, height: params.height, n_lambda: params.n_lambda })
    }

    fn check(&self, rc: c_int) -> Result<(), SrtError> {
        if rc == sys::SRT_OK { Ok(()) } else { Err(last_error(self.ctx, rc)) }
    }

    /// Frames `[first_frame, first_frame + n_frames)`: one sample per pixel and frame, radiance ADDED to the buffer.
    pub fn render_frames(&mut self, first_frame: u32, n_frames: u32) -> Result<(), SrtError> {
        let rc = unsafe { sys::srt_render_frames(self.ctx, first_frame, n_frames) };
        self.check(rc)
    }

    /// `App::render`'s per-frame protocol in batches (main.rs:1338-1357).  `on_update(frames_done, frames_total,
    /// rgba8)` returns `true` to abort.  `Ok(true)`: the render was aborted.
    pub fn render_progressive<F>(&mut self, first_frame: u32, n_frames: u32, frames_per_update: u32, want_preview: bool,
                                 mut on_update: F) -> Result<bool, SrtError>
    where
        F: FnMut(u32, u32, Option<&[u8]>) -> bool,
    {
        struct Env<'a> {
            f: &'a mut dyn FnMut(u32, u32, Option<&[u8]>) -> bool,
            bytes: usize,
        }
        unsafe extern "C" fn trampoline(user: *mut c_void, done: u32, total: u32, rgba8: *const u8) -> c_int {
            let env = &mut *(user as *mut Env);
            let img = if rgba8.is_null() { None } else { Some(std::slice::from_raw_parts(rgba8, env.bytes)) };
            // a panic must not unwind into C
            match std::panic::catch_unwind(std::panic::AssertUnwindSafe(|| (env.f)(done, total, img))) {
                Ok(abort) => abort as c_int,
                Err(_) => 1,
            }
        }
        let mut env = Env { f: &mut on_update, bytes: (self.width as usize) * (self.height as usize) * 4 };
        let rc = unsafe {
            sys::srt_render_progressive(self.ctx, first_frame, n_frames, frames_per_update, want_preview as c_int,
                                        Some(trampoline), &mut env as *mut Env as *mut c_void)
        };
        if rc == sys::SRT_ERR_ABORTED {
            return Ok(true);
        }
        self.check(rc).map(|_| false)
    }

    /// A handle other threads can use to stop a running render at the next frame boundary (srt_abort).
    pub fn abort_handle(&self) -> AbortHandle {
        AbortHandle { ctx: self.ctx }
    }

    pub fn clear(&mut self) -> Result<(), SrtError> {
        let rc = unsafe { sys::srt_clear(self.ctx) };
        self.check(rc)
    }

    pub fn set_deterministic(&mut self, on: bool) -> Result<(), SrtError> {
        let rc = unsafe { sys::srt_set_deterministic(self.ctx, on as c_int) };
        self.check(rc)
    }

    pub fn frames_accumulated(&self) -> u64 {
        unsafe { sys::srt_frames_accumulated(self.ctx) }
    }

    /// `CustomImage.data` (custom_image.rs:9-13): W*H*4 f32, alpha = 1, linear RGB before clamping.
    pub fn resolve_rgba_f32(&mut self) -> Result<Vec<f32>, SrtError> {
        let mut out = vec![0.0f32; (self.width as usize) * (self.height as usize) * 4];
        let rc = unsafe { sys::srt_resolve_rgba_f32(self.ctx, out.as_mut_ptr()) };
        self.check(rc).map(|_| out)
    }

    /// `From<CustomImage> for DynamicImage` (custom_image.rs:92-101): clamp, * 255, truncating cast.
    pub fn resolve_rgba_u8(&mut self) -> Result<Vec<u8>, SrtError> {
        let mut out = vec![0u8; (self.width as usize) * (self.height as usize) * 4];
        let rc = unsafe { sys::srt_resolve_rgba_u8(self.ctx, out.as_mut_ptr()) };
        self.check(rc).map(|_| out)
    }

    pub fn counters(&mut self) -> Result<sys::srt_counters, SrtError> {
        let mut c = sys::srt_counters::default();
        let rc = unsafe { sys::srt_get_counters(self.ctx, &mut c) };
        self.check(rc).map(|_| c)
    }

    pub fn save_checkpoint(&mut self, path: &std::path::Path) -> Result<(), SrtError> {
        let p = std::ffi::CString::new(path.to_string_lossy().as_bytes()).map_err(|_| SrtError { code: 1, message: "path contains NUL".into() })?;
        let rc = unsafe { sys::srt_checkpoint_save(self.ctx, p.as_ptr()) };
        self.check(rc)
    }

    pub fn as_raw(&self) -> *mut sys::srt_ctx {
        self.ctx
    }
}

impl Drop for Renderer {
    fn drop(&mut self) {
        unsafe { sys::srt_destroy(self.ctx) }
    }
}

/// See [`Renderer::abort_handle`].  Must not outlive the renderer it came from.
#[derive(Clone, Copy)]
pub struct AbortHandle {
    ctx: *mut sys::srt_ctx,
}

unsafe impl Send for AbortHandle {}
unsafe impl Sync for AbortHandle {}

impl AbortHandle {
    pub fn abort(&self) {
        unsafe {
            sys::srt_abort(self.ctx);
        }
    }
}

/// Stateless `Spectrum::get_rgb_early` (spectrum.rs:238-261) for `spectra.len() / n_lambda` spectra.
pub fn spectrum_to_rgb(spectra: &[f32], n_lambda: u32, lambda_min: f32, lambda_max: f32) -> Result<Vec<[f32; 3]>, SrtError> {
    let n = spectra.len() / n_lambda.max(1) as usize;
    let mut rgb = vec![[0.0f32; 3]; n];
    let rc = unsafe { sys::srt_spectrum_to_rgb(spectra.as_ptr(), n as u32, n_lambda, lambda_min, lambda_max, rgb.as_mut_ptr() as *mut f32) };
    if rc == sys::SRT_OK { Ok(rgb) } else { Err(last_error(std::ptr::null(), rc)) }
}
