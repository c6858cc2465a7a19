//! The headless sibling of `App::dispatch_render` (main.rs:1376-1427): the same preparation, but the frames are
//! rendered by libsrt.so instead of `thread::spawn(App::render)`, and the function returns the `CustomImage` data the
//! Display tab would show.
//!
//! The reference keeps its own scene types; the glue that turns them into a [`FlatScene`] touches `pub(crate)`
//! fields and therefore lives inside the reference (`impl RaytracingUniforms { fn flatten() }` in `shader.rs`, and
//! `App::dispatch_render_headless` in `main.rs`): INTEGRATION.md section 3 has that patch.  Everything that does not
//! need the reference's types is here.

use crate::sys;
use crate::{FlatScene, Renderer, SrtError};

/// What `dispatch_render` takes from `UIFields` (main.rs:1511-1535) besides the scene.
#[derive(Clone, Copy, Debug)]
pub struct RenderSettings {
    pub width: u32,
    pub height: u32,
    /// `nbr_of_iterations` = frames = samples per pixel (main.rs:1338)
    pub iterations: u32,
    /// `uniforms.max_bounces` (main.rs:1403)
    pub max_bounces: u32,
    /// `SRT_RNG_PCG3D_REFERENCE` reproduces the reference's samples; `SRT_RNG_PHILOX` is the backend's own RNG
    pub rng_mode: u32,
    /// CUDA device ordinal, -1 = current
    pub device: i32,
    /// frames per progress update (the reference updates after every frame; larger batches render faster)
    pub frames_per_update: u32,
}

impl Default for RenderSettings {
    fn default() -> Self {
        // UIFields::default(): 600x400 (main.rs:1733-1757), 100 iterations / 30 bounces (main.rs:29-35)
        RenderSettings { width: 600, height: 400, iterations: 100, max_bounces: 30, rng_mode: sys::SRT_RNG_PCG3D_REFERENCE,
                         device: -1, frames_per_update: 16 }
    }
}

/// VISIBLE_LIGHT_WAVELENGTH_LOWER_BOUND / UPPER_BOUND, spectrum.rs:5-6
pub const LAMBDA_MIN: f32 = 380.0;
pub const LAMBDA_MAX: f32 = 780.0;

pub fn params_for(scene: &FlatScene, s: &RenderSettings) -> sys::srt_params {
    sys::srt_params {
        width: s.width,
        height: s.height,
        n_lambda: scene.n_lambda,
        lambda_min: LAMBDA_MIN,
        lambda_max: LAMBDA_MAX,
        max_bounces: s.max_bounces,
        intended_frames: s.iterations,
        rng_mode: s.rng_mode,
        math_mode: sys::SRT_MATH_FAST,
        accel: sys::SRT_ACCEL_AUTO,
        integrator: sys::SRT_INTEGRATOR_AUTO,
        device: s.device,
        pool_paths: 0,
        philox_seed_lo: 0,
        philox_seed_hi: 0,
    }
}

/// Progress of a headless render, the counterpart of `AppActions::RenderingProgressUpdate` (main.rs:1346-1347).
pub type Progress<'a> = &'a mut dyn FnMut(f32) -> bool;

/// Renders `settings.iterations` frames of `scene` and returns `CustomImage.data` (W*H*4 f32, alpha 1).
/// `progress(fraction)` is called after every batch and returns `true` to abort (AppToRenderMessages::AbortRender,
/// main.rs:1351-1357); an aborted render returns the image of the frames completed so far.
pub fn dispatch_render_headless(scene: &FlatScene, settings: &RenderSettings, progress: Option<Progress>) -> Result<Vec<f32>, SrtError> {
    let params = params_for(scene, settings);
    // srt_create validates what dispatch_render asserts: camera direction / up not collinear (main.rs:1407-1412),
    // sample count a multiple of 8 up to 128 (spectrum.rs:37-38), non-empty image
    let mut r = Renderer::new(scene, &params)?;
    let per = settings.frames_per_update.max(1);
    match progress {
        Some(p) => {
            r.render_progressive(0, settings.iterations, per, false, |done, total, _| p(done as f32 / total as f32))?;
        }
        None => r.render_frames(0, settings.iterations)?,
    }
    r.resolve_rgba_f32()
}

/// ... and as the RGBA8 image `DynamicImage::from(custom_image)` would hold (custom_image.rs:92-101), ready for
/// `image::RgbaImage::from_raw(w, h, bytes)` + `.save(path)` (main.rs:2325-2326).
pub fn dispatch_render_headless_rgba8(scene: &FlatScene, settings: &RenderSettings) -> Result<Vec<u8>, SrtError> {
    let params = params_for(scene, settings);
    let mut r = Renderer::new(scene, &params)?;
    r.render_frames(0, settings.iterations)?;
    r.resolve_rgba_u8()
}
