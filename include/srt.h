/* srt.h -- C ABI of the B200 spectral render backend (libsrt.so).
 *
 * Drop-in for the per-pixel render path of happy737/spectral-raytracer: everything
 * below App::render (main.rs:1327) -- apply_shader2 (main.rs:1280-1322), the whole
 * of shader.rs (ray generation, "ray acceleration structure" = submit_ray,
 * intersection, the implicit any-hit logic, hit and miss shaders), the per-sample
 * part of spectrum.rs (get_rgb_early, spectrum.rs:238-261) and the accumulation
 * image of custom_image.rs (blend_pixel :59-79, RGBA8 export :92-101).
 *
 * The reference has no FFI of its own; this header is the seam a
 * Rust `-sys` crate binds (see INTEGRATION.md for the `extern "C"` block, the
 * #[repr(C)] mirrors and the headless entry beside App::dispatch_render,
 * main.rs:1376).  All structs are plain old data, 4-byte aligned, no padding.
 *
 * Conventions
 *   - every call returns an int status: 0 = SRT_OK, otherwise an srt_status code;
 *     srt_last_error(ctx) (or srt_last_error(NULL) for create / stateless calls)
 *     returns a message.  Nothing unwinds across this boundary (the reference
 *     panics instead: main.rs:1412, shader.rs:330/481, spectrum.rs asserts).
 *   - a context is driven by one host thread at a time (the reference renders from
 *     one dedicated thread, main.rs:1424); contexts are independent.
 *   - the library copies the scene at srt_create and owns all device memory; the
 *     caller owns every host buffer it passes in.
 *   - there is NO CPU fallback: if no CUDA device is usable every entry point
 *     that needs one fails with SRT_ERR_CUDA.
 */
#ifndef SRT_H
#define SRT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRT_ABI_VERSION 1u

typedef enum srt_status {
    SRT_OK = 0,
    SRT_ERR_INVALID_ARGUMENT = 1, /* null pointer, zero size, index out of range            */
    SRT_ERR_SPECTRUM_SAMPLES = 2, /* n_lambda not a multiple of 8 in 8..=128 (spectrum.rs:37-38) */
    SRT_ERR_CAMERA_COLLINEAR = 3, /* direction x up ~ 0 (main.rs:1407-1412, :2200-2203)     */
    SRT_ERR_CUDA = 4,             /* no device / launch failure / out of memory             */
    SRT_ERR_UNSUPPORTED = 5,      /* e.g. more lights than SRT_MAX_LIGHTS                   */
    SRT_ERR_ABORTED = 6           /* srt_abort() was called while rendering (main.rs:1351-1357) */
} srt_status;

/* AABBType, shader.rs:168-172 */
enum { SRT_PLAIN_BOX = 0, SRT_SPHERE = 1, SRT_ROTATED_BOX = 2 };

/* Aabb, shader.rs:99-104.  min/max are exactly what Aabb::new_sphere / new_box /
 * new_rotated_box compute (shader.rs:108-166); center/dims/rot are the RotatedBox
 * payload (row-major Rotation3), ignored for the other kinds. */
typedef struct srt_object {
    float min[3];
    float max[3];
    uint32_t kind;
    float center[3];
    float dims[3];
    float rot[9];
    uint32_t material; /* index into the material table */
} srt_object;

/* Material, shader.rs:253-258.  `reflectance` indexes the spectra table; the
 * reflective spectrum must already be clamped by min1 (spectrum.rs:486-494).
 * transmissive / ior_a / ior_b are the beyond-reference dispersion extension
 * (Cauchy n(lambda) = ior_a + ior_b / lambda_nm^2); transmissive = 0 gives the
 * reference's behaviour. */
typedef struct srt_material {
    float metallicness;
    float roughness;
    uint32_t reflectance;
    uint32_t transmissive;
    float ior_a;
    float ior_b;
} srt_material;

/* Light, shader.rs:192-195 */
typedef struct srt_light {
    float position[3];
    uint32_t spectrum; /* index into the spectra table (raw emission, not clamped) */
} srt_light;

/* Camera, shader.rs:213-218 */
typedef struct srt_camera {
    float position[3];
    float direction[3];
    float up[3];
    float fov_y_deg;
} srt_camera;

enum { SRT_RNG_PCG3D_REFERENCE = 0, SRT_RNG_PHILOX = 1 };
enum { SRT_MATH_FAST = 0, SRT_MATH_EXACT = 1 };
enum { SRT_ACCEL_AUTO = 0, SRT_ACCEL_LINEAR = 1, SRT_ACCEL_BVH = 2 };
/* WAVEFRONT: per-stage kernels over SoA path pools in HBM, compaction between bounces.  RESIDENT: one
 * persistent kernel, a path stays in its lane with its state in registers / shared memory (every legal n_lambda
 * for linear-scan scenes; BVH scenes at the default n_lambda = 32 only -- otherwise the wavefront is used).
 * AUTO: resident for linear-scan scenes, wavefront for BVH scenes (latency-bound traversal wants the wavefront's
 * higher occupancy). */
enum { SRT_INTEGRATOR_WAVEFRONT = 0, SRT_INTEGRATOR_RESIDENT = 1, SRT_INTEGRATOR_AUTO = 2 };

/* Per-render constants: RaytracingUniforms minus the scene vectors
 * (shader.rs:32-41) plus the image size (custom_image.rs:18-22) and backend knobs. */
typedef struct srt_params {
    uint32_t width, height;
    uint32_t n_lambda;            /* Spectrum::nbr_of_samples, spectrum.rs:25-30  */
    float lambda_min, lambda_max; /* EquidistantSamples(min,max); 380/780, spectrum.rs:5-6 */
    uint32_t max_bounces;         /* uniforms.max_bounces, main.rs:1403           */
    uint32_t intended_frames;     /* uniforms.intended_frames_amount, main.rs:1400 */
    uint32_t rng_mode;            /* SRT_RNG_*: pcg3d(px,py,frame+remaining) of shader.rs:389-391, or Philox4x32-10 keyed (pixel,frame,bounce) */
    uint32_t math_mode;           /* SRT_MATH_FAST: CUDA f32 libm; SRT_MATH_EXACT: correctly rounded sin/cos/asin (sample-exact vs the oracle) */
    uint32_t accel;               /* SRT_ACCEL_* */
    uint32_t integrator;          /* SRT_INTEGRATOR_* */
    int32_t device;               /* CUDA device ordinal; -1 = current device     */
    uint32_t pool_paths;          /* path-pool capacity; 0 = default              */
    uint32_t philox_seed_lo, philox_seed_hi;
} srt_params;

/* Event counters (device-side, accumulated over all render calls of a context). */
typedef struct srt_counters {
    uint64_t samples;           /* ray_generation_shader evaluations            */
    uint64_t rays_primary;
    uint64_t rays_continuation;
    uint64_t rays_shadow;
    uint64_t hits;              /* hit_shader evaluations                       */
    uint64_t self_hits;         /* hits with t < 1e-4                           */
    uint64_t misses;
    uint64_t lit;               /* unoccluded light samples                     */
    uint64_t spec_hits;
    uint64_t spec_dropped;      /* specular children dropped by the 1e-4 gate (shader.rs:407) */
    uint64_t iterations;        /* wavefront iterations launched                */
    uint64_t kernel_launches;   /* kernels launched by this context             */
    uint64_t shadow_skipped;    /* shadow rays NOT traced because their light term is exactly zero
                                   (light behind the surface / surface seen from behind); the reference
                                   traces them, rays_shadow counts only rays actually traced */
} srt_counters;

typedef struct srt_ctx srt_ctx;

uint32_t srt_abi_version(void);

/* Number of usable CUDA devices (0 if none; never falls back to the CPU). */
int srt_device_count(void);

/* Build a render context: validates like dispatch_render (main.rs:1376-1427),
 * copies the scene to the device, allocates the path pool and the W*H*n_lambda
 * spectral accumulation buffer.  spectra = n_spectra rows of n_lambda floats. */
int srt_create(const srt_params* params, const srt_camera* camera,
               const srt_object* objects, uint32_t n_objects,
               const srt_material* materials, uint32_t n_materials,
               const srt_light* lights, uint32_t n_lights,
               const float* spectra, uint32_t n_spectra, srt_ctx** out);
void srt_destroy(srt_ctx* ctx);
const char* srt_last_error(const srt_ctx* ctx);

/* App::render's frame loop (main.rs:1338-1341) for frame ids
 * [first_frame, first_frame + n_frames): every pixel gets one sample per frame and
 * the per-wavelength radiance is ADDED to the accumulation buffer.  Asynchronous
 * w.r.t. the host only inside the call; on return the work is complete. */
int srt_render_frames(srt_ctx* ctx, uint32_t first_frame, uint32_t n_frames);
/* App::render's per-frame protocol (main.rs:1338-1357) in batches: after every frames_per_update frames
 * (0 = 1, the reference's granularity) the callback receives frames_done / frames_total
 * (AppActions::RenderingProgressUpdate) and, if want_preview, the RGBA8 image of the frames accumulated so far
 * (AppActions::FrameUpdate, From<CustomImage> for DynamicImage, custom_image.rs:92-101; valid during the call
 * only); a nonzero return aborts (AppToRenderMessages::AbortRender, polled once per update like try_recv,
 * main.rs:1351).  The preview is copied out on a second stream while the next batch renders, and the callback
 * of update k runs while batch k+1 is on the GPU, so an abort takes effect one batch later; frames of completed
 * batches stay accumulated (srt_frames_accumulated) and the image is consistent.  Returns SRT_ERR_ABORTED when
 * the render stopped early.  The callback runs on the calling thread. */
typedef int (*srt_progress_fn)(void* user, uint32_t frames_done, uint32_t frames_total, const uint8_t* rgba8);
int srt_render_progressive(srt_ctx* ctx, uint32_t first_frame, uint32_t n_frames, uint32_t frames_per_update,
                           int want_preview, srt_progress_fn callback, void* user);
/* Request that a running / the next srt_render_frames / srt_render_progressive stops.  May be called from another
 * thread while a render call is in progress.  The render stops at a FRAME boundary: frames already started are
 * completed (the paths in flight are traced to their end), no further frame is begun, srt_frames_accumulated counts
 * exactly the frames the buffer holds, and the call returns SRT_ERR_ABORTED.  (The resident integrator renders a
 * call's frames in one launch: the flag is looked at between calls / batches there.) */
int srt_abort(srt_ctx* ctx);
/* Bit-reproducible accumulation.  Radiance is added to the buffer with f32 atomics; when one launch spans several
 * frames, different warps add to the same pixel record in an order that changes from run to run, so images agree
 * only to f32 rounding (a few ulp of the sum).  With deterministic = 1 srt_render_frames renders frame by frame (one
 * launch per frame): a pixel record then receives its terms in frame order, within a frame in bounce order, and two
 * runs -- or an interrupted and resumed run -- give identical bits.  Costs the tail of one launch per frame
 * (a few percent at 1080p).  Default 0. */
int srt_set_deterministic(srt_ctx* ctx, int on);
/* Zero the accumulation buffer and the frame count (a fresh CustomImage::new). */
int srt_clear(srt_ctx* ctx);
uint64_t srt_frames_accumulated(const srt_ctx* ctx);
/* For frame-sharded multi-GPU: after summing accumulation buffers externally
 * (NCCL reduce), tell the root context how many frames the sum now holds. */
int srt_set_frames_accumulated(srt_ctx* ctx, uint64_t n_frames);

/* Device pointer of the accumulation buffer (W*H*n_lambda f32, pixel-major,
 * top-left origin) for zero-copy use by a collective; *n_floats receives its size. */
void* srt_accum_device_ptr(srt_ctx* ctx, size_t* n_floats);
/* CUDA device ordinal the context lives on. */
int srt_device(const srt_ctx* ctx);
/* The CUDA stream (cudaStream_t) the context launches on. */
void* srt_stream(srt_ctx* ctx);
/* Copy the accumulation buffer to / from the host (checkpoint / resume, tests). */
int srt_read_accum(srt_ctx* ctx, float* out);
int srt_write_accum(srt_ctx* ctx, const float* in, uint64_t n_frames);

/* The per-render constants a context was created with (device = the ordinal it lives on). */
int srt_get_params(const srt_ctx* ctx, srt_params* out);

/* Checkpoint / resume.  save writes one file with the scene as given to srt_create, the render constants, the frame
 * count and the accumulation buffer (FNV-1a checksums; written aside and renamed).  load restores buffer and frame
 * count into a context of the SAME scene and constants (else SRT_ERR_INVALID_ARGUMENT; accel / integrator / device /
 * pool size may differ -- they do not change the image).  open builds a new context from the file alone (device = -1:
 * current device); render further frames with srt_render_frames(ctx, srt_frames_accumulated(ctx), n).  A resumed render
 * equals the uninterrupted one bit for bit in deterministic mode (srt_set_deterministic) or with one frame per call;
 * otherwise to f32 rounding of the sums.
 * (The reference keeps the image in memory only and lists scene saving as a TODO, main.rs:73.) */
int srt_checkpoint_save(srt_ctx* ctx, const char* path);
int srt_checkpoint_load(srt_ctx* ctx, const char* path);
int srt_checkpoint_open(const char* path, int32_t device, srt_ctx** out);

/* Sum the accumulation buffers of n contexts (one per device, same image size)
 * into ctxs[0] with NCCL (single-process multi-device).  Afterwards ctxs[0] holds
 * the whole render -- its frame count is the sum of all n frame counts -- and
 * ctxs[1..] are cleared (empty image, frame count 0), so the call can be repeated
 * after further srt_render_frames rounds without counting radiance twice.  The
 * NCCL communicators are created on the first call for a given set of devices and
 * reused; srt_reduce_shutdown destroys them.  srt_reduce_last_ms: device time of
 * the last reduce (CUDA events on ctxs[0]'s stream).  These four live in
 * libsrt_nccl.so. */
int srt_reduce(srt_ctx* const* ctxs, uint32_t n);
void srt_reduce_shutdown(void);
float srt_reduce_last_ms(void);
const char* srt_reduce_last_error(void);

/* mean spectrum -> XYZ -> RGB (get_rgb_early, spectrum.rs:238-261), alpha = 1:
 * out = W*H*4 f32, the layout of CustomImage.data (custom_image.rs:9-13). */
int srt_resolve_rgba_f32(srt_ctx* ctx, float* out);
/* ... followed by From<CustomImage> for DynamicImage (custom_image.rs:92-101):
 * clamp(0,1) * 255, truncating cast, NaN -> 0.  out = W*H*4 bytes. */
int srt_resolve_rgba_u8(srt_ctx* ctx, uint8_t* out);
/* Same two, leaving the result in device memory (device pointers). */
int srt_resolve_rgba_f32_device(srt_ctx* ctx, float* d_out);

/* Primary-hit object ids for one frame (ray generation + submit_ray's closest
 * hit, shader.rs:271-294 / :468-483): ids[W*H] = index into `objects`, -1 = miss;
 * t[W*H] (optional) = hit distance, +inf for a miss. */
int srt_primary_ids(srt_ctx* ctx, uint32_t frame, int32_t* ids, float* t);

/* Stateless Spectrum::get_rgb_early for n spectra of n_lambda samples each
 * (runs on the current CUDA device). rgb = n*3 floats. */
int srt_spectrum_to_rgb(const float* spectra, uint32_t n, uint32_t n_lambda,
                        float lambda_min, float lambda_max, float* rgb);

/* Spectrum tooling on the input side of the path (spectrum.rs:285-374), for n spectra at once, stateless, on the
 * current CUDA device.
 *   resample:  Spectrum::resample (spectrum.rs:285-323) from n_old to n_new samples per spectrum (in = n rows of
 *              n_old floats, out = n rows of n_new).  SRT_ERR_UNSUPPORTED for the reductions the reference panics
 *              on (a second trip of its down-sampling loop, spectrum.rs:298, or the assert of
 *              linear_interpolate_halved, spectrum.rs:616 -- e.g. 128 -> 24, 64 -> 8).
 *   radiance:  Spectrum::get_radiance (spectrum.rs:357-362), sum of I_i * step folded from 0 in sample order.
 *   normalize: Spectrum::normalize (spectrum.rs:369-374), every sample divided by max(r, g, b) of get_rgb_early. */
int srt_spectra_resample(const float* in, uint32_t n, uint32_t n_old, uint32_t n_new, float* out);
int srt_spectra_radiance(const float* in, uint32_t n, uint32_t n_lambda, float lambda_min, float lambda_max, float* radiance);
int srt_spectra_normalize(const float* in, uint32_t n, uint32_t n_lambda, float lambda_min, float lambda_max, float* out);

/* Device self-test of the kernels' exact-arithmetic helpers: the batched reciprocal / quotient sequences the
 * intersection and normalisation code uses instead of one IEEE operation per value are compared bit for bit
 * with the IEEE operations (1/x, a/b round-to-nearest; what Rust's f32 `/` does, shader.rs:531-556 and
 * nalgebra's normalize) on n pseudo-random operand sets; *mismatches receives the number of differing results. */
int srt_selftest_arith(uint64_t n, uint32_t seed, uint64_t* mismatches);

int srt_get_counters(srt_ctx* ctx, srt_counters* out);
int srt_reset_counters(srt_ctx* ctx);

/* Device time (ms, CUDA events on the context's stream) of the last
 * srt_render_frames call, and the number of kernels it launched. */
int srt_last_render_stats(srt_ctx* ctx, float* device_ms, uint64_t* kernel_launches);

/* Bytes of the scene block every kernel launch carries host -> device as its parameter (for e2e byte counts). */
uint32_t srt_launch_param_bytes(void);

/* Optional per-stage timing: when on, CUDA events are recorded around every stage
 * kernel and srt_last_stage_times returns, for the last srt_render_frames call, the
 * summed device time in ms and the launch count of ms[0] ray generation,
 * ms[1] extend (acceleration structure + intersection), ms[2] shade (hit / miss). */
int srt_set_profiling(srt_ctx* ctx, int on);
int srt_last_stage_times(srt_ctx* ctx, float* ms /* [3] */, uint64_t* launches /* [3] */);

#ifdef __cplusplus
}
#endif
#endif /* SRT_H */
