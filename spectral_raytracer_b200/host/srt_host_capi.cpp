// srt_host_capi.cpp -- extern "C" view of the C++ host mirror (srt_host.hpp) so that the
// Python plumbing (tests, bench.py) builds its scenes with the product's own builders, and
// can run the whole headless dispatch in one call.
#include <thread>

#include "srt_host.hpp"

using namespace srt_host;

namespace {
thread_local std::string g_err;
UIFields make_preset(const char* name, uint32_t n_lambda, uint32_t arg) {
    std::string n(name ? name : "");
    if (n == "default") return UIFields::default_scene(n_lambda);
    if (n == "cornell") return UIFields::cornell_box(n_lambda);
    if (n == "spheres") return UIFields::random_spheres(arg, n_lambda);
    if (n == "prism") return UIFields::prism(n_lambda);
    throw std::invalid_argument("unknown preset '" + n + "' (default | cornell | spheres | prism)");
}
}  // namespace

extern "C" {

struct srth_scene {
    FlatScene flat;
};

const char* srth_last_error(void) { return g_err.c_str(); }

// Build a preset's UIFields, run dispatch_render's uniform assembly and flatten it.
srth_scene* srth_scene_preset(const char* name, uint32_t n_lambda, uint32_t arg) {
    try {
        UIFields ui = make_preset(name, n_lambda, arg);
        RaytracingUniforms u = build_uniforms(ui);
        auto* s = new srth_scene();
        s->flat = flatten(u);
        return s;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
void srth_scene_free(srth_scene* s) { delete s; }

void srth_scene_counts(const srth_scene* s, uint32_t* n_lambda, uint32_t* n_obj, uint32_t* n_mat, uint32_t* n_light,
                       uint32_t* n_spec) {
    *n_lambda = s->flat.n_lambda;
    *n_obj = (uint32_t)s->flat.objects.size();
    *n_mat = (uint32_t)s->flat.materials.size();
    *n_light = (uint32_t)s->flat.lights.size();
    *n_spec = s->flat.n_spectra;
}
void srth_scene_copy(const srth_scene* s, srt_object* objects, srt_material* materials, srt_light* lights, float* spectra,
                     srt_camera* camera, float* lambda_min_max) {
    const FlatScene& f = s->flat;
    if (objects) std::memcpy(objects, f.objects.data(), f.objects.size() * sizeof(srt_object));
    if (materials) std::memcpy(materials, f.materials.data(), f.materials.size() * sizeof(srt_material));
    if (lights) std::memcpy(lights, f.lights.data(), f.lights.size() * sizeof(srt_light));
    if (spectra) std::memcpy(spectra, f.spectra.data(), f.spectra.size() * sizeof(float));
    if (camera) *camera = f.camera;
    if (lambda_min_max) {
        lambda_min_max[0] = f.lambda_min;
        lambda_min_max[1] = f.lambda_max;
    }
}

// Spectrum constructors (spectrum.rs:73-187) for known-answer tests:
// kind 0 temperature(arg0 = T, arg1 = factor), 1 flat(arg0), 2 red, 3 green, 4 blue, 5 sunlight(arg0 = factor)
int srth_spectrum(uint32_t kind, uint32_t n, float arg0, float arg1, float* out) {
    try {
        const float lo = VISIBLE_LIGHT_WAVELENGTH_LOWER_BOUND, hi = VISIBLE_LIGHT_WAVELENGTH_UPPER_BOUND;
        Spectrum s;
        switch (kind) {
            case 0: s = Spectrum::new_temperature_spectrum(lo, hi, arg0, n, arg1); break;
            case 1: s = Spectrum::new_singular_reflectance_factor(lo, hi, n, arg0); break;
            case 2: s = Spectrum::new_reflective_spectrum_red(lo, hi, n, arg0); break;
            case 3: s = Spectrum::new_reflective_spectrum_green(lo, hi, n, arg0); break;
            case 4: s = Spectrum::new_reflective_spectrum_blue(lo, hi, n, arg0); break;
            case 5: s = Spectrum::new_sunlight_spectrum(lo, hi, n, arg0); break;
            default: throw std::invalid_argument("unknown spectrum kind");
        }
        std::memcpy(out, s.intensities.data(), n * sizeof(float));
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}
// Spectrum::resample / get_radiance / normalize through the host mirror's methods (which run on the device):
// op 0 resample to n_new (out = n_new floats), 1 get_radiance (out[0]), 2 normalize (out = n floats)
int srth_spectrum_tool(uint32_t op, const float* in, uint32_t n, uint32_t n_new, float* out) {
    try {
        std::array<float, NBR_OF_SAMPLES_MAX> a{};
        std::memcpy(a.data(), in, n * sizeof(float));
        Spectrum s = Spectrum::new_from_list(a, VISIBLE_LIGHT_WAVELENGTH_LOWER_BOUND, VISIBLE_LIGHT_WAVELENGTH_UPPER_BOUND, n);
        if (op == 0) {
            s.resample(n_new);
            std::memcpy(out, s.intensities.data(), n_new * sizeof(float));
        } else if (op == 1) {
            out[0] = s.get_radiance();
        } else if (op == 2) {
            Spectrum t = s.normalize();
            std::memcpy(out, t.intensities.data(), n * sizeof(float));
        } else {
            throw std::invalid_argument("unknown spectrum tool");
        }
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}
double srth_black_body(double wavelength_nm, double temperature_k) {
    try {
        return black_body_radiation(wavelength_nm, temperature_k);
    } catch (const std::exception& e) {
        g_err = e.what();
        return std::nan("");
    }
}
void srth_to_rgba8(const float* data, size_t n, uint8_t* out) {
    CustomImage img;
    img.data.assign(data, data + n);
    std::vector<uint8_t> v = img.to_rgba8();
    std::memcpy(out, v.data(), n);
}

// CustomImage -> DynamicImage -> save(path) (main.rs:2325-2326) for an image given as CustomImage.data: the RGBA8
// conversion of custom_image.rs:92-101 written as a PNG.  Returns 0 on success.
int srth_save_png(const char* path, uint32_t width, uint32_t height, const float* data) {
    try {
        CustomImage img;
        img.width = width;
        img.height = height;
        img.data.assign(data, data + (size_t)width * height * 4);
        return write_png_rgba8(path, width, height, img.to_rgba8()) ? 0 : 1;
    } catch (...) {
        return 1;
    }
}

// dispatch_render_headless for a preset: the call a user of the reference makes when they press
// "start render", minus the GUI.  image = width*height*4 f32 (CustomImage.data).
int srth_dispatch_render(const char* preset, uint32_t arg, uint32_t width, uint32_t height, uint32_t n_lambda,
                         uint32_t iterations, uint32_t bounces, uint32_t rng_mode, uint32_t math_mode, int32_t device,
                         uint32_t first_frame, uint32_t n_frames, float* image, double* device_seconds,
                         uint64_t* kernel_launches, srt_counters* counters) {
    try {
        UIFields ui = make_preset(preset, n_lambda, arg);
        ui.width = width;
        ui.height = height;
        ui.nbr_of_iterations = iterations;
        ui.nbr_of_ray_bounces = bounces;
        RenderOptions opt;
        opt.rng_mode = rng_mode;
        opt.math_mode = math_mode;
        opt.device = device;
        opt.first_frame = first_frame;
        opt.n_frames = n_frames;
        RenderResult r = dispatch_render_headless(ui, opt);
        if (image) std::memcpy(image, r.image.data.data(), r.image.data.size() * sizeof(float));
        if (device_seconds) *device_seconds = r.device_seconds;
        if (kernel_launches) *kernel_launches = r.kernel_launches;
        if (counters) *counters = r.counters;
        return SRT_OK;
    } catch (const SrtError& e) {
        g_err = e.what();
        return e.code;
    } catch (const std::exception& e) {
        g_err = e.what();
        return SRT_ERR_INVALID_ARGUMENT;
    }
}

// App::render's action protocol for a preset (srt_host::render): runs `iterations` frames with an update every
// frames_per_update frames; abort_at_update >= 0 sets the AbortRender message while that update's actions are
// being pushed.  kinds / progress receive the first max_actions entries of the action list (AppAction::Kind,
// progress value or seconds), last_frame (optional, width*height*4) the image of the last FrameUpdate.
// Returns the number of actions, or a negative srt status.
int srth_render_protocol(const char* preset, uint32_t arg, uint32_t width, uint32_t height, uint32_t n_lambda,
                         uint32_t iterations, uint32_t bounces, uint32_t frames_per_update, int32_t abort_at_update,
                         int32_t* kinds, float* progress, uint32_t max_actions, uint8_t* last_frame,
                         uint64_t* frames_accumulated, int32_t* completed) {
    try {
        UIFields ui = make_preset(preset, n_lambda, arg);
        ui.width = width;
        ui.height = height;
        ui.nbr_of_iterations = iterations;
        ui.nbr_of_ray_bounces = bounces;
        RaytracingUniforms uniforms = build_uniforms(ui);
        Context ctx(uniforms, width, height, RenderOptions());
        RenderChannel ch;
        // the "UI thread": watches the action list and sends AbortRender once update abort_at_update arrived
        std::atomic<bool> stop{false};
        std::thread ui_thread([&] {
            while (!stop) {
                if (abort_at_update >= 0) {
                    std::lock_guard<std::mutex> g(ch.lock);
                    int updates = 0;
                    for (const AppAction& a : ch.action_list) updates += a.kind == AppAction::RenderingProgressUpdate;
                    if (updates > abort_at_update) ch.abort_render = true;
                }
                std::this_thread::yield();
            }
        });
        bool done = false;
        try {
            done = render(ctx, iterations, ch, frames_per_update);
        } catch (...) {
            stop = true;
            ui_thread.join();
            throw;
        }
        stop = true;
        ui_thread.join();
        if (completed) *completed = done ? 1 : 0;
        if (frames_accumulated) *frames_accumulated = srt_frames_accumulated(ctx.h);
        uint32_t i = 0;
        const AppAction* last = nullptr;
        for (const AppAction& a : ch.action_list) {
            if (i < max_actions) {
                if (kinds) kinds[i] = (int32_t)a.kind;
                if (progress) progress[i] = a.kind == AppAction::TrueTimeUpdate ? (float)a.seconds : a.progress;
            }
            if (a.kind == AppAction::FrameUpdate) last = &a;
            ++i;
        }
        if (last_frame && last && !last->image.empty()) std::memcpy(last_frame, last->image.data(), last->image.size());
        return (int)ch.action_list.size();
    } catch (const SrtError& e) {
        g_err = e.what();
        return -e.code;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -SRT_ERR_INVALID_ARGUMENT;
    }
}

}  // extern "C"
