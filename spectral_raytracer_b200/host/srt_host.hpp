// srt_host.hpp -- host side of the B200 render backend: a C++ mirror of the
// reference's scene types, Spectrum constructors, scene presets and of
// App::dispatch_render, written above the C ABI of include/srt.h.
//
// The reference is a Rust crate and this image has no Rust toolchain, so the host
// layer that the north star wants in Rust ("keeps the reference's scene, Spectrum
// and material types and its shader-stage API, calls CUDA through a thin extern "C"
// FFI crate, headless entry beside main::App::dispatch_render") is written in C++
// with the same names, argument meaning and error behaviour; INTEGRATION.md shows
// the Rust binding a maintainer adds on a machine that has cargo.
//
// Nothing here renders: the per-pixel path is libsrt.so.  This header only builds
// the inputs (bit-faithfully -- the spectra and primitive bounds are computed with
// the reference's f32/f64 operation order) and drives srt_* calls.
//
// Citations are into /root/reference/src.
#pragma once
#include <array>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/srt.h"

namespace srt_host {

// ============================================================== spectrum.rs
constexpr float VISIBLE_LIGHT_WAVELENGTH_LOWER_BOUND = 380.0f;  // spectrum.rs:5
constexpr float VISIBLE_LIGHT_WAVELENGTH_UPPER_BOUND = 780.0f;  // spectrum.rs:6
constexpr size_t NBR_OF_SAMPLES_MAX = 128;                      // spectrum.rs:8

// black_body_radiation, spectrum.rs:582-594 (f64; panics on non-positive input).
inline double black_body_radiation(double wavelength_nm, double temperature_k) {
    if (!(wavelength_nm > 0.0)) throw std::invalid_argument("wavelength must be positive");
    if (!(temperature_k > 0.0)) throw std::invalid_argument("temperature must be positive");
    const double c = 299792458.0, h = 6.62607015e-34, kb = 1.380649e-23;
    const double lambda = wavelength_nm / 1e9;
    const double hc22 = 2.0 * h * c * c;
    const double l5 = lambda * lambda * lambda * lambda * lambda;
    const double big_denominator = std::exp((h * c) / (lambda * temperature_k * kb)) - 1.0;
    return (hc22 / l5) * (1.0 / big_denominator) * 1e-9;
}

// Spectrum, spectrum.rs:25-30: fixed [f32; 128] storage, the first nbr_of_samples
// entries are meaningful, SpectrumType::EquidistantSamples(lowest, highest).
class Spectrum {
public:
    size_t nbr_of_samples = 0;
    std::array<float, NBR_OF_SAMPLES_MAX> intensities{};
    float lowest_wavelength = VISIBLE_LIGHT_WAVELENGTH_LOWER_BOUND;
    float highest_wavelength = VISIBLE_LIGHT_WAVELENGTH_UPPER_BOUND;

    // Spectrum::new asserts (spectrum.rs:37-38)
    static void check_samples(size_t n) {
        if (n == 0 || n % 8 != 0 || n > NBR_OF_SAMPLES_MAX)
            throw std::invalid_argument("number of spectral samples must be a multiple of 8 in 8..=128");
    }
    static Spectrum new_from_list(const std::array<float, NBR_OF_SAMPLES_MAX>& list, float lo, float hi, size_t n) {  // :62-68
        Spectrum s;
        s.nbr_of_samples = n;
        s.intensities = list;
        s.lowest_wavelength = lo;
        s.highest_wavelength = hi;
        return s;
    }
    static Spectrum new_singular_reflectance_factor(float lo, float hi, size_t n, float factor) {  // :100-106
        check_samples(n);
        std::array<float, NBR_OF_SAMPLES_MAX> a;
        a.fill(factor);
        return new_from_list(a, lo, hi, n);
    }
    static Spectrum new_equal_size_empty_spectrum(const Spectrum& other) {  // :49-58
        return new_singular_reflectance_factor(other.lowest_wavelength, other.highest_wavelength, other.nbr_of_samples, 0.0f);
    }
    // :112-122 -- wavelength_i = lo + step * i in f32, black body in f64, cast, * factor
    static Spectrum new_temperature_spectrum(float lo, float hi, float temperature, size_t n, float factor) {
        check_samples(n);
        std::array<float, NBR_OF_SAMPLES_MAX> a{};
        const float step = (hi - lo) / (float)(n - 1);
        for (size_t i = 0; i < NBR_OF_SAMPLES_MAX; ++i) {
            const float wavelength = lo + step * (float)i;
            a[i] = (float)black_body_radiation((double)wavelength, (double)temperature) * factor;
        }
        return new_from_list(a, lo, hi, n);
    }
    // :73-96 -- the measured table is disabled in the reference; sunlight is a 6500 K black body
    static Spectrum new_sunlight_spectrum(float lo, float hi, size_t n, float factor) {
        return new_temperature_spectrum(lo, hi, 6500.0f, n, factor);
    }
    template <class Pred>
    static Spectrum step_spectrum(float lo, float hi, size_t n, float factor, Pred in_band) {
        check_samples(n);
        std::array<float, NBR_OF_SAMPLES_MAX> a{};
        const float step = (hi - lo) / (float)(n - 1);
        for (size_t i = 0; i < n; ++i)
            if (in_band(lo + step * (float)i)) a[i] = factor;
        return new_from_list(a, lo, hi, n);
    }
    static Spectrum new_reflective_spectrum_red(float lo, float hi, size_t n, float f) {  // :141-154
        return step_spectrum(lo, hi, n, f, [](float w) { return 550.0f < w; });
    }
    static Spectrum new_reflective_spectrum_green(float lo, float hi, size_t n, float f) {  // :158-171
        return step_spectrum(lo, hi, n, f, [](float w) { return 500.0f < w && w < 575.0f; });
    }
    static Spectrum new_reflective_spectrum_blue(float lo, float hi, size_t n, float f) {  // :175-187
        return step_spectrum(lo, hi, n, f, [](float w) { return w < 475.0f; });
    }
    void max0() {  // :215-221
        for (size_t i = 0; i < nbr_of_samples; ++i) intensities[i] = std::fmax(intensities[i], 0.0f);
    }
    void min1() {  // :224-230
        for (size_t i = 0; i < nbr_of_samples; ++i) intensities[i] = std::fmin(intensities[i], 1.0f);
    }
    size_t get_nbr_of_samples() const { return nbr_of_samples; }

    // The spectrum editor's tooling (spectrum.rs:285-374) runs on the device (srt_spectra_*, include/srt.h):
    // resample -- a Custom spectrum follows a change of the sample count (main.rs:1193-1195); throws where the
    // reference panics (reductions beyond what its down-sampling loop survives)
    void resample(size_t new_sample_amount) {  // :285-323
        check_samples(new_sample_amount);
        if (new_sample_amount == nbr_of_samples) return;
        std::array<float, NBR_OF_SAMPLES_MAX> out{};
        const int rc = srt_spectra_resample(intensities.data(), 1, (uint32_t)nbr_of_samples, (uint32_t)new_sample_amount, out.data());
        if (rc != SRT_OK) throw std::runtime_error(srt_last_error(nullptr));
        intensities = out;
        nbr_of_samples = new_sample_amount;
    }
    float get_radiance() const {  // :357-362
        float r = 0.0f;
        if (srt_spectra_radiance(intensities.data(), 1, (uint32_t)nbr_of_samples, lowest_wavelength, highest_wavelength, &r) != SRT_OK)
            throw std::runtime_error(srt_last_error(nullptr));
        return r;
    }
    Spectrum normalize() const {  // :369-374
        Spectrum s = *this;
        s.intensities.fill(0.0f);
        if (srt_spectra_normalize(intensities.data(), 1, (uint32_t)nbr_of_samples, lowest_wavelength, highest_wavelength,
                                  s.intensities.data()) != SRT_OK)
            throw std::runtime_error(srt_last_error(nullptr));
        return s;
    }
};

// ============================================================== nalgebra bits
struct Vector3 {
    float x = 0, y = 0, z = 0;
};
using Point3 = Vector3;
struct Rotation3 {
    float m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};  // row-major
    // Rotation3::from_euler_angles(roll, pitch, yaw) = Rz(yaw) * Ry(pitch) * Rx(roll)  (nalgebra 0.33.2)
    static Rotation3 from_euler_angles(float roll, float pitch, float yaw) {
        const float sr = std::sin(roll), cr = std::cos(roll), sp = std::sin(pitch), cp = std::cos(pitch);
        const float sy = std::sin(yaw), cy = std::cos(yaw);
        Rotation3 r;
        r.m[0] = cy * cp; r.m[1] = cy * sp * sr - sy * cr; r.m[2] = cy * sp * cr + sy * sr;
        r.m[3] = sy * cp; r.m[4] = sy * sp * sr + cy * cr; r.m[5] = sy * sp * cr - cy * sr;
        r.m[6] = -sp;     r.m[7] = cp * sr;                r.m[8] = cp * cr;
        return r;
    }
    Vector3 operator*(const Vector3& v) const {
        return {(m[0] * v.x + m[1] * v.y) + m[2] * v.z, (m[3] * v.x + m[4] * v.y) + m[5] * v.z,
                (m[6] * v.x + m[7] * v.y) + m[8] * v.z};
    }
};
inline Vector3 cross(const Vector3& a, const Vector3& b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// ============================================================== shader.rs types
constexpr float F32_DELTA = 0.00001f;  // shader.rs:7

struct Material {  // shader.rs:253-258 (+ the dispersion extension, off by default)
    Spectrum reflective_spectrum;
    float metallicness = 0.0f;
    float roughness = 0.0f;
    bool transmissive = false;
    float ior_a = 1.0f, ior_b = 0.0f;
};

enum class AABBType { PlainBox, Sphere, RotatedBox };  // shader.rs:168-172

struct Aabb {  // shader.rs:99-104
    Point3 min, max;
    AABBType aabb_type = AABBType::PlainBox;
    Point3 rb_center;      // RotatedBox(center, dims, rotation)
    Vector3 rb_dims;
    Rotation3 rb_rotation;
    std::shared_ptr<const Material> material;

    static Aabb new_sphere(const Point3& c, float radius, std::shared_ptr<const Material> m) {  // :108-115
        Aabb a;
        a.min = {c.x - radius, c.y - radius, c.z - radius};
        a.max = {c.x + radius, c.y + radius, c.z + radius};
        a.aabb_type = AABBType::Sphere;
        a.material = std::move(m);
        return a;
    }
    static Aabb new_box(const Point3& c, float xl, float yl, float zl, std::shared_ptr<const Material> m) {  // :120-130
        const float xh = xl / 2.0f, yh = yl / 2.0f, zh = zl / 2.0f;
        Aabb a;
        a.min = {c.x - xh, c.y - yh, c.z - zh};
        a.max = {c.x + xh, c.y + yh, c.z + zh};
        a.aabb_type = AABBType::PlainBox;
        a.material = std::move(m);
        return a;
    }
    // :134-166 -- bounds = component-wise min / max of the eight rotated corners
    static Aabb new_rotated_box(const Point3& c, float xl, float yl, float zl, const Rotation3& rot,
                                std::shared_ptr<const Material> m) {
        const float xh = xl / 2.0f, yh = yl / 2.0f, zh = zl / 2.0f;
        Aabb a;
        bool first = true;
        for (float sx : {-xh, xh})
            for (float sy : {-yh, yh})
                for (float sz : {-zh, zh}) {
                    const Vector3 r = rot * Vector3{sx, sy, sz};
                    const Point3 p{c.x + r.x, c.y + r.y, c.z + r.z};
                    if (first) {
                        a.min = a.max = p;
                        first = false;
                    } else {
                        a.min = {std::fmin(a.min.x, p.x), std::fmin(a.min.y, p.y), std::fmin(a.min.z, p.z)};
                        a.max = {std::fmax(a.max.x, p.x), std::fmax(a.max.y, p.y), std::fmax(a.max.z, p.z)};
                    }
                }
        a.aabb_type = AABBType::RotatedBox;
        a.rb_center = c;
        a.rb_dims = {xl, yl, zl};
        a.rb_rotation = rot;
        a.material = std::move(m);
        return a;
    }
};

struct Light {  // shader.rs:192-195
    Point3 position;
    Spectrum spectrum;
};

struct Camera {  // shader.rs:213-218
    Point3 position{0.0f, 0.0f, -2.0f};
    Vector3 direction{0.0f, 0.0f, 1.0f};
    Vector3 up{0.0f, 1.0f, 0.0f};
    float fov_y_deg = 60.0f;
};

struct RaytracingUniforms {  // shader.rs:32-41
    std::vector<Aabb> aabbs;
    std::vector<Light> lights;
    Camera camera;
    uint32_t frame_id = 0;
    uint32_t intended_frames_amount = 1;
    Spectrum example_spectrum;
    uint32_t max_bounces = 30;
};

// are_linear_dependent, main.rs:2200-2203
inline bool are_linear_dependent(const Vector3& a, const Vector3& b) {
    const Vector3 c = cross(a, b);
    return std::fabs(c.x) < F32_DELTA && std::fabs(c.y) < F32_DELTA && std::fabs(c.z) < F32_DELTA;
}

// ============================================================== UI model (main.rs:1511-1535, :1761-2196)
enum class SpectrumEffectType { Emissive, Reflective };
enum class UISpectrumKind { Custom, Solar, PlainReflective, Temperature, ReflectiveRed, ReflectiveGreen, ReflectiveBlue };

struct UISpectrum {
    std::string name;
    UISpectrumKind kind = UISpectrumKind::Custom;
    float arg0 = 1.0f, arg1 = 1.0f;  // Solar(factor) / PlainReflective(factor) / Temperature(temp, factor) / Reflective*(factor)
    SpectrumEffectType effect = SpectrumEffectType::Emissive;
    Spectrum spectrum;

    // update_all_spectrum_sample_sizes, main.rs:1186-1228: Custom spectra are resampled (on the device),
    // every other kind is rebuilt from its parameters
    void regenerate(float lo, float hi, size_t n) {
        switch (kind) {
            case UISpectrumKind::Custom: spectrum.resample(n); break;
            case UISpectrumKind::Solar: spectrum = Spectrum::new_sunlight_spectrum(lo, hi, n, arg0); break;
            case UISpectrumKind::PlainReflective: spectrum = Spectrum::new_singular_reflectance_factor(lo, hi, n, arg0); break;
            case UISpectrumKind::Temperature: spectrum = Spectrum::new_temperature_spectrum(lo, hi, arg0, n, arg1); break;
            case UISpectrumKind::ReflectiveRed: spectrum = Spectrum::new_reflective_spectrum_red(lo, hi, n, arg0); break;
            case UISpectrumKind::ReflectiveGreen: spectrum = Spectrum::new_reflective_spectrum_green(lo, hi, n, arg0); break;
            case UISpectrumKind::ReflectiveBlue: spectrum = Spectrum::new_reflective_spectrum_blue(lo, hi, n, arg0); break;
        }
    }
    // From<&UISpectrum> for Spectrum, spectrum.rs:486-494: reflective spectra are clamped to <= 1
    Spectrum to_render_spectrum() const {
        Spectrum s = spectrum;
        if (effect == SpectrumEffectType::Reflective) s.min1();
        return s;
    }
};
using UISpectrumRef = std::shared_ptr<UISpectrum>;

struct UIMaterial {
    std::string name;
    float metallicness = 0.0f, roughness = 0.0f;
    UISpectrumRef spectrum;
    bool transmissive = false;  // extension
    float ior_a = 1.0f, ior_b = 0.0f;
};
using UIMaterialRef = std::shared_ptr<UIMaterial>;

struct UILight {
    float pos_x = 0, pos_y = 0, pos_z = 0;
    UISpectrumRef spectrum;
    std::string name;
    bool hidden = false;
};

enum class UIObjectKind { PlainBox, Sphere, RotatedBox };
struct UIObject {
    float pos_x = 0, pos_y = 0, pos_z = 0;
    UIMaterialRef material;
    UIObjectKind kind = UIObjectKind::PlainBox;
    float p[6] = {0, 0, 0, 0, 0, 0};  // PlainBox(x,y,z) / Sphere(r) / RotatedBox(x,y,z, rot_x,rot_y,rot_z)
    std::string name;
    bool hidden = false;
};

struct UIFields {
    uint32_t width = 600, height = 400;                 // main.rs:1734-1735
    uint32_t nbr_of_iterations = 100;                   // NBR_OF_ITERATIONS_DEFAULT, main.rs:31
    uint32_t nbr_of_ray_bounces = 30;                   // NEW_RAY_MAX_BOUNCES_DEFAULT, main.rs:33
    size_t spectrum_number_of_samples = 32;             // NBR_OF_SPECTRUM_SAMPLES_DEFAULT, main.rs:32
    float spectrum_lower_bound = VISIBLE_LIGHT_WAVELENGTH_LOWER_BOUND;
    float spectrum_upper_bound = VISIBLE_LIGHT_WAVELENGTH_UPPER_BOUND;
    Camera ui_camera;                                   // UICamera::default(), main.rs:1970-1985
    std::vector<UILight> ui_lights;
    std::vector<UIObject> ui_objects;
    std::vector<UISpectrumRef> spectra;
    std::vector<UIMaterialRef> materials;

    UISpectrumRef add_spectrum(std::string name, UISpectrumKind kind, SpectrumEffectType effect, float arg0, float arg1 = 1.0f) {
        auto s = std::make_shared<UISpectrum>();
        s->name = std::move(name);
        s->kind = kind;
        s->effect = effect;
        s->arg0 = arg0;
        s->arg1 = arg1;
        s->regenerate(spectrum_lower_bound, spectrum_upper_bound, spectrum_number_of_samples);
        spectra.push_back(s);
        return s;
    }
    UIMaterialRef add_material(std::string name, float metallicness, float roughness, UISpectrumRef spectrum) {
        auto m = std::make_shared<UIMaterial>();
        m->name = std::move(name);
        m->metallicness = metallicness;
        m->roughness = roughness;
        m->spectrum = std::move(spectrum);
        materials.push_back(m);
        return m;
    }
    void add_light(float x, float y, float z, UISpectrumRef s, std::string name) {
        ui_lights.push_back(UILight{x, y, z, std::move(s), std::move(name), false});
    }
    void add_object(float x, float y, float z, UIMaterialRef m, UIObjectKind kind, std::initializer_list<float> params,
                    std::string name) {
        UIObject o;
        o.pos_x = x; o.pos_y = y; o.pos_z = z;
        o.material = std::move(m);
        o.kind = kind;
        size_t i = 0;
        for (float v : params) o.p[i++] = v;
        o.name = std::move(name);
        ui_objects.push_back(std::move(o));
    }

    // ---- presets
    // UIFields::default(), main.rs:1638-1758 (the README example scene)
    static UIFields default_scene(size_t n_samples = 32) {
        UIFields f;
        f.spectrum_number_of_samples = n_samples;
        auto sun10 = f.add_spectrum("Close light spectrum", UISpectrumKind::Solar, SpectrumEffectType::Emissive, 0.001f);
        auto sun1mil = f.add_spectrum("Far away sun spectrum", UISpectrumKind::Solar, SpectrumEffectType::Emissive, 100.0f);
        auto grey = f.add_spectrum("Grey reflecting spectrum", UISpectrumKind::PlainReflective, SpectrumEffectType::Reflective, 0.7f);
        auto white = f.add_spectrum("White reflecting spectrum", UISpectrumKind::PlainReflective, SpectrumEffectType::Reflective, 1.0f);
        f.add_light(0.0f, 2.0f, -1.0f, sun10, "Close light");
        f.add_light(0.0f, 1000.0f, 0.0f, sun1mil, "Far away sun light");
        auto mirror = f.add_material("Perfect Mirror", 1.0f, 0.2f, white);
        auto plastic = f.add_material("Grey plastic", 0.0f, 0.0f, grey);
        f.add_object(-1.5f, 0.0f, 1.0f, mirror, UIObjectKind::PlainBox, {0.25f, 3.0f, 30.0f}, "Left mirror");
        f.add_object(0.0f, 0.0f, 1.0f, plastic, UIObjectKind::Sphere, {1.0f}, "Left sphere");
        f.add_object(1.0f, 0.0f, 1.0f, plastic, UIObjectKind::Sphere, {1.0f}, "Right sphere");
        f.add_object(0.0f, -1.0f, 0.0f, plastic, UIObjectKind::PlainBox, {50.0f, 0.1f, 50.0f}, "Floor");
        return f;
    }
    // UIFields::cornell_box(), main.rs:1538-1635
    static UIFields cornell_box(size_t n_samples = 32) {
        UIFields f;
        f.spectrum_number_of_samples = n_samples;
        auto sun = f.add_spectrum("Solar light spectrum", UISpectrumKind::Solar, SpectrumEffectType::Emissive, 0.0001f);
        auto grey = f.add_spectrum("Reflective gray", UISpectrumKind::PlainReflective, SpectrumEffectType::Reflective, 0.7f);
        auto red = f.add_spectrum("Reflective red", UISpectrumKind::ReflectiveRed, SpectrumEffectType::Reflective, 1.0f);
        auto green = f.add_spectrum("Reflective green", UISpectrumKind::ReflectiveGreen, SpectrumEffectType::Reflective, 1.0f);
        f.add_light(0.0f, 0.9f, 0.0f, sun, "Top light");
        auto m_grey = f.add_material("Grey plastic", 0.0f, 0.0f, grey);
        auto m_green = f.add_material("Green plastic", 0.0f, 0.0f, green);
        auto m_red = f.add_material("Red plastic", 0.0f, 0.0f, red);
        f.add_object(0.0f, 0.0f, 2.0f, m_grey, UIObjectKind::PlainBox, {2.0f, 2.0f, 2.0f}, "Central wall");
        f.add_object(0.0f, 2.0f, 0.0f, m_grey, UIObjectKind::PlainBox, {2.0f, 2.0f, 2.0f}, "Ceiling");
        f.add_object(0.0f, -2.0f, 0.0f, m_grey, UIObjectKind::PlainBox, {2.0f, 2.0f, 2.0f}, "Floor");
        f.add_object(-2.0f, 0.0f, 0.0f, m_red, UIObjectKind::PlainBox, {2.0f, 2.0f, 2.0f}, "Left wall");
        f.add_object(2.0f, 0.0f, 0.0f, m_green, UIObjectKind::PlainBox, {2.0f, 2.0f, 2.0f}, "Right wall");
        f.add_object(0.5f, -0.75f, -0.5f, m_grey, UIObjectKind::RotatedBox, {0.5f, 0.5f, 0.5f, 0.0f, 1.0f, 0.0f}, "Right front box");
        f.add_object(-0.5f, -0.4f, 0.5f, m_grey, UIObjectKind::RotatedBox, {0.5f, 1.2f, 0.5f, 0.0f, -0.5f, 0.0f}, "Left back box");
        return f;
    }
    // BASELINE.json config 5 / SURVEY.md 8(d) C4: floor + n spheres placed with the reference's own
    // random_pcg3d hash (shader.rs:685-705), the two lights of the default scene.
    static UIFields random_spheres(uint32_t n_spheres, size_t n_samples = 32);
    // BASELINE.json config 3 (beyond-reference extension): Cornell walls + one dispersive glass sphere.
    static UIFields prism(size_t n_samples = 32);
};

// random_pcg3d, shader.rs:685-705 -- host copy, used only to place the spheres of config C4.
inline void random_pcg3d(uint32_t x, uint32_t y, uint32_t z, float out[3]) {
    x = x * 1664525u + 1013904223u;
    y = y * 1664525u + 1013904223u;
    z = z * 1664525u + 1013904223u;
    x += y * z; y += z * x; z += x * y;
    x ^= x >> 16; y ^= y >> 16; z ^= z >> 16;
    x += y * z; y += z * x; z += x * y;
    const float reciprocal = 1.0f / (float)0xffffffffu;
    out[0] = (float)x * reciprocal;
    out[1] = (float)y * reciprocal;
    out[2] = (float)z * reciprocal;
}

inline UIFields UIFields::random_spheres(uint32_t n_spheres, size_t n_samples) {
    UIFields f;
    f.spectrum_number_of_samples = n_samples;
    auto sun10 = f.add_spectrum("Close light spectrum", UISpectrumKind::Solar, SpectrumEffectType::Emissive, 0.001f);
    auto sun1mil = f.add_spectrum("Far away sun spectrum", UISpectrumKind::Solar, SpectrumEffectType::Emissive, 100.0f);
    auto grey = f.add_spectrum("grey", UISpectrumKind::PlainReflective, SpectrumEffectType::Reflective, 0.7f);
    auto white = f.add_spectrum("white", UISpectrumKind::PlainReflective, SpectrumEffectType::Reflective, 1.0f);
    auto red = f.add_spectrum("red", UISpectrumKind::ReflectiveRed, SpectrumEffectType::Reflective, 1.0f);
    auto green = f.add_spectrum("green", UISpectrumKind::ReflectiveGreen, SpectrumEffectType::Reflective, 1.0f);
    auto blue = f.add_spectrum("blue", UISpectrumKind::ReflectiveBlue, SpectrumEffectType::Reflective, 1.0f);
    f.add_light(0.0f, 2.0f, -1.0f, sun10, "Close light");
    f.add_light(0.0f, 1000.0f, 0.0f, sun1mil, "Far away sun light");
    UIMaterialRef mats[8] = {
        f.add_material("grey", 0.0f, 0.0f, grey),      f.add_material("red", 0.0f, 0.0f, red),
        f.add_material("green", 0.0f, 0.0f, green),    f.add_material("blue", 0.0f, 0.0f, blue),
        f.add_material("mirror", 1.0f, 0.0f, white),   f.add_material("metal 0.1", 1.0f, 0.1f, white),
        f.add_material("metal 0.2", 1.0f, 0.2f, white), f.add_material("metal 0.4", 1.0f, 0.4f, white)};
    f.add_object(0.0f, -1.0f, 0.0f, mats[0], UIObjectKind::PlainBox, {50.0f, 0.1f, 50.0f}, "Floor");
    for (uint32_t i = 0; i < n_spheres; ++i) {
        float a[3], b[3];
        random_pcg3d(i, 0x5EEDu, 1u, a);
        random_pcg3d(i, 0x5EEDu, 2u, b);
        const float r = 0.03f + 0.09f * a[2];
        uint32_t mi = (uint32_t)(8.0f * b[0]);
        if (mi > 7) mi = 7;
        f.add_object(-8.0f + 16.0f * a[0], -0.9f + r, 0.0f + 16.0f * a[1], mats[mi], UIObjectKind::Sphere, {r}, "sphere");
    }
    return f;
}

inline UIFields UIFields::prism(size_t n_samples) {
    UIFields f = cornell_box(n_samples);
    auto white = f.add_spectrum("Glass", UISpectrumKind::PlainReflective, SpectrumEffectType::Reflective, 1.0f);
    auto glass = f.add_material("Dispersive glass", 0.0f, 0.0f, white);
    glass->transmissive = true;
    glass->ior_a = 1.30f;
    glass->ior_b = 6000.0f;  // Cauchy n = A + B / lambda_nm^2  
    f.add_object(0.0f, -0.2f, -0.2f, glass, UIObjectKind::Sphere, {0.35f}, "Glass sphere");
    return f;
}

// ============================================================== flattening
// RaytracingUniforms -> the POD arrays of the C ABI.
struct FlatScene {
    std::vector<srt_object> objects;
    std::vector<srt_material> materials;
    std::vector<srt_light> lights;
    std::vector<float> spectra;  // n_spectra * n_lambda
    uint32_t n_lambda = 0, n_spectra = 0;
    srt_camera camera{};
    float lambda_min = VISIBLE_LIGHT_WAVELENGTH_LOWER_BOUND, lambda_max = VISIBLE_LIGHT_WAVELENGTH_UPPER_BOUND;

    uint32_t add_spectrum(const Spectrum& s) {
        spectra.insert(spectra.end(), s.intensities.begin(), s.intensities.begin() + n_lambda);
        return n_spectra++;
    }
};

inline FlatScene flatten(const RaytracingUniforms& u) {
    FlatScene f;
    f.n_lambda = (uint32_t)u.example_spectrum.nbr_of_samples;
    f.lambda_min = u.example_spectrum.lowest_wavelength;
    f.lambda_max = u.example_spectrum.highest_wavelength;
    std::vector<const Material*> seen;
    for (const Aabb& a : u.aabbs) {
        if (a.material->reflective_spectrum.nbr_of_samples != f.n_lambda)
            throw std::invalid_argument("material spectrum has the wrong sample count");  // check_render_legality, main.rs:1452-1464
        uint32_t mi = 0;
        for (; mi < seen.size(); ++mi)
            if (seen[mi] == a.material.get()) break;
        if (mi == seen.size()) {
            seen.push_back(a.material.get());
            srt_material m{};
            m.metallicness = a.material->metallicness;
            m.roughness = a.material->roughness;
            m.reflectance = f.add_spectrum(a.material->reflective_spectrum);
            m.transmissive = a.material->transmissive ? 1u : 0u;
            m.ior_a = a.material->ior_a;
            m.ior_b = a.material->ior_b;
            f.materials.push_back(m);
        }
        srt_object o{};
        o.min[0] = a.min.x; o.min[1] = a.min.y; o.min[2] = a.min.z;
        o.max[0] = a.max.x; o.max[1] = a.max.y; o.max[2] = a.max.z;
        o.kind = a.aabb_type == AABBType::PlainBox ? SRT_PLAIN_BOX : (a.aabb_type == AABBType::Sphere ? SRT_SPHERE : SRT_ROTATED_BOX);
        o.center[0] = a.rb_center.x; o.center[1] = a.rb_center.y; o.center[2] = a.rb_center.z;
        o.dims[0] = a.rb_dims.x; o.dims[1] = a.rb_dims.y; o.dims[2] = a.rb_dims.z;
        std::memcpy(o.rot, a.rb_rotation.m, sizeof(o.rot));
        o.material = mi;
        f.objects.push_back(o);
    }
    for (const Light& l : u.lights) {
        if (l.spectrum.nbr_of_samples != f.n_lambda) throw std::invalid_argument("light spectrum has the wrong sample count");
        srt_light sl{};
        sl.position[0] = l.position.x; sl.position[1] = l.position.y; sl.position[2] = l.position.z;
        sl.spectrum = f.add_spectrum(l.spectrum);
        f.lights.push_back(sl);
    }
    const Camera& c = u.camera;
    f.camera = srt_camera{{c.position.x, c.position.y, c.position.z}, {c.direction.x, c.direction.y, c.direction.z},
                          {c.up.x, c.up.y, c.up.z}, c.fov_y_deg};
    return f;
}

// The uniform assembly of App::dispatch_render, main.rs:1377-1404: regenerate every spectrum at the
// current sample count, drop hidden objects / lights, convert UI types to render types
// (From<&UIObject> shader.rs:174-190, From<&UILight> :205-210, From<&UIMaterial> :260-268).
inline RaytracingUniforms build_uniforms(UIFields& ui) {
    Spectrum::check_samples(ui.spectrum_number_of_samples);
    for (auto& s : ui.spectra) s->regenerate(ui.spectrum_lower_bound, ui.spectrum_upper_bound, ui.spectrum_number_of_samples);
    RaytracingUniforms u;
    u.example_spectrum = Spectrum::new_singular_reflectance_factor(ui.spectrum_lower_bound, ui.spectrum_upper_bound,
                                                                   ui.spectrum_number_of_samples, 0.0f);
    std::vector<std::pair<const UIMaterial*, std::shared_ptr<const Material>>> converted;
    auto convert = [&](const UIMaterialRef& um) {
        for (auto& kv : converted)
            if (kv.first == um.get()) return kv.second;
        auto m = std::make_shared<Material>();
        m->reflective_spectrum = um->spectrum->to_render_spectrum();
        m->metallicness = um->metallicness;
        m->roughness = um->roughness;
        m->transmissive = um->transmissive;
        m->ior_a = um->ior_a;
        m->ior_b = um->ior_b;
        converted.emplace_back(um.get(), m);
        return std::shared_ptr<const Material>(m);
    };
    for (const UIObject& o : ui.ui_objects) {
        if (o.hidden) continue;
        const Point3 pos{o.pos_x, o.pos_y, o.pos_z};
        auto m = convert(o.material);
        switch (o.kind) {
            case UIObjectKind::PlainBox: u.aabbs.push_back(Aabb::new_box(pos, o.p[0], o.p[1], o.p[2], m)); break;
            case UIObjectKind::Sphere: u.aabbs.push_back(Aabb::new_sphere(pos, o.p[0], m)); break;
            case UIObjectKind::RotatedBox:
                u.aabbs.push_back(Aabb::new_rotated_box(pos, o.p[0], o.p[1], o.p[2],
                                                        Rotation3::from_euler_angles(o.p[3], o.p[4], o.p[5]), m));
                break;
        }
    }
    for (const UILight& l : ui.ui_lights) {
        if (l.hidden) continue;
        u.lights.push_back(Light{{l.pos_x, l.pos_y, l.pos_z}, l.spectrum->spectrum});  // raw, not clamped (shader.rs:207-208)
    }
    u.camera = ui.ui_camera;
    u.frame_id = 0;
    u.intended_frames_amount = ui.nbr_of_iterations;
    u.max_bounces = ui.nbr_of_ray_bounces;
    return u;
}

// ============================================================== custom_image.rs
// A PNG writer without dependencies: 8-bit RGBA, filter 0, zlib stream of STORED deflate blocks (no compression --
// every PNG reader accepts it).  CRC-32 over chunk type + data, Adler-32 over the raw scanlines (RFC 1950 / 1951 / 2083).
inline uint32_t crc32_update(uint32_t crc, const uint8_t* p, size_t n) {
    static uint32_t table[256];
    static bool ready = false;
    if (!ready) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        ready = true;
    }
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xffu] ^ (crc >> 8);
    return crc;
}
inline void put_be32(std::vector<uint8_t>& v, uint32_t x) {
    for (int s = 24; s >= 0; s -= 8) v.push_back((uint8_t)(x >> s));
}
inline bool write_chunk(FILE* f, const char type[4], const std::vector<uint8_t>& data) {
    std::vector<uint8_t> head;
    put_be32(head, (uint32_t)data.size());
    head.insert(head.end(), type, type + 4);
    uint32_t crc = crc32_update(0xffffffffu, reinterpret_cast<const uint8_t*>(type), 4);
    crc = crc32_update(crc, data.data(), data.size()) ^ 0xffffffffu;
    std::vector<uint8_t> tail;
    put_be32(tail, crc);
    return std::fwrite(head.data(), 1, head.size(), f) == head.size() &&
           (data.empty() || std::fwrite(data.data(), 1, data.size(), f) == data.size()) &&
           std::fwrite(tail.data(), 1, 4, f) == 4;
}
inline bool write_png_rgba8(const std::string& path, uint32_t w, uint32_t h, const std::vector<uint8_t>& rgba) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    bool ok = std::fwrite(sig, 1, 8, f) == 8;
    std::vector<uint8_t> ihdr;
    put_be32(ihdr, w);
    put_be32(ihdr, h);
    const uint8_t rest[5] = {8, 6, 0, 0, 0};  // bit depth 8, colour type 6 (RGBA), deflate, adaptive filtering, no interlace
    ihdr.insert(ihdr.end(), rest, rest + 5);
    ok = ok && write_chunk(f, "IHDR", ihdr);
    // raw data: every scanline preceded by its filter byte (0 = none)
    std::vector<uint8_t> raw;
    raw.reserve((size_t)h * (4 * (size_t)w + 1));
    for (uint32_t y = 0; y < h; ++y) {
        raw.push_back(0);
        raw.insert(raw.end(), rgba.begin() + (size_t)y * w * 4, rgba.begin() + (size_t)(y + 1) * w * 4);
    }
    uint32_t a = 1, b = 0;  // Adler-32
    for (size_t i = 0; i < raw.size();) {
        const size_t n = std::min<size_t>(5552, raw.size() - i);
        for (size_t k = 0; k < n; ++k) {
            a += raw[i + k];
            b += a;
        }
        a %= 65521u;
        b %= 65521u;
        i += n;
    }
    std::vector<uint8_t> z;
    z.reserve(raw.size() + raw.size() / 65535 * 5 + 16);
    z.push_back(0x78);
    z.push_back(0x01);
    for (size_t i = 0; i < raw.size() || i == 0;) {
        const size_t n = std::min<size_t>(65535, raw.size() - i);
        const bool last = i + n >= raw.size();
        z.push_back(last ? 1 : 0);
        z.push_back((uint8_t)(n & 0xff));
        z.push_back((uint8_t)(n >> 8));
        z.push_back((uint8_t)(~n & 0xff));
        z.push_back((uint8_t)((~n >> 8) & 0xff));
        z.insert(z.end(), raw.begin() + i, raw.begin() + i + n);
        i += n;
        if (last) break;
    }
    put_be32(z, (b << 16) | a);
    ok = ok && write_chunk(f, "IDAT", z) && write_chunk(f, "IEND", {});
    return (std::fclose(f) == 0) && ok;
}

struct CustomImage {  // custom_image.rs:9-22: RGBA f32, row-major, top-left origin
    uint32_t width = 0, height = 0;
    std::vector<float> data;
    CustomImage() = default;
    CustomImage(uint32_t w, uint32_t h) : width(w), height(h), data((size_t)w * h * 4, 0.0f) {}
    // From<CustomImage> for DynamicImage, custom_image.rs:92-101
    std::vector<uint8_t> to_rgba8() const {
        std::vector<uint8_t> out(data.size());
        for (size_t i = 0; i < data.size(); ++i) {
            float f = data[i];
            if (f != f) { out[i] = 0; continue; }
            f = f < 0.0f ? 0.0f : (f > 1.0f ? 1.0f : f);
            out[i] = (uint8_t)(f * 255.0f);
        }
        return out;
    }
};

// ============================================================== headless render
struct RenderOptions {
    uint32_t rng_mode = SRT_RNG_PCG3D_REFERENCE;
    uint32_t math_mode = SRT_MATH_FAST;
    uint32_t accel = SRT_ACCEL_AUTO;
    uint32_t integrator = SRT_INTEGRATOR_AUTO;
    int32_t device = -1;
    uint32_t pool_paths = 0;
    uint32_t frames_per_batch = 16;  // progress / abort granularity (the reference polls once per frame, main.rs:1351)
    uint32_t first_frame = 0;        // frame-sharded multi-GPU: this context renders [first_frame, first_frame + n_frames)
    uint32_t n_frames = 0;           // 0 = all of nbr_of_iterations
    bool (*progress)(float fraction, void* user) = nullptr;  // return false to abort (AppToRenderMessages::AbortRender)
    void* user = nullptr;
};

struct RenderResult {
    CustomImage image;
    double device_seconds = 0.0;
    uint64_t kernel_launches = 0;
    srt_counters counters{};
    bool aborted = false;
};

class SrtError : public std::runtime_error {
public:
    int code;
    SrtError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// RAII srt_ctx built from the uniforms, i.e. dispatch_render's validation + upload.
class Context {
public:
    srt_ctx* h = nullptr;
    uint32_t width, height;
    Context(const RaytracingUniforms& u, uint32_t w, uint32_t hgt, const RenderOptions& opt) : width(w), height(hgt) {
        // dispatch_render asserts this (main.rs:1407-1412); srt_create reports it as a status
        FlatScene f = flatten(u);
        srt_params p{};
        p.width = w;
        p.height = hgt;
        p.n_lambda = f.n_lambda;
        p.lambda_min = f.lambda_min;
        p.lambda_max = f.lambda_max;
        p.max_bounces = u.max_bounces;
        p.intended_frames = u.intended_frames_amount;
        p.rng_mode = opt.rng_mode;
        p.math_mode = opt.math_mode;
        p.accel = opt.accel;
        p.integrator = opt.integrator;
        p.device = opt.device;
        p.pool_paths = opt.pool_paths;
        int rc = srt_create(&p, &f.camera, f.objects.data(), (uint32_t)f.objects.size(), f.materials.data(),
                            (uint32_t)f.materials.size(), f.lights.data(), (uint32_t)f.lights.size(), f.spectra.data(),
                            f.n_spectra, &h);
        if (rc != SRT_OK) throw SrtError(rc, srt_last_error(nullptr));
    }
    ~Context() { srt_destroy(h); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    void check(int rc) const {
        if (rc != SRT_OK) throw SrtError(rc, srt_last_error(h));
    }
};

// Headless sibling of App::dispatch_render + App::render (main.rs:1376-1427, :1327-1371): builds the
// uniforms exactly like the GUI path, then renders on the GPU in batches of frames instead of
// spawning the CPU render thread, and returns the CustomImage the Display tab would show.
inline RenderResult dispatch_render_headless(UIFields& ui, const RenderOptions& opt = RenderOptions()) {
    RaytracingUniforms uniforms = build_uniforms(ui);
    if (are_linear_dependent(uniforms.camera.direction, uniforms.camera.up))
        throw SrtError(SRT_ERR_CAMERA_COLLINEAR, "View Direction and Up Direction are linearly dependent!");
    Context ctx(uniforms, ui.width, ui.height, opt);
    RenderResult res;
    const uint32_t total = opt.n_frames ? opt.n_frames : ui.nbr_of_iterations;
    const uint32_t batch = opt.frames_per_batch ? opt.frames_per_batch : total;
    for (uint32_t done = 0; done < total && !res.aborted;) {
        const uint32_t n = std::min(batch, total - done);
        ctx.check(srt_render_frames(ctx.h, opt.first_frame + done, n));
        done += n;
        float ms = 0.0f;
        uint64_t launches = 0;
        srt_last_render_stats(ctx.h, &ms, &launches);
        res.device_seconds += ms * 1e-3;
        res.kernel_launches += launches;
        if (opt.progress && !opt.progress((float)done / (float)total, opt.user)) res.aborted = true;
    }
    res.image = CustomImage(ui.width, ui.height);
    ctx.check(srt_resolve_rgba_f32(ctx.h, res.image.data.data()));
    ctx.check(srt_get_counters(ctx.h, &res.counters));
    return res;
}

// ============================================================== App::render protocol
// AppActions the render thread pushes for the UI thread (main.rs:1343-1348, :1366-1370) and the message it
// polls (AppToRenderMessages::AbortRender, main.rs:1351-1357).
struct AppAction {
    enum Kind { FrameUpdate, RenderingProgressUpdate, TrueTimeUpdate, DestroySender } kind;
    std::vector<uint8_t> image;  // FrameUpdate: RGBA8, image_float.clone().into()
    float progress = 0.0f;       // RenderingProgressUpdate: (frame_number + 1) / nbr_of_iterations
    double seconds = 0.0;        // TrueTimeUpdate
};
struct RenderChannel {
    std::mutex lock;
    std::vector<AppAction> action_list;     // Arc<Mutex<Vec<AppActions>>>
    std::atomic<bool> abort_render{false};  // the Receiver<AppToRenderMessages> side: set to request AbortRender
    std::atomic<bool> rendering{false};     // Arc<Mutex<bool>> `rendering`
};

// Sibling of App::render (main.rs:1327-1371) on the GPU: same action sequence -- per update FrameUpdate +
// RenderingProgressUpdate, abort polled once per update, at the end TrueTimeUpdate + DestroySender -- driven by
// srt_render_progressive (updates every frames_per_update frames; 1 = the reference's granularity).
inline bool render(Context& ctx, uint32_t nbr_of_iterations, RenderChannel& ch, uint32_t frames_per_update = 1,
                   uint32_t first_frame = 0) {
    ch.rendering = true;
    const auto begin_time = std::chrono::steady_clock::now();
    struct Env {
        Context* ctx;
        RenderChannel* ch;
    } env{&ctx, &ch};
    auto cb = [](void* user, uint32_t done, uint32_t total, const uint8_t* rgba8) -> int {
        Env* e = static_cast<Env*>(user);
        {
            std::lock_guard<std::mutex> g(e->ch->lock);
            AppAction frame{AppAction::FrameUpdate};
            if (rgba8) frame.image.assign(rgba8, rgba8 + (size_t)e->ctx->width * e->ctx->height * 4);
            e->ch->action_list.push_back(std::move(frame));
            AppAction prog{AppAction::RenderingProgressUpdate};
            prog.progress = (float)done / (float)total;
            e->ch->action_list.push_back(std::move(prog));
        }
        return e->ch->abort_render.exchange(false) ? 1 : 0;
    };
    const int rc = srt_render_progressive(ctx.h, first_frame, nbr_of_iterations, frames_per_update, 1, cb, &env);
    ch.rendering = false;
    {
        std::lock_guard<std::mutex> g(ch.lock);
        AppAction t{AppAction::TrueTimeUpdate};
        t.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - begin_time).count();
        ch.action_list.push_back(std::move(t));
        ch.action_list.push_back(AppAction{AppAction::DestroySender});
    }
    if (rc == SRT_ERR_ABORTED) return false;
    ctx.check(rc);
    return true;
}

}  // namespace srt_host
