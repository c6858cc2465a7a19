// srt_headless.cpp -- command-line sibling of App::dispatch_render (main.rs:1376): renders a preset
// without the eframe UI and writes the image -- the RGBA8 conversion of custom_image.rs:92-101, as a PNG
// like the reference's "Save Image" does through DynamicImage::save (main.rs:2325-2326), or as a binary PPM
// (RGB) when the name ends in .ppm -- plus a one-line JSON summary.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "srt_host.hpp"

using namespace srt_host;

namespace {
bool ends_with(const std::string& s, const char* suffix) {
    const size_t n = std::strlen(suffix);
    return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0;
}
}  // namespace

static void usage() {
    std::fprintf(stderr,
                 "usage: srt_headless [--scene default|cornell|spheres|prism] [--spheres N] [--width W] [--height H]\n"
                 "                    [--spp N] [--bounces N] [--nlambda N] [--rng pcg3d|philox] [--math fast|exact]\n"
                 "                    [--device D] [--out image.png | image.ppm]\n");
}

int main(int argc, char** argv) {
    std::string scene = "default", out;
    uint32_t n_spheres = 10000, width = 0, height = 0, spp = 0, bounces = 0, nlambda = 32;
    RenderOptions opt;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() -> const char* {
            if (i + 1 >= argc) { usage(); std::exit(2); }
            return argv[++i];
        };
        if (a == "--scene") scene = next();
        else if (a == "--spheres") n_spheres = (uint32_t)std::atoi(next());
        else if (a == "--width") width = (uint32_t)std::atoi(next());
        else if (a == "--height") height = (uint32_t)std::atoi(next());
        else if (a == "--spp") spp = (uint32_t)std::atoi(next());
        else if (a == "--bounces") bounces = (uint32_t)std::atoi(next());
        else if (a == "--nlambda") nlambda = (uint32_t)std::atoi(next());
        else if (a == "--rng") opt.rng_mode = std::string(next()) == "philox" ? SRT_RNG_PHILOX : SRT_RNG_PCG3D_REFERENCE;
        else if (a == "--math") opt.math_mode = std::string(next()) == "exact" ? SRT_MATH_EXACT : SRT_MATH_FAST;
        else if (a == "--device") opt.device = std::atoi(next());
        else if (a == "--out") out = next();
        else { usage(); return 2; }
    }
    try {
        UIFields ui = scene == "cornell" ? UIFields::cornell_box(nlambda)
                      : scene == "spheres" ? UIFields::random_spheres(n_spheres, nlambda)
                      : scene == "prism"   ? UIFields::prism(nlambda)
                                           : UIFields::default_scene(nlambda);
        if (width) ui.width = width;
        if (height) ui.height = height;
        if (spp) ui.nbr_of_iterations = spp;
        if (bounces) ui.nbr_of_ray_bounces = bounces;
        opt.progress = [](float f, void*) { std::fprintf(stderr, "\rprogress %5.1f %%", 100.0f * f); return true; };
        auto t0 = std::chrono::steady_clock::now();
        RenderResult r = dispatch_render_headless(ui, opt);
        double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::fprintf(stderr, "\n");
        if (!out.empty()) {
            std::vector<uint8_t> px = r.image.to_rgba8();
            if (ends_with(out, ".ppm")) {
                FILE* f = std::fopen(out.c_str(), "wb");
                if (!f) { std::perror("fopen"); return 1; }
                std::fprintf(f, "P6\n%u %u\n255\n", ui.width, ui.height);
                for (size_t i = 0; i < (size_t)ui.width * ui.height; ++i) std::fwrite(&px[4 * i], 1, 3, f);
                std::fclose(f);
            } else if (!write_png_rgba8(out, ui.width, ui.height, px)) {
                std::fprintf(stderr, "srt_headless: cannot write %s\n", out.c_str());
                return 1;
            }
        }
        const double samples = (double)r.counters.samples;
        const double rays = (double)(r.counters.rays_primary + r.counters.rays_continuation + r.counters.rays_shadow);
        std::printf("{\"scene\": \"%s\", \"width\": %u, \"height\": %u, \"spp\": %u, \"samples_per_s\": %.6g, "
                    "\"mrays_per_s\": %.6g, \"device_s\": %.6g, \"wall_s\": %.6g, \"kernel_launches\": %llu}\n",
                    scene.c_str(), ui.width, ui.height, ui.nbr_of_iterations, samples / r.device_seconds,
                    rays / r.device_seconds / 1e6, r.device_seconds, wall, (unsigned long long)r.kernel_launches);
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "srt_headless: %s\n", e.what());
        return 1;
    }
}
