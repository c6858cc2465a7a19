/* Compiled by tests/test_host_and_abi.py with `gcc -std=c11 -pedantic -Wall -Wextra -Werror`: include/srt.h must be
 * plain C, and the layout of every struct that crosses the boundary is what the #[repr(C)] / ctypes mirrors assume
 * (the C++ side asserts the same numbers in spectral_raytracer_b200/csrc/srt_api.cu). */
#include <stddef.h>

#include "srt.h"

_Static_assert(sizeof(srt_object) == 92, "srt_object");
_Static_assert(offsetof(srt_object, kind) == 24 && offsetof(srt_object, center) == 28 && offsetof(srt_object, dims) == 40 &&
                   offsetof(srt_object, rot) == 52 && offsetof(srt_object, material) == 88, "srt_object fields");
_Static_assert(sizeof(srt_material) == 24 && offsetof(srt_material, reflectance) == 8 && offsetof(srt_material, ior_b) == 20, "srt_material");
_Static_assert(sizeof(srt_light) == 16 && offsetof(srt_light, spectrum) == 12, "srt_light");
_Static_assert(sizeof(srt_camera) == 40 && offsetof(srt_camera, fov_y_deg) == 36, "srt_camera");
_Static_assert(sizeof(srt_params) == 60 && offsetof(srt_params, device) == 44 && offsetof(srt_params, philox_seed_hi) == 56, "srt_params");
_Static_assert(sizeof(srt_counters) == 104 && offsetof(srt_counters, shadow_skipped) == 96, "srt_counters");

/* every entry point is callable from C with the documented argument types */
int srt_h_c11_uses_the_api(void) {
    srt_ctx* ctx = NULL;
    srt_params p = {0};
    srt_camera cam = {{0.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 1.0f}, {0.0f, 1.0f, 0.0f}, 60.0f};
    int rc = srt_create(&p, &cam, NULL, 0, NULL, 0, NULL, 0, NULL, 0, &ctx);
    if (rc == SRT_OK) {
        rc = srt_render_frames(ctx, 0, 1);
        srt_destroy(ctx);
    }
    (void)srt_last_error(NULL);
    return rc + (int)srt_abi_version() + srt_device_count();
}
