// srt_resident_inst.cuh -- included by every srt_resident_nl*.cu after defining SRT_RESIDENT_CAP (quads) and
// SRT_RESIDENT_FN (the selector's name): instantiates k_resident for that spectral capacity and defines the selector.
//
// Instantiations per capacity: production / exact math x pcg3d / Philox; linear scan: Cornell-like (diffuse, plain +
// rotated boxes) and its pair mode for one-light scenes, [default-like (specular + spheres) and prism-like (everything but
// the specular lobe): the default width only], everything; BVH: everything, the
// default width only (BVH scenes run the wavefront unless the resident integrator is asked for).  Partial widths
// (capacity > n_lambda4) exist for the capacities that have widths below them: 8 (24), 16 (40..56), 32 (72..120).
#define SRT_KERNELS_RESIDENT_ONLY 1
#include "srt_kernels.cuh"
#include "srt_resident.h"

namespace srt {
namespace {

template <class Accel, int NL4, int FEAT>
ResidentKernel pick_mode(bool exact, bool philox) {
    ResidentKernel k;
    k.cap = nl4_cap(NL4);
    k.min_blocks = resident_min_blocks(nl4_cap(NL4), (FEAT & kFeatPair) != 0, (FEAT & (kFeatSpecular | kFeatTransmissive)) != 0);
    k.feat = FEAT;
#ifdef SRT_DEV_MINIMAL  // developer builds for kernel A/B runs: production mode only (compiles in seconds)
    (void)exact; (void)philox;
    k.fn = reinterpret_cast<const void*>(&k_resident<Accel, false, false, NL4, FEAT>);
#else
    if (exact && philox) k.fn = reinterpret_cast<const void*>(&k_resident<Accel, true, true, NL4, FEAT>);
    else if (exact) k.fn = reinterpret_cast<const void*>(&k_resident<Accel, true, false, NL4, FEAT>);
    else if (philox) k.fn = reinterpret_cast<const void*>(&k_resident<Accel, false, true, NL4, FEAT>);
    else k.fn = reinterpret_cast<const void*>(&k_resident<Accel, false, false, NL4, FEAT>);
#endif
    return k;
}

constexpr int kCornellLike = kFeatRot, kDefaultLike = kFeatSpecular | kFeatSphere, kPrismLike = kFeatTransmissive | kFeatSphere | kFeatRot;

template <int NL4>
ResidentKernel pick(bool bvh, bool exact, bool philox, int need) {
    constexpr bool kDefaultWidth = NL4 == 8;  // NBR_OF_SPECTRUM_SAMPLES_DEFAULT = 32 (main.rs:32)
    if (bvh) {
        if constexpr (kDefaultWidth) return pick_mode<AccelBvh, NL4, kFeatAll>(exact, philox);
        return ResidentKernel{};
    }
    // the smallest instantiated superset of what the scene contains
    const bool one_light = (need & kFeatPair) != 0;  // (set by srt_create; not a lobe)
    need &= kFeatAll;
    if ((need & ~kCornellLike) == 0) {
        if (one_light) return pick_mode<AccelLinear, NL4, kCornellLike | kFeatPair>(exact, philox);  // pair mode, see k_resident
        return pick_mode<AccelLinear, NL4, kCornellLike>(exact, philox);
    }
    if constexpr (kDefaultWidth) {
        if ((need & ~kDefaultLike) == 0) return pick_mode<AccelLinear, NL4, kDefaultLike>(exact, philox);
        // (no metal in the scene: the prism config, +4 % over the kernel with every lobe -- 1374 -> 1429 M samples/s)
        if ((need & ~kPrismLike) == 0) return pick_mode<AccelLinear, NL4, kPrismLike>(exact, philox);
    }
    return pick_mode<AccelLinear, NL4, kFeatAll>(exact, philox);
}

}  // namespace

ResidentKernel SRT_RESIDENT_FN(bool bvh, bool exact, bool philox, int need, bool partial) {
    if (!partial) return pick<SRT_RESIDENT_CAP>(bvh, exact, philox, need);
#if SRT_RESIDENT_CAP >= 8
    return pick<-SRT_RESIDENT_CAP>(bvh, exact, philox, need);
#else
    return ResidentKernel{};
#endif
}

}  // namespace srt
