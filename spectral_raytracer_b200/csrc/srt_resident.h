// srt_resident.h -- the resident integrator's kernels, one translation unit per spectral width so that the
// library builds in parallel (srt_resident_nl*.cu); srt_api.cu picks the instantiation for a scene.
#pragma once
#include <cstddef>

namespace srt {

// What srt_create needs to launch k_resident for a scene: the kernel (the smallest instantiated superset of the
// lobes / primitive kinds the scene contains), its spectral capacity in quads and the blocks per SM it was compiled
// for.  fn == nullptr: no resident kernel for this combination (the wavefront integrator is used).
struct ResidentKernel {
    const void* fn = nullptr;
    int cap = 0;         // nl4_cap(NL4): quads of 4 wavelengths the per-thread throughput storage holds
    int min_blocks = 0;  // __launch_bounds__ blocks per SM
    int feat = 0;        // kFeat* bits compiled in
};

// need = kFeat* bits of the scene.  Widths: nl4 in {2, 4, 8, 16, 32} run a kernel whose wavelength loops have
// compile-time trip counts; the widths in between (n_lambda = 24, 40, ... 120; spectrum.rs:37-38 allows every
// multiple of 8) run the next larger capacity with guarded loops.
ResidentKernel resident_kernel_nl2(bool bvh, bool exact, bool philox, int need, bool partial);
ResidentKernel resident_kernel_nl4(bool bvh, bool exact, bool philox, int need, bool partial);
ResidentKernel resident_kernel_nl8(bool bvh, bool exact, bool philox, int need, bool partial);
ResidentKernel resident_kernel_nl16(bool bvh, bool exact, bool philox, int need, bool partial);
ResidentKernel resident_kernel_nl32(bool bvh, bool exact, bool philox, int need, bool partial);

}  // namespace srt
