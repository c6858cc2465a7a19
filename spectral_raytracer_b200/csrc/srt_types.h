// srt_types.h -- device-side scene layout shared by the host API and the kernels.
// Everything here is derived from the C-ABI structs of include/srt.h at
// srt_create time; nothing in this file is visible to callers.
#pragma once
#include <cstdint>

#include <cuda_runtime.h>

namespace srt {

constexpr int kMaxLambda = 128;        // spectrum.rs:8  NBR_OF_SAMPLES_MAX
#ifndef SRT_MAX_CONST_OBJECTS
#define SRT_MAX_CONST_OBJECTS 64
#endif
constexpr int kMaxConstObjects = SRT_MAX_CONST_OBJECTS;   // linear-scan scenes live in the kernel-parameter constant bank
constexpr int kMaxLights = 8;
#ifndef SRT_BLOCK
#define SRT_BLOCK 256
#endif
constexpr int kBlock = SRT_BLOCK;

constexpr float kF32Delta = 0.00001f;          // shader.rs:7
constexpr float kNewRayOffset = 0.00001f;      // shader.rs:8
constexpr float kSpecularMinDistance = 0.0001f;  // shader.rs:14

enum : uint32_t { kPlainBox = 0, kSphere = 1, kRotatedBox = 2 };

// One primitive = 7 float4 (28 words), read with 128-bit loads.  The scene's primitives are
// stored sorted by kind (plain boxes, spheres, rotated boxes; original order inside a kind) so
// the linear scan runs three branch-free loops; `kind_orig` keeps the caller's object index for
// reporting and for the reference's tie-break (stable sort => lowest original index wins).
// For a sphere c = centre and h[0] = radius exactly as intersection_shader re-derives them from
// the bounds (shader.rs:305-306); for a rotated box c/h/rot are position, dims*0.5 and the
// row-major Rotation3 (shader.rs:560-571); the inverse rotation is applied by transposed indexing.
struct alignas(16) DevObject {
    float mn[3];
    uint32_t kind_orig;  // (original index << 2) | kind; linear-scan scenes: (orig << 10) | (staged index << 2) | kind
    float mx[3];
    uint32_t material;
    float c[3];
    float pad0;
    float h[3];
    float pad1;
    float rot[9];
    float pad2[3];
};
static_assert(sizeof(DevObject) == 28 * 4, "DevObject is 7 float4");
constexpr int kObjQuads = 7;

// Host-precomputed camera frame: the part of ray_generation_shader that does not
// depend on the pixel (shader.rs:272-278, :286-289), evaluated once in f32 with the
// same operation order, so the device only does IEEE add/mul/div/sqrt per pixel.
struct DevCamera {
    float pos[3];
    float fwd_focal[3];  // forward * focal_distance
    float right[3];
    float true_up[3];
    float aspect;
    float width_f, height_f;
};

struct DevBvhNode {  // 32 bytes
    float mn[3];
    uint32_t left_or_first;  // inner: index of left child (right = left+1); leaf: first primitive slot
    float mx[3];
    uint32_t count;          // 0 = inner node, >0 = leaf with `count` primitives
};

// Kernel-parameter block (passed __grid_constant__, so it sits in the constant bank
// and uniform reads are broadcast / folded into instruction operands).
struct SceneParams {
    DevCamera cam;
    uint32_t width, height, npix;
    uint32_t n_lambda, n_lambda4;
    uint32_t max_bounces, intended_frames;
    uint32_t n_objects, n_lights, n_materials;
    uint32_t n_plain, n_sphere, n_rot;      // primitives are sorted by kind in this order
    uint32_t philox_key[2];
    uint32_t tame;                          // every reflectance in [0,1], every emission in [0,1e18] (no NaN / negative radiance terms)
    float lambda_min, lambda_step;          // wavelength of sample i = lambda_min + lambda_step * i (dispersion extension)
    // materials (global memory; tiny, L1-resident)
    const float2* mat_params;               // [n_materials] (metallicness, roughness)
    const float4* mat_ext;                  // [n_materials] (transmissive, ior_a, ior_b, 0)
    const float4* mat_refl;                 // [n_lambda4][n_materials] reflectance, 4 wavelengths per entry
    // large scenes: primitives + BVH in global memory
    const DevObject* objects_g;
    const DevBvhNode* bvh_nodes;
    const uint32_t* bvh_prims;              // primitive slot -> object index
    const float4* bvh_leaf;                 // primitive slot -> (mn.xyz, kind_orig), (mx.xyz, object index): all a sphere or
                                            // plain-box test needs, 32 contiguous bytes per primitive in leaf order
    // face_towards() of every box-face normal (k_build_frames): 3 float4 per frame; frames 0..5 = the axis
    // normals of plain boxes, 6 + 6*r + face = rotated box r (r < n_frame_rot)
    const float4* frames;
    uint32_t n_frame_rot, pad_frames;
    // lights
    float light_pos[kMaxLights][3];
    float light_e[kMaxLights][kMaxLambda];  // raw emission spectra
    // small scenes: primitives in the constant bank
    DevObject obj[kMaxConstObjects];
};

// Path-pool state word (ray_d.w): bits 0..6 remaining bounces (Ray::max_bounces, <= 100 in the UI,
// main.rs:34), bit 7 parent lobe was specular, bit 8 some ancestor was diffuse (its max0 scrubs NaN,
// shader.rs:448), bit 9 fresh path (throughput == 1), bits 10..17 hero wavelength index + 1 of a path
// that went through a dispersive interface (0 = full spectrum; extension), bits 18..31 frame id
// relative to the render call.
constexpr uint32_t kRemMask = 0x7Fu;
constexpr uint32_t kFlagPrevSpec = 1u << 7;
constexpr uint32_t kFlagDiffAncestor = 1u << 8;
constexpr uint32_t kFlagFresh = 1u << 9;
constexpr int kHeroShift = 10;
constexpr uint32_t kHeroMask = 0xFFu << kHeroShift;
constexpr int kFrameShift = 18;
constexpr uint32_t kMaxFramesPerCall = 1u << (32 - kFrameShift);  // larger requests are split by srt_render_frames
enum : int { kLobeDiffuse = 0, kLobeSpecular = 1, kLobeTransmissive = 2 };

struct PathPool {
    float4* ray_o;  // origin.xyz, w = pixel index (bits)
    float4* ray_d;  // direction.xyz, w = state word (bits)
    float4* thr;    // [n_lambda4][capacity] path throughput
};

// Shadow-ray queue of the BVH wavefront: one region of `capacity` records per light, filled by k_shade and traced by
// k_shadow (srt_kernels.cuh).  a = origin.xyz, |L|; b = direction.xyz, pixel (bits); c = the light's factors
// (production math: c1*c2/|L|^2, -, -; exact math: |L|^2, c1, c2) and the slot that holds the path's advanced throughput
// (bits 0..29; bit 30: in the pool the path came from -- its last bounce -- instead of the next pool; bit 31: scrub).
struct ShadowQueue {
    float4* a = nullptr;
    float4* b = nullptr;
    float4* c = nullptr;
    uint32_t* count = nullptr;  // [kMaxLights] records queued in this iteration (reset by k_generate)
};

// Double-buffered control block (see srt_kernels.cuh: iteration `it` reads
// ctl[it&1] and builds ctl[(it+1)&1]).
struct PoolCtl {
    uint32_t count;             // live paths handed over by the previous shade
    uint32_t pad;
    unsigned long long next_sample;  // next (frame_local * npix + pixel) to generate
};

enum : int {
    kCtrSamples = 0, kCtrPrimary, kCtrContinuation, kCtrShadow, kCtrHits, kCtrSelfHits,
    kCtrMisses, kCtrLit, kCtrSpecHits, kCtrSpecDropped, kCtrShadowSkipped, kNumCounters
};
constexpr int kCtrStride = 32;  // one counter per 256-byte line
struct DevCounters {
    unsigned long long v[kNumCounters * kCtrStride];
};

}  // namespace srt
