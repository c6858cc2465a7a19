// k_resident for spectra of 128 samples (32 quads) -- see srt_resident_inst.cuh.
#define SRT_RESIDENT_CAP 32
#define SRT_RESIDENT_FN resident_kernel_nl32
#include "srt_resident_inst.cuh"
