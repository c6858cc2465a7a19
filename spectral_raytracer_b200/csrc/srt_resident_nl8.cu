// k_resident for spectra of 32 samples (8 quads) -- see srt_resident_inst.cuh.
#define SRT_RESIDENT_CAP 8
#define SRT_RESIDENT_FN resident_kernel_nl8
#include "srt_resident_inst.cuh"
