// k_resident for spectra of 8 samples (2 quads) -- see srt_resident_inst.cuh.
#define SRT_RESIDENT_CAP 2
#define SRT_RESIDENT_FN resident_kernel_nl2
#include "srt_resident_inst.cuh"
