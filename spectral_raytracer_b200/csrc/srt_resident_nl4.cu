// k_resident for spectra of 16 samples (4 quads) -- see srt_resident_inst.cuh.
#define SRT_RESIDENT_CAP 4
#define SRT_RESIDENT_FN resident_kernel_nl4
#include "srt_resident_inst.cuh"
