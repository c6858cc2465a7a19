// k_resident for spectra of 64 samples (16 quads) -- see srt_resident_inst.cuh.
#define SRT_RESIDENT_CAP 16
#define SRT_RESIDENT_FN resident_kernel_nl16
#include "srt_resident_inst.cuh"
