// single-precision instantiation of the FFT passes
#define SCB_T float
#include "fft_passes_impl.cuh"
