// double-precision instantiation of the FFT passes
#define SCB_T double
#include "fft_passes_impl.cuh"
namespace scb {
bool fft_len_supported(int N) { return N >= 8 && N <= 2048 && (N & (N - 1)) == 0; }
}
