/* oracle/ref_shim/fftw3.h -- TEST INFRASTRUCTURE.  Stand-in for FFTW3's header
 * (absent from this image) so the reference's cpuLS.hpp:31 / cpuLS_main.cpp:28
 * include resolves; the implementation is oracle/fft_shim.c. */
#ifndef REF_SHIM_FFTW3_H
#define REF_SHIM_FFTW3_H
#include "../fft_shim.h"
#endif
