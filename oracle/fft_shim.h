/*
 * oracle/fft_shim.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * FFTW3f-compatible subset (see fft_shim.c); stands in for <fftw3.h>, which the
 * reference includes at cpuLS.hpp:31 / cpuLS_main.cpp:28 but this image lacks.
 */
#ifndef ORACLE_FFT_SHIM_H
#define ORACLE_FFT_SHIM_H

#ifdef __cplusplus
extern "C" {
#endif

typedef float fftwf_complex[2];
typedef struct fftwf_plan_s *fftwf_plan;

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)

fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags);
void fftwf_execute(const fftwf_plan p);
void fftwf_destroy_plan(fftwf_plan p);

/* direct entry points for the restated oracle and the cross-check tests */
void oracle_fft_f32(int n, const float *in, float *out, int sign);
void oracle_fft_f64(int n, const double *in, double *out, int sign);

#ifdef __cplusplus
}
#endif
#endif
