/*
 * oracle/cpuls_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference receive path, function by function, with
 * runtime dimensions (the reference fixes them with -D macros,
 * ShMemSymBuff.hpp:42-67).  Every routine cites the reference lines it follows.
 * Arithmetic is plain fp32 in the reference's operation order; build with
 * -ffp-contract=off so no FMA contraction changes the rounding.
 *
 * Parity pin: tests/test_oracle_vs_ref.py compares this file bit-for-bit with
 * the reference's own cpuLS.hpp functions compiled from /root/reference
 * (oracle/_ref, see oracle/Makefile + oracle/ref_driver.cpp).  The FFT on both
 * sides is oracle/fft_shim.c because FFTW3f (un-versioned third-party
 * dependency of the reference, cpuLS.hpp:31,51) is absent from this image; the
 * shim is cross-checked against numpy.fft in tests/test_oracle_fft.py.
 * The reference ships no golden vectors or tests.
 *
 * Documented deviations from the reference as shipped:
 *  - the pilot read that is commented out at cpuLS.hpp:266-272 is restored
 *    (otherwise H == 0 and every output is NaN);
 *  - the hard QAM demapper is new (the reference has none).
 */
#include "cpuls_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "fft_shim.h"

size_t oracle_bits_row_bytes(int K, int qam_bits)
{
    return ((size_t)K * (size_t)qam_bits + 7u) / 8u;
}

/* cpuLS.hpp:101-112: temp <- second half; first half -> tail; temp -> head */
void oracle_pilot_to_bin_order(const oc_complex *pilot_asc, oc_complex *x_bin, int K)
{
    int half_lo = (K - 1) / 2; /* elements moved to the front */
    int half_hi = (K + 1) / 2; /* elements moved to the back */
    oc_complex *tmp = (oc_complex *)malloc(sizeof(oc_complex) * (size_t)(K > 0 ? K : 1));
    memcpy(tmp, pilot_asc, sizeof(oc_complex) * (size_t)K);
    memcpy(x_bin, &tmp[half_hi], sizeof(oc_complex) * (size_t)half_lo);
    memcpy(&x_bin[half_lo], tmp, sizeof(oc_complex) * (size_t)half_hi);
    free(tmp);
}

/* cpuLS.hpp:135-149 */
void oracle_shift_one_row(oc_complex *row, int K)
{
    int n_tail = (K + 1) / 2;
    int n_head = (K - 1) / 2;
    oc_complex *tmp = (oc_complex *)malloc(sizeof(oc_complex) * (size_t)n_tail);
    memmove(tmp, &row[n_head], sizeof(oc_complex) * (size_t)n_tail);
    memmove(&row[n_tail], row, sizeof(oc_complex) * (size_t)n_head);
    memmove(row, tmp, sizeof(oc_complex) * (size_t)n_tail);
    free(tmp);
}

/* ShMemSymBuff.hpp:281-294: Y[a][n] = slot[a][n + C] */
static void strip_prefix(const oc_complex *rx_sym, oc_complex *y, int A, int N, int C)
{
    int a;
    for (a = 0; a < A; a++)
        memcpy(&y[(size_t)a * N], &rx_sym[(size_t)a * (N + C) + C], sizeof(oc_complex) * (size_t)N);
}

/* cpuLS.hpp:165-174 */
static void fft_one_row(oc_complex *y, int N, int row)
{
    float *p = (float *)&y[(size_t)row * N];
    oracle_fft_f32(N, p, p, FFTW_FORWARD);
}

/* cpuLS.hpp:233-244 */
static void divide_one_row(oc_complex *a, const oc_complex *b, int cols, int row)
{
    int j;
    for (j = 0; j < cols; j++) {
        float fxa = a[(size_t)row * cols + j].real;
        float fxb = a[(size_t)row * cols + j].imag;
        float fya = b[j].real;
        float fyb = b[j].imag;
        a[(size_t)row * cols + j].real = ((fxa * fya + fxb * fyb) / (fya * fya + fyb * fyb));
        a[(size_t)row * cols + j].imag = ((fxb * fya - fxa * fyb) / (fya * fya + fyb * fyb));
    }
}

/* cpuLS.hpp:211-228; the reference stores the sum into X[j].real, we keep a
 * separate float row (same values). */
static void find_dist_sqrd(const oc_complex *h, float *hsqrd, int rows, int cols)
{
    int i, j;
    for (j = 0; j < cols; j++)
        hsqrd[j] = (h[j].real * h[j].real) + (h[j].imag * h[j].imag);
    for (i = 1; i < rows; i++)
        for (j = 0; j < cols; j++)
            hsqrd[j] = hsqrd[j] + (h[(size_t)i * cols + j].real * h[(size_t)i * cols + j].real) +
                       (h[(size_t)i * cols + j].imag * h[(size_t)i * cols + j].imag);
}

/* cpuLS.hpp:187-208 (cols here is N; rows of width N-1) */
static void matrix_mult_then_sum(const oc_complex *y, const oc_complex *hconj, oc_complex *yf,
                                 int rows, int cols)
{
    int i, j;
    for (i = 0; i < rows; i++) {
        for (j = 0; j < cols - 1; j++) {
            float yr = y[(size_t)i * (cols - 1) + j].real;
            float yi = y[(size_t)i * (cols - 1) + j].imag;
            float hr = hconj[(size_t)i * (cols - 1) + j].real;
            float hi = hconj[(size_t)i * (cols - 1) + j].imag;
            if (i == 0) {
                yf[j].real = 0;
                yf[j].imag = 0;
            }
            yf[j].real = yf[j].real + (yr * hr - yi * hi);
            yf[j].imag = yf[j].imag + (yr * hi + yi * hr);
        }
    }
}

void oracle_first_vector(const oc_complex *rx_sym, const oc_complex *x_bin, oc_complex *hconj,
                         float *hsqrd, int A, int N, int C)
{
    int K = N - 1, row, j;
    oc_complex *y = (oc_complex *)malloc(sizeof(oc_complex) * (size_t)A * N);
    /* restored pilot read, cpuLS.hpp:266-272 -> readNextSymbol -> CP strip */
    strip_prefix(rx_sym, y, A, N, C);
    /* cpuLS.hpp:278-281 */
    for (row = 0; row < A; row++) fft_one_row(y, N, row);
    /* cpuLS.hpp:290-299: drop bin 0, divide by X */
    for (row = 0; row < A; row++) {
        memcpy(&hconj[(size_t)row * K], &y[(size_t)row * N + 1], sizeof(oc_complex) * (size_t)K);
        divide_one_row(hconj, x_bin, K, row);
    }
    /* cpuLS.hpp:303-307: conjugate */
    for (row = 0; row < A; row++)
        for (j = 0; j < K; j++)
            hconj[(size_t)row * K + j].imag = -1 * hconj[(size_t)row * K + j].imag;
    /* cpuLS.hpp:311 */
    find_dist_sqrd(hconj, hsqrd, A, K);
    free(y);
}

void oracle_one_symbol(const oc_complex *rx_sym, const oc_complex *hconj, const float *hsqrd,
                       oc_complex *out_sorted, int A, int N, int C)
{
    int K = N - 1, row, j;
    oc_complex *y = (oc_complex *)malloc(sizeof(oc_complex) * (size_t)A * N);
    oc_complex *ytemp = (oc_complex *)malloc(sizeof(oc_complex) * (size_t)A * K);
    /* cpuLS.hpp:321-328 -> ShMemSymBuff.hpp:237-295 */
    strip_prefix(rx_sym, y, A, N, C);
    /* cpuLS.hpp:342-345 */
    for (row = 0; row < A; row++) fft_one_row(y, N, row);
    /* cpuLS.hpp:354-357 */
    for (row = 0; row < A; row++)
        memcpy(&ytemp[(size_t)row * K], &y[(size_t)row * N + 1], sizeof(oc_complex) * (size_t)K);
    /* cpuLS.hpp:360 */
    matrix_mult_then_sum(ytemp, hconj, out_sorted, A, N);
    /* cpuLS.hpp:364-367 */
    for (j = 0; j < K; j++) {
        out_sorted[j].real = out_sorted[j].real / hsqrd[j];
        out_sorted[j].imag = out_sorted[j].imag / hsqrd[j];
    }
    /* cpuLS.hpp:368 */
    oracle_shift_one_row(out_sorted, K);
    free(y);
    free(ytemp);
}

/* Hard decision, Gray-mapped square QAM of 3GPP TS 38.211 5.1.3-5.1.5, unit
 * average power.  Strict comparisons: +-0.0 -> bit 0.  Bits are packed LSB
 * first; symbol i occupies stream bits [i*b, i*b+b).  The same fp32 constants
 * appear in the CUDA kernel (csrc/lsmrc_kernels.cuh demap_symbol). */
#define QAM16_T ((float)0.6324555320336759)   /* 2/sqrt(10) */
#define QAM64_T4 ((float)0.6172133998483676)  /* 4/sqrt(42) */
#define QAM64_T2 ((float)0.3086066999241838)  /* 2/sqrt(42) */

static unsigned demap_one(float re, float im, int qam_bits)
{
    unsigned v = 0;
    float are = fabsf(re), aim = fabsf(im);
    if (re < 0.0f) v |= 1u;
    if (im < 0.0f) v |= 2u;
    if (qam_bits == 4) {
        if (are > QAM16_T) v |= 4u;
        if (aim > QAM16_T) v |= 8u;
    } else if (qam_bits == 6) {
        if (are > QAM64_T4) v |= 4u;
        if (aim > QAM64_T4) v |= 8u;
        if (fabsf(are - QAM64_T4) > QAM64_T2) v |= 16u;
        if (fabsf(aim - QAM64_T4) > QAM64_T2) v |= 32u;
    }
    return v;
}

void oracle_demap_row(const oc_complex *sym, int K, int qam_bits, uint8_t *packed, uint8_t *idx)
{
    size_t nbytes = oracle_bits_row_bytes(K, qam_bits);
    int i, q;
    if (packed) memset(packed, 0, nbytes);
    for (i = 0; i < K; i++) {
        unsigned v = demap_one(sym[i].real, sym[i].imag, qam_bits);
        if (idx) idx[i] = (uint8_t)v;
        if (packed) {
            for (q = 0; q < qam_bits; q++) {
                size_t pos = (size_t)i * (size_t)qam_bits + (size_t)q;
                if (v & (1u << q)) packed[pos >> 3] |= (uint8_t)(1u << (pos & 7u));
            }
        }
    }
}

/* Max-log LLRs (not in the reference).  Piecewise-linear form for the mapping above;
 * LLR > 0 <=> bit 0; rho = sum|H|^2 / noise_var.  Same constants and order as soft_symbol() in the kernel. */
static void soft_one(float re, float im, float rho, int qam_bits, float *out)
{
    const float a = qam_bits == 2 ? (float)0.7071067811865476 : qam_bits == 4 ? (float)0.31622776601683794 : (float)0.1543033499620919;
    const float g = (4.0f * a) * rho;
    out[0] = g * re;
    out[1] = g * im;
    if (qam_bits == 4) {
        out[2] = g * (2.0f * a - fabsf(re));
        out[3] = g * (2.0f * a - fabsf(im));
    } else if (qam_bits == 6) {
        out[2] = g * (4.0f * a - fabsf(re));
        out[3] = g * (4.0f * a - fabsf(im));
        out[4] = g * (2.0f * a - fabsf(fabsf(re) - 4.0f * a));
        out[5] = g * (2.0f * a - fabsf(fabsf(im) - 4.0f * a));
    }
}

/* sym [K] combined symbols (ascending frequency), hsqrd_bin [K] in FFT-bin order, llr [K][b] */
void oracle_soft_demap_row(const oc_complex *sym, const float *hsqrd_bin, int K, int qam_bits, float noise_var, float *llr)
{
    int i;
    const float inv = 1.0f / noise_var;
    for (i = 0; i < K; i++) {
        const int bin_idx = (i + (K - 1) / 2) % K; /* inverse of shiftOneRow: sorted[i] = out[(i+(K-1)/2) mod K] */
        soft_one(sym[i].real, sym[i].imag, hsqrd_bin[bin_idx] * inv, qam_bits, llr + (size_t)i * qam_bits);
    }
}

/* nearest constellation level of one axis, from the same comparisons as demap_one() */
static float slice_axis(float u, int qam_bits)
{
    const float au = fabsf(u);
    float lev;
    if (qam_bits == 2) {
        lev = (float)0.7071067811865476;
    } else if (qam_bits == 4) {
        const float a = (float)0.31622776601683794;
        lev = (au > (float)0.6324555320336759) ? 3.0f * a : a;
    } else {
        const float a = (float)0.1543033499620919;
        const float t4 = (float)0.6172133998483676, t2 = (float)0.3086066999241838;
        const int outer = au > t4, far = fabsf(au - t4) > t2;
        lev = outer ? (far ? 7.0f * a : 5.0f * a) : (far ? a : 3.0f * a);
    }
    return u < 0.0f ? -lev : lev;
}

/* Decision-directed noise-variance estimate of one frame (new):
 *   mean over data symbols s and subcarriers i of  sum|H|^2[bin(i)] * |y[s][i] - slice(y[s][i])|^2
 * (after MRC the symbol error has variance noise_var / sum|H|^2).  combined [n_rows][K] ascending frequency,
 * hsqrd_bin [K] FFT-bin order.  Accumulated in double: the CUDA kernel's fp32 tree sum is compared to 1e-5. */
double oracle_noise_var_frame(const oc_complex *combined, const float *hsqrd_bin, int K, int n_rows, int qam_bits)
{
    double acc = 0.0;
    int s, i;
    for (s = 0; s < n_rows; s++)
        for (i = 0; i < K; i++) {
            const oc_complex y = combined[(size_t)s * K + i];
            const float dr = y.real - slice_axis(y.real, qam_bits), di = y.imag - slice_axis(y.imag, qam_bits);
            const float e2 = dr * dr + di * di;
            acc += (double)(hsqrd_bin[(i + (K - 1) / 2) % K] * e2);
        }
    return acc / ((double)n_rows * (double)K);
}

typedef struct {
    const oc_complex *rx;
    const oc_complex *x_bin;
    int f0, f1, S, A, N, C, b;
    oc_complex *hconj;
    float *hsqrd;
    oc_complex *combined;
    uint8_t *bits;
} frame_job;

/* frame loop of cpuLS_main.cpp:80-93: symbol 0 -> firstVector, 1..S-1 -> doOneSymbol */
static void *frame_worker(void *arg)
{
    frame_job *jb = (frame_job *)arg;
    int K = jb->N - 1, f, s;
    size_t slot = (size_t)jb->A * (size_t)(jb->N + jb->C);
    size_t row_bytes = oracle_bits_row_bytes(K, jb->b);
    for (f = jb->f0; f < jb->f1; f++) {
        const oc_complex *frame = jb->rx + (size_t)f * jb->S * slot;
        oc_complex *hc = jb->hconj + (size_t)f * jb->A * K;
        float *hs = jb->hsqrd + (size_t)f * K;
        oracle_first_vector(frame, jb->x_bin, hc, hs, jb->A, jb->N, jb->C);
        for (s = 1; s < jb->S; s++) {
            oc_complex *out = jb->combined + ((size_t)f * (jb->S - 1) + (size_t)(s - 1)) * K;
            oracle_one_symbol(frame + (size_t)s * slot, hc, hs, out, jb->A, jb->N, jb->C);
            if (jb->bits)
                oracle_demap_row(out, K, jb->b,
                                 jb->bits + ((size_t)f * (jb->S - 1) + (size_t)(s - 1)) * row_bytes,
                                 NULL);
        }
    }
    return NULL;
}

int oracle_demod_frames(const oc_complex *rx, const oc_complex *pilot_asc, int F, int S, int A,
                        int N, int C, int qam_bits, oc_complex *hconj, float *hsqrd,
                        oc_complex *combined, uint8_t *bits, int n_threads)
{
    int K = N - 1, t;
    oc_complex *x_bin;
    pthread_t *th;
    frame_job *jobs;
    if (F < 0 || S < 1 || A < 1 || N < 2 || (N & (N - 1)) || C < 0) return -1;
    if (qam_bits != 2 && qam_bits != 4 && qam_bits != 6) return -2;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > F) n_threads = F > 0 ? F : 1;
    x_bin = (oc_complex *)malloc(sizeof(oc_complex) * (size_t)K);
    oracle_pilot_to_bin_order(pilot_asc, x_bin, K);
    th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    jobs = (frame_job *)malloc(sizeof(frame_job) * (size_t)n_threads);
    for (t = 0; t < n_threads; t++) {
        frame_job *jb = &jobs[t];
        jb->rx = rx;
        jb->x_bin = x_bin;
        jb->f0 = (int)(((long long)F * t) / n_threads);
        jb->f1 = (int)(((long long)F * (t + 1)) / n_threads);
        jb->S = S;
        jb->A = A;
        jb->N = N;
        jb->C = C;
        jb->b = qam_bits;
        jb->hconj = hconj;
        jb->hsqrd = hsqrd;
        jb->combined = combined;
        jb->bits = bits;
        if (n_threads == 1)
            frame_worker(jb);
        else
            pthread_create(&th[t], NULL, frame_worker, jb);
    }
    if (n_threads > 1)
        for (t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    free(jobs);
    free(th);
    free(x_bin);
    return 0;
}
