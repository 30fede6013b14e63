/*
 * oracle/fft_shim.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A small single-precision CPU FFT that stands in for FFTW3f, which the
 * reference's CPU path calls (cpuLS.hpp:157-159 backward, :170-172 forward)
 * but which is not installed in this image (no fftw3.h / libfftw3f, no network).
 * It exposes the three FFTW entry points the reference uses
 * (fftwf_plan_dft_1d / fftwf_execute / fftwf_destroy_plan) so that
 *   (a) the reference's own cpuLS.hpp compiles and links against it unmodified
 *       (oracle/_ref, built by oracle/Makefile), and
 *   (b) the restated oracle (cpuls_oracle.c) runs the very same transform,
 *       so the two can be compared bit for bit.
 *
 * Transform definition (same as FFTW's): unnormalised,
 *   FFTW_FORWARD  (-1):  X[k] = sum_n x[n] exp(-2*pi*i*n*k/N)
 *   FFTW_BACKWARD (+1):  X[k] = sum_n x[n] exp(+2*pi*i*n*k/N)
 * Arithmetic: fp32 butterflies (iterative radix-2 decimation in time), twiddles
 * generated in double and rounded once to fp32.  A double-precision variant
 * (oracle_fft_f64) is provided for cross-checks.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may link or call this file.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#include "fft_shim.h"

#define SHIM_MAX_LOG2 24

typedef struct {
    int n;
    float *tw;          /* n/2 complex twiddles, exp(-2*pi*i*j/n) */
    unsigned *rev;      /* bit-reversal permutation */
} shim_table;

static shim_table g_tables[SHIM_MAX_LOG2 + 1];
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;

static int ilog2_exact(int n)
{
    int l = 0;
    if (n <= 0) return -1;
    while ((1 << l) < n) l++;
    return ((1 << l) == n) ? l : -1;
}

static const shim_table *get_table(int n)
{
    int l = ilog2_exact(n);
    if (l < 0 || l > SHIM_MAX_LOG2) return NULL;
    pthread_mutex_lock(&g_lock);
    if (g_tables[l].n != n) {
        shim_table t;
        int j, b;
        t.n = n;
        t.tw = (float *)malloc(sizeof(float) * (size_t)(n > 1 ? n : 2));
        t.rev = (unsigned *)malloc(sizeof(unsigned) * (size_t)n);
        for (j = 0; j < n / 2; j++) {
            double ang = -2.0 * M_PI * (double)j / (double)n;
            t.tw[2 * j] = (float)cos(ang);
            t.tw[2 * j + 1] = (float)sin(ang);
        }
        for (j = 0; j < n; j++) {
            unsigned r = 0;
            for (b = 0; b < l; b++)
                if (j & (1 << b)) r |= 1u << (l - 1 - b);
            t.rev[j] = r;
        }
        g_tables[l] = t;
    }
    pthread_mutex_unlock(&g_lock);
    return &g_tables[l];
}

/* in-place or out-of-place, interleaved (re,im) fp32 */
static void shim_fft_f32(int n, const float *in, float *out, int sign)
{
    const shim_table *t = get_table(n);
    int j, len;
    if (!t) abort();
    if (in == out) {
        for (j = 0; j < n; j++) {
            unsigned r = t->rev[j];
            if ((unsigned)j < r) {
                float a = out[2 * j], b = out[2 * j + 1];
                out[2 * j] = out[2 * r];
                out[2 * j + 1] = out[2 * r + 1];
                out[2 * r] = a;
                out[2 * r + 1] = b;
            }
        }
    } else {
        for (j = 0; j < n; j++) {
            unsigned r = t->rev[j];
            out[2 * r] = in[2 * j];
            out[2 * r + 1] = in[2 * j + 1];
        }
    }
    for (len = 2; len <= n; len <<= 1) {
        int half = len >> 1;
        int step = n / len;
        int base, k;
        for (base = 0; base < n; base += len) {
            for (k = 0; k < half; k++) {
                float wr = t->tw[2 * (k * step)];
                float wi = t->tw[2 * (k * step) + 1];
                float *p = out + 2 * (base + k);
                float *q = out + 2 * (base + k + half);
                float xr, xi, ur, ui;
                if (sign > 0) wi = -wi;
                xr = q[0] * wr - q[1] * wi;
                xi = q[0] * wi + q[1] * wr;
                ur = p[0];
                ui = p[1];
                p[0] = ur + xr;
                p[1] = ui + xi;
                q[0] = ur - xr;
                q[1] = ui - xi;
            }
        }
    }
}

void oracle_fft_f32(int n, const float *in, float *out, int sign)
{
    shim_fft_f32(n, in, out, sign);
}

/* double-precision direct evaluation path for cross-checks: radix-2 in f64 */
void oracle_fft_f64(int n, const double *in, double *out, int sign)
{
    int l = ilog2_exact(n), j, b, len;
    if (l < 0) abort();
    for (j = 0; j < n; j++) {
        unsigned r = 0;
        for (b = 0; b < l; b++)
            if (j & (1 << b)) r |= 1u << (l - 1 - b);
        out[2 * r] = in[2 * j];
        out[2 * r + 1] = in[2 * j + 1];
    }
    for (len = 2; len <= n; len <<= 1) {
        int half = len >> 1, base, k;
        for (base = 0; base < n; base += len) {
            for (k = 0; k < half; k++) {
                double ang = (sign > 0 ? 2.0 : -2.0) * M_PI * (double)k / (double)len;
                double wr = cos(ang), wi = sin(ang);
                double *p = out + 2 * (base + k);
                double *q = out + 2 * (base + k + half);
                double xr = q[0] * wr - q[1] * wi;
                double xi = q[0] * wi + q[1] * wr;
                double ur = p[0], ui = p[1];
                p[0] = ur + xr;
                p[1] = ui + xi;
                q[0] = ur - xr;
                q[1] = ui - xi;
            }
        }
    }
}

/* ---- the FFTW3f surface the reference calls (cpuLS.hpp:157-159,170-172) ---- */

struct fftwf_plan_s {
    int n;
    float *in;
    float *out;
    int sign;
};

fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags)
{
    fftwf_plan p = (fftwf_plan)malloc(sizeof(struct fftwf_plan_s));
    (void)flags;
    p->n = n;
    p->in = (float *)in;
    p->out = (float *)out;
    p->sign = sign;
    (void)get_table(n);
    return p;
}

void fftwf_execute(const fftwf_plan p)
{
    shim_fft_f32(p->n, p->in, p->out, p->sign);
}

void fftwf_destroy_plan(fftwf_plan p)
{
    free(p);
}
