/*
 * oracle/zf_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's multi-user zero-forcing helpers:
 *   createZeroForcingMatrix   cpuLS.hpp:415-447  (with rotCube :400-413)
 *   multiplyWithChannelInv    cpuLS.hpp:449-463
 * Nothing in the reference calls them and they need CBLAS + LAPACK (cgemm, cgetrf, cgetri, cgemv), which are
 * not installed here: PARITY UNPINNED.  What they compute is mathematically fixed, so the restatement follows
 * the call sequence with plain loops and an LU-free Gauss-Jordan inverse with partial pivoting in double
 * precision (any stable inverse agrees with LAPACK's to rounding):
 *   per subcarrier k:  Xk = X[:, :, k]  (users x antennas, after rotCube)
 *                      G  = Xk * Xk^H                       (cgemm NoTrans/ConjTrans, :437)
 *                      Gi = inv(G)                          (cgetrf + cgetri, :438-439)
 *                      Hk = Xk^H * Gi  (antennas x users)   (cgemm ConjTrans/NoTrans, :440), column-major, ld = antennas
 *   apply:             HX[a][k] = sum_u Hk[a + A*u] * Xd[u][k]   (cgemv per subcarrier, :459)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "cpuls_oracle.h"

typedef struct {
    double re, im;
} zc;
static zc zmul(zc a, zc b) { zc r = {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; return r; }
static zc zdiv(zc a, zc b)
{
    const double d = b.re * b.re + b.im * b.im;
    zc r = {(a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d};
    return r;
}

/* in-place inverse of the U x U column-major matrix g (Gauss-Jordan, partial pivoting); returns 0, or -1 if singular */
static int zinv(zc *g, int U)
{
    zc *w = (zc *)calloc((size_t)U * 2 * U, sizeof(zc)); /* [row][2U] augmented, row-major */
    int i, j, c;
    double gmax = 0.0;
    for (i = 0; i < U; i++) {
        for (j = 0; j < U; j++) w[i * 2 * U + j] = g[j * U + i];
        w[i * 2 * U + U + i].re = 1.0;
    }
    for (i = 0; i < U; i++)
        for (j = 0; j < U; j++) {
            const double m = fabs(w[i * 2 * U + j].re) + fabs(w[i * 2 * U + j].im);
            if (m > gmax) gmax = m;
        }
    for (c = 0; c < U; c++) {
        int p = c;
        double best = -1.0;
        for (i = c; i < U; i++) {
            const double m = fabs(w[i * 2 * U + c].re) + fabs(w[i * 2 * U + c].im); /* icamax's |re|+|im| */
            if (m > best) {
                best = m;
                p = i;
            }
        }
        if (!(best > 1e-6 * gmax)) { /* same rule as the kernel: below fp32 rounding noise of the Gram matrix */
            free(w);
            return -1;
        }
        if (p != c)
            for (j = 0; j < 2 * U; j++) {
                zc t = w[c * 2 * U + j];
                w[c * 2 * U + j] = w[p * 2 * U + j];
                w[p * 2 * U + j] = t;
            }
        {
            const zc piv = w[c * 2 * U + c];
            for (j = 0; j < 2 * U; j++) w[c * 2 * U + j] = zdiv(w[c * 2 * U + j], piv);
        }
        for (i = 0; i < U; i++) {
            if (i == c) continue;
            {
                const zc f = w[i * 2 * U + c];
                for (j = 0; j < 2 * U; j++) {
                    const zc t = zmul(f, w[c * 2 * U + j]);
                    w[i * 2 * U + j].re -= t.re;
                    w[i * 2 * U + j].im -= t.im;
                }
            }
        }
    }
    for (i = 0; i < U; i++)
        for (j = 0; j < U; j++) g[j * U + i] = w[i * 2 * U + U + j];
    free(w);
    return 0;
}

/* X [U][A][K] (user-major, as handed to createZeroForcingMatrix before rotCube); Hzf [K][U][A] = per subcarrier the
 * A x U column-major matrix of cpuLS.hpp:440.  Returns the number of singular subcarriers (their block is zeroed). */
int oracle_zf_create(const oc_complex *X, oc_complex *Hzf, int A, int K, int U)
{
    zc *xk = (zc *)malloc((size_t)U * A * sizeof(zc)); /* users x antennas, column-major ld = U (after rotCube) */
    zc *g = (zc *)malloc((size_t)U * U * sizeof(zc));
    int k, u, v, a, bad = 0;
    for (k = 0; k < K; k++) {
        for (a = 0; a < A; a++)
            for (u = 0; u < U; u++) {
                const oc_complex s = X[((size_t)u * A + a) * K + k];
                xk[a * U + u].re = s.real;
                xk[a * U + u].im = s.imag;
            }
        for (v = 0; v < U; v++)
            for (u = 0; u < U; u++) { /* G[u][v] = sum_a x[u,a] conj(x[v,a]) */
                zc acc = {0.0, 0.0};
                for (a = 0; a < A; a++) {
                    zc cj = {xk[a * U + v].re, -xk[a * U + v].im};
                    zc t = zmul(xk[a * U + u], cj);
                    acc.re += t.re;
                    acc.im += t.im;
                }
                g[v * U + u] = acc;
            }
        if (zinv(g, U) != 0) {
            memset(Hzf + (size_t)k * A * U, 0, (size_t)A * U * sizeof(oc_complex));
            bad++;
            continue;
        }
        for (u = 0; u < U; u++)
            for (a = 0; a < A; a++) { /* H[a][u] = sum_v conj(x[v,a]) Gi[v][u] */
                zc acc = {0.0, 0.0};
                for (v = 0; v < U; v++) {
                    zc cj = {xk[a * U + v].re, -xk[a * U + v].im};
                    zc t = zmul(cj, g[u * U + v]);
                    acc.re += t.re;
                    acc.im += t.im;
                }
                Hzf[(size_t)k * A * U + (size_t)u * A + a].real = (float)acc.re;
                Hzf[(size_t)k * A * U + (size_t)u * A + a].imag = (float)acc.im;
            }
    }
    free(xk);
    free(g);
    return bad;
}

/* Xd [U][K] user symbols, Hzf [K][U][A]; HX [A][K] (cpuLS.hpp:449-463) */
void oracle_zf_apply(const oc_complex *Hzf, const oc_complex *Xd, oc_complex *HX, int A, int K, int U)
{
    int k, a, u;
    for (k = 0; k < K; k++)
        for (a = 0; a < A; a++) {
            double re = 0.0, im = 0.0;
            for (u = 0; u < U; u++) {
                const oc_complex h = Hzf[(size_t)k * A * U + (size_t)u * A + a], x = Xd[(size_t)u * K + k];
                re += (double)h.real * x.real - (double)h.imag * x.imag;
                im += (double)h.real * x.imag + (double)h.imag * x.real;
            }
            HX[(size_t)a * K + k].real = (float)re;
            HX[(size_t)a * K + k].imag = (float)im;
        }
}
