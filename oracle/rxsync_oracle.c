/*
 * oracle/rxsync_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the receive front end that sits immediately before the hot path in the
 * reference: the PN-sequence correlator that finds the frame start and
 * the stitching of a frame out of two capture buffers (rx_and_corr.cpp:332-393), plus the
 * per-symbol slot gather of copy_to_shared_mem (rx_and_corr.cpp:64-87).
 *
 * PARITY UNPINNED: rx_and_corr.cpp needs UHD and Boost (absent here) and USRP hardware, so the
 * reference code itself cannot be built or run; this file restates its loops line by line and
 * the tests check it on synthetic captures with known frame positions.
 */
#include <math.h>
#include <string.h>

#include "cpuls_oracle.h"

/* rx_and_corr.cpp:335-360.  Channels are scanned in order, offsets ascending; the scan stops at
 * the first offset whose normalised correlation magnitude reaches `thres`.
 *   temp[i] = sum_j pn[j] * buf[ch][i+j]        (plain complex product, no conjugate)
 *   metric  = |temp[i]| / L
 * Returns the offset (the reference's `length`) or -1; *ch_out / *metric_out describe the hit. */
int oracle_sync_correlate(const oc_complex *buf, int A, int samps, const oc_complex *pn, int L, float thres,
                          int *ch_out, float *metric_out, float *metric_all)
{
    int ch, i, j;
    int found = -1;
    for (ch = 0; ch < A; ch++) {
        for (i = 0; i < samps - L + 1; i++) {
            float re = 0.f, im = 0.f, m;
            const oc_complex *x = buf + (size_t)ch * samps + i;
            for (j = 0; j < L; j++) {
                re = re + (pn[j].real * x[j].real - pn[j].imag * x[j].imag);
                im = im + (pn[j].real * x[j].imag + pn[j].imag * x[j].real);
            }
            m = sqrtf(re * re + im * im) / (float)L;
            if (metric_all) metric_all[(size_t)ch * samps + i] = m;
            if (found < 0 && m >= thres) {
                found = i;
                if (ch_out) *ch_out = ch;
                if (metric_out) *metric_out = m;
                if (!metric_all) return found;
            }
        }
        if (found >= 0 && !metric_all) break;
    }
    return found;
}

/* rx_and_corr.cpp:372-393: the frame starts right after the PN sequence; its tail wraps into the
 * next capture buffer.  copy_buff holds samps - L samples per channel. */
void oracle_sync_assemble(const oc_complex *buf1, const oc_complex *buf2, int A, int samps, int off, int L,
                          oc_complex *copy_buff)
{
    int ch;
    const int n_first = samps - off - L;
    for (ch = 0; ch < A; ch++) {
        oc_complex *dst = copy_buff + (size_t)ch * (samps - L);
        memcpy(dst, buf1 + (size_t)ch * samps + off + L, sizeof(oc_complex) * (size_t)n_first);
        memcpy(dst + n_first, buf2 + (size_t)ch * samps, sizeof(oc_complex) * (size_t)off);
    }
}

/* rx_and_corr.cpp:64-87 (copy_to_shared_mem): symbol i of channel j is
 * copy_buff[j][i*(N+cp) + cp ...] -- the producer strips the CP (keep_cp = 0).  keep_cp = 1 gives
 * the prefix > 0 slot layout instead (the commented-out memcpy at :75), which lets the GPU strip it. */
void oracle_sync_to_slots(const oc_complex *copy_buff, int A, int per_chan, int S, int N, int cp, int keep_cp,
                          oc_complex *slots)
{
    int s, a;
    const int w = N + (keep_cp ? cp : 0);
    for (s = 0; s < S; s++)
        for (a = 0; a < A; a++)
            memcpy(slots + ((size_t)s * A + a) * w, copy_buff + (size_t)a * per_chan + (size_t)s * (N + cp) + (keep_cp ? 0 : cp),
                   sizeof(oc_complex) * (size_t)w);
}
