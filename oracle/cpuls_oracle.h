/*
 * oracle/cpuls_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's uplink receive path (cpuLS.hpp) with
 * runtime dimensions.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (libofdm_lsmrc.so) never links or calls it.
 *
 * Conventions: A = RX antennas (numOfRows), N = FFT size (dimension),
 * C = cyclic-prefix length (prefix), S = symbols per frame (lenOfBuffer,
 * symbol 0 = pilot), K = N-1 used subcarriers, b = bits per QAM symbol.
 */
#ifndef CPULS_ORACLE_H
#define CPULS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    float real;
    float imag;
} oc_complex; /* == complexF, ShMemSymBuff.hpp:86-89 */

/* bytes of one packed demap row: ceil(K*b/8) */
size_t oracle_bits_row_bytes(int K, int qam_bits);

/* cpuLS.hpp:105-112 -- roll the ascending-frequency pilot file into FFT-bin order */
void oracle_pilot_to_bin_order(const oc_complex *pilot_asc, oc_complex *x_bin, int K);

/* cpuLS.hpp:135-149 -- shiftOneRow: bin order -> ascending frequency, in place */
void oracle_shift_one_row(oc_complex *row, int K);

/* cpuLS.hpp:247-317 with the pilot read restored (:266-272) and the CP strip of
 * ShMemSymBuff.hpp:281-294.  rx_sym is one ring slot [A][N+C]. */
void oracle_first_vector(const oc_complex *rx_sym, const oc_complex *x_bin, oc_complex *hconj,
                         float *hsqrd, int A, int N, int C);

/* cpuLS.hpp:319-389 (doOneSymbol) minus the file write; out is ascending frequency */
void oracle_one_symbol(const oc_complex *rx_sym, const oc_complex *hconj, const float *hsqrd,
                       oc_complex *out_sorted, int A, int N, int C);

/* hard demap (not in the reference; definition in DESIGN.md / SURVEY.md 8c) */
void oracle_demap_row(const oc_complex *sym, int K, int qam_bits, uint8_t *packed, uint8_t *idx);

/* max-log LLRs of one combined row (new; definition in DESIGN.md): llr [K][b], LLR > 0 <=> bit 0 */
void oracle_soft_demap_row(const oc_complex *sym, const float *hsqrd_bin, int K, int qam_bits, float noise_var, float *llr);

/* decision-directed noise-variance estimate of one frame (new; definition in DESIGN.md) */
double oracle_noise_var_frame(const oc_complex *combined, const float *hsqrd_bin, int K, int n_rows, int qam_bits);

/* whole batch: rx [F][S][A][N+C]; hconj [F][A][K]; hsqrd [F][K]; combined
 * [F][S-1][K]; bits [F][S-1][row_bytes].  n_threads >= 1 splits frames. */
int oracle_demod_frames(const oc_complex *rx, const oc_complex *pilot_asc, int F, int S, int A,
                        int N, int C, int qam_bits, oc_complex *hconj, float *hsqrd,
                        oc_complex *combined, uint8_t *bits, int n_threads);

/* receive front end before the hot path (rxsync_oracle.c; rx_and_corr.cpp:64-87,332-393) */
int oracle_sync_correlate(const oc_complex *buf, int A, int samps, const oc_complex *pn, int L, float thres,
                          int *ch_out, float *metric_out, float *metric_all);
void oracle_sync_assemble(const oc_complex *buf1, const oc_complex *buf2, int A, int samps, int off, int L,
                          oc_complex *copy_buff);
void oracle_sync_to_slots(const oc_complex *copy_buff, int A, int per_chan, int S, int N, int cp, int keep_cp,
                          oc_complex *slots);

/* multi-user zero-forcing helpers (zf_oracle.c; cpuLS.hpp:400-463): X [U][A][K] -> Hzf [K][U][A]; apply -> HX [A][K] */
int oracle_zf_create(const oc_complex *X, oc_complex *Hzf, int A, int K, int U);
void oracle_zf_apply(const oc_complex *Hzf, const oc_complex *Xd, oc_complex *HX, int A, int K, int U);

#ifdef __cplusplus
}
#endif
#endif
