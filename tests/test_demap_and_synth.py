"""The hard demapper (new code; 3GPP TS 38.211 Gray-mapped square QAM) and the synthetic
generator: map -> demap round trips, bit packing, and a noiseless end-to-end run of the oracle."""
import numpy as np
import pytest


@pytest.mark.parametrize("b", [2, 4, 6])
def test_demap_inverts_the_constellation(oracle, ofdm, b):
    idx = np.arange(1 << b, dtype=np.uint8)
    pts = ofdm.synth.qam_map_indices(idx, b).astype(np.complex64)
    assert abs(np.mean(np.abs(pts) ** 2) - 1.0) < 1e-6          # unit average power
    packed, got = oracle.demap_row(pts, b)
    assert np.array_equal(got, idx)
    assert np.array_equal(packed, ofdm.synth.pack_bits_rows(idx, b))
    # robust to noise well inside the decision regions
    rng = np.random.default_rng(b)
    noisy = (pts + 0.05 * (rng.standard_normal(pts.shape) + 1j * rng.standard_normal(pts.shape)) / np.sqrt({2: 2, 4: 10, 6: 42}[b])).astype(np.complex64)
    assert np.array_equal(oracle.demap_row(noisy, b)[1], idx)


def test_demap_ties_and_packing_layout(oracle):
    z = np.array([0 + 0j, -0.0 - 0.0j], np.complex64)           # +-0 -> bit 0 (strict comparisons)
    assert list(oracle.demap_row(z, 2)[1]) == [0, 0]
    sym = np.array([-1 - 1j, 1 + 1j, -1 + 1j], np.complex64)    # QPSK indices 3, 0, 1
    packed, idx = oracle.demap_row(sym, 2)
    assert list(idx) == [3, 0, 1] and packed[0] == (3 | (0 << 2) | (1 << 4))
    k = 1023
    assert oracle.bits_row_bytes(k, 4) == 512 and oracle.bits_row_bytes(63, 2) == 16 and oracle.bits_row_bytes(4095, 6) == 3072


def test_pilot_roll_and_output_roll_are_inverse_orders(oracle, ofdm):
    K = 63
    p = (np.arange(K) + 1j * np.arange(K)).astype(np.complex64)
    xb = oracle.pilot_to_bin_order(p)
    assert np.array_equal(xb, ofdm.synth.asc_to_bin(p))          # X[k] = P[(k + (K+1)/2) mod K]
    assert np.array_equal(oracle.shift_one_row(xb), p)           # shiftOneRow undoes it
    assert xb[0] == p[(K + 1) // 2]


@pytest.mark.parametrize("dims", [(4, 64, 16, 16, 2), (3, 256, 18, 4, 4), (2, 1024, 64, 3, 6)])
def test_noiseless_frames_decode_to_the_source(oracle, ofdm, dims):
    A, N, C, S, b = dims
    d = ofdm.synth.make_frames(2, A, N, C, S, b, snr_db=None, seed=3)
    out = oracle.demod_frames(d["rx"], d["pilot_asc"], b, C)
    assert np.abs(np.conj(out["hconj"]) - d["h_true"]).max() < 2e-5 * np.abs(d["h_true"]).max() + 1e-5
    want = ofdm.synth.qam_map_indices(d["src_idx"], b)
    assert np.abs(out["combined"] - want).max() < 1e-4
    assert np.array_equal(out["bits"], ofdm.synth.pack_bits_rows(d["src_idx"], b))


def test_identity_channel_passes_symbols_through(oracle, ofdm):
    d = ofdm.synth.make_frames(1, 2, 128, 8, 4, 4, snr_db=None, seed=1, channel="identity")
    out = oracle.demod_frames(d["rx"], d["pilot_asc"], 4, 8)
    assert np.abs(out["hconj"] - 1).max() < 1e-5
    assert np.abs(out["hsqrd"] - 2).max() < 1e-4


def test_oracle_threads_do_not_change_results(oracle, ofdm):
    d = ofdm.synth.make_frames(5, 2, 64, 16, 4, 2, snr_db=10.0, seed=8)
    a = oracle.demod_frames(d["rx"], d["pilot_asc"], 2, 16, n_threads=1)
    b = oracle.demod_frames(d["rx"], d["pilot_asc"], 2, 16, n_threads=3, fast=True)
    assert np.abs(a["combined"] - b["combined"]).max() <= 1e-5 * np.abs(a["combined"]).max()
    assert np.array_equal(a["bits"], b["bits"])
