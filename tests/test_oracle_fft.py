"""The oracle's FFT (oracle/fft_shim.c, standing in for the absent FFTW3f) against numpy's
pocketfft and closed-form known answers."""
import numpy as np
import pytest


@pytest.mark.parametrize("n", [64, 128, 256, 512, 1024, 2048, 4096])
def test_shim_fft_matches_numpy(oracle, n):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal((5, n)) + 1j * rng.standard_normal((5, n))).astype(np.complex64)
    want = np.fft.fft(x.astype(np.complex128), axis=-1)
    for row, w in zip(x, want):
        got = oracle.fft_f32(row)
        assert np.abs(got - w).max() / np.abs(w).max() < 2e-6
        back = oracle.fft_f32(got, sign=+1) / n
        assert np.abs(back - row).max() / np.abs(row).max() < 2e-6


@pytest.mark.parametrize("n", [64, 1024, 4096])
def test_shim_fft_known_answers(oracle, n):
    imp = np.zeros(n, np.complex64)
    imp[0] = 1
    assert np.allclose(oracle.fft_f32(imp), np.ones(n), atol=0)          # impulse -> all ones, exactly
    ones = np.ones(n, np.complex64)
    got = oracle.fft_f32(ones)
    assert got[0] == n and np.abs(got[1:]).max() < 1e-3                    # constant -> N at DC
    k0 = 5
    tone = np.exp(2j * np.pi * k0 * np.arange(n) / n).astype(np.complex64)
    got = oracle.fft_f32(tone)                                             # forward kernel exp(-2*pi*i*n*k/N)
    assert abs(got[k0] - n) / n < 1e-5 and np.abs(np.delete(got, k0)).max() / n < 1e-5
