"""Soft output (max-log LLRs; SURVEY 8f rank 2 -- new, the reference has no demapper at all): the oracle's
definition is self-consistent with the hard demapper, and the CUDA epilogue reproduces it."""
import numpy as np
import pytest

from util import assert_close


@pytest.mark.parametrize("b", [2, 4, 6])
def test_llr_sign_is_the_hard_decision_and_scales_with_snr(oracle, ofdm, b):
    rng = np.random.default_rng(b)
    K = 63
    sym = (1.3 * (rng.standard_normal((1, 2, K)) + 1j * rng.standard_normal((1, 2, K))) / np.sqrt(2)).astype(np.complex64)
    e = (1.0 + rng.random((1, K))).astype(np.float32)
    llr = oracle.soft_demap(sym, e, b, noise_var=0.5)
    idx = np.stack([oracle.demap_row(sym[0, s], b)[1] for s in range(2)])
    hard = (idx[..., None] >> np.arange(b)) & 1
    nz = np.abs(llr[0]) > 1e-6
    assert np.array_equal((llr[0] < 0)[nz], hard.astype(bool)[nz])       # LLR < 0 <=> bit 1
    assert np.allclose(oracle.soft_demap(sym, e, b, noise_var=0.25), 2 * llr, rtol=1e-6)   # 1/noise_var scaling
    assert np.allclose(oracle.soft_demap(sym, 3 * e, b, noise_var=0.5), 3 * llr, rtol=1e-6)  # sum|H|^2 scaling


def test_llr_known_values(oracle):
    a = 1 / np.sqrt(10)
    sym = np.array([[[3 * a - 1j * a, 0.5 * a + 2 * a * 1j, 0]]], np.complex64)   # K = 3
    e = np.ones((1, 3), np.float32)
    llr = oracle.soft_demap(sym, e, 4, noise_var=1.0)[0, 0]
    want0 = 4 * a * np.array([3 * a, -a, 2 * a - 3 * a, 2 * a - a])
    assert np.allclose(llr[0], want0, rtol=1e-6)
    assert np.allclose(llr[2], 4 * a * np.array([0, 0, 2 * a, 2 * a]), rtol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("dims", [(4, 64, 16, 6, 2, 3), (8, 1024, 64, 4, 4, 2), (6, 1024, 64, 3, 6, 2), (5, 2048, 144, 3, 4, 1)])
def test_gpu_llrs_match_oracle(ofdm, oracle, dims):
    import torch

    A, N, C, S, b, F = dims
    K = N - 1
    d = ofdm.synth.make_frames(F, A, N, C, S, b, snr_db={2: 10.0, 4: 15.0, 6: 20.0}[b], seed=17)
    ref = oracle.demod_frames(d["rx"], d["pilot_asc"], b, C)
    noise_var = 0.37
    want = oracle.soft_demap(ref["combined"], ref["hsqrd"], b, noise_var)
    dev = torch.device("cuda:0")
    rx = torch.view_as_real(torch.from_numpy(d["rx"]).to(dev)).contiguous()
    comb = torch.empty((F, S - 1, K, 2), device=dev)
    bits = torch.empty((F, S - 1, (K * b + 7) // 8), device=dev, dtype=torch.uint8)
    llr = torch.full((F, S - 1, K, b), float("nan"), device=dev)
    with ofdm.LsMrcReceiver(A, N, C, S, b) as r:
        r.set_pilot(d["pilot_asc"])
        r.demod_frames_device_soft(rx, F, comb, llr, noise_var, bits)
        r.sync()
        with pytest.raises(ofdm.LsmrcError):
            r.demod_frames_device_soft(rx, F, comb, llr, 0.0, bits)
    got = llr.cpu().numpy()
    assert np.isfinite(got).all()
    assert_close(got, want, "LLRs", tol=2e-5)
    assert np.array_equal(bits.cpu().numpy(), ref["bits"])
    hard = np.unpackbits(ref["bits"], axis=-1, bitorder="little")[..., : K * b].reshape(F, S - 1, K, b).astype(bool)
    big = np.abs(want) > 1e-3 * np.abs(want).max()
    assert np.array_equal((got < 0)[big], hard[big])


@pytest.mark.parametrize("b,snr_db", [(2, 8.0), (4, 16.0), (6, 24.0)])
def test_noise_estimate_recovers_the_noise_level(oracle, ofdm, b, snr_db):
    """decision-directed estimate on synthetic frames: unit-power subcarriers at per-antenna SNR `snr_db` have
    frequency-domain noise variance ~10^(-snr/10) per bin; the estimate is the EFFECTIVE post-combining noise, which
    also carries the LS estimation error of H (one unit-modulus pilot: x2 for a unit-power symbol, plus
    second-order terms that grow at low SNR)"""
    A, N, C, S, F = 8, 256, 16, 9, 2
    d = ofdm.synth.make_frames(F, A, N, C, S, b, snr_db=snr_db, seed=3)
    ref = oracle.demod_frames(d["rx"], d["pilot_asc"], b, C)
    est = oracle.noise_var(ref["combined"], ref["hsqrd"], b)
    true = 10.0 ** (-snr_db / 10.0)
    assert est.shape == (F,)
    assert np.all(est > 1.5 * true) and np.all(est < 3.2 * true), (est, true)


@pytest.mark.gpu
@pytest.mark.parametrize("dims", [(4, 64, 16, 6, 2, 5), (8, 1024, 64, 4, 4, 3), (6, 512, 32, 3, 6, 2)])
def test_gpu_noise_estimate_and_llrs_from_combined(ofdm, oracle, dims):
    import torch

    A, N, C, S, b, F = dims
    K = N - 1
    d = ofdm.synth.make_frames(F, A, N, C, S, b, snr_db={2: 10.0, 4: 15.0, 6: 20.0}[b], seed=29)
    ref = oracle.demod_frames(d["rx"], d["pilot_asc"], b, C)
    want_nv = oracle.noise_var(ref["combined"], ref["hsqrd"], b)
    dev = torch.device("cuda:0")
    rx = torch.view_as_real(torch.from_numpy(d["rx"]).to(dev)).contiguous()
    comb = torch.empty((F, S - 1, K, 2), device=dev)
    hsq = torch.empty((F, K), device=dev)
    nv = torch.zeros(F, device=dev)
    llr = torch.full((F, S - 1, K, b), float("nan"), device=dev)
    with ofdm.LsMrcReceiver(A, N, C, S, b) as r:
        r.set_pilot(d["pilot_asc"])
        r.demod_frames_device(rx, F, comb, None, None, hsq)
        r.estimate_noise_var(comb, hsq, F, nv)
        r.llr_from_combined(comb, hsq, nv, F, llr)
        r.sync()
    got_nv = nv.cpu().numpy()
    assert np.allclose(got_nv, want_nv, rtol=2e-5), (got_nv, want_nv)
    want = np.stack([oracle.soft_demap(ref["combined"][f:f + 1], ref["hsqrd"][f:f + 1], b, float(np.float32(want_nv[f])))[0]
                     for f in range(F)])
    assert_close(llr.cpu().numpy(), want, "LLRs from combined", tol=5e-5)


@pytest.mark.parametrize("b", [2, 4, 6])
def test_llrs_against_brute_force_max_log_over_the_constellation(oracle, ofdm, b):
    """Independent pin of the soft demapper (the reference has none): brute-force max-log LLRs in float64,
        LLR_j = rho * ( min_{s: bit j = 1} |y - s|^2  -  min_{s: bit j = 0} |y - s|^2 ),   rho = sum|H|^2 / noise_var,
    over all 2^b points of the TS 38.211 constellation (ofdm.synth.qam_map_indices, the transmit-side mapper).
    The shipped form is the usual piecewise-linear simplification; where the two coincide they must agree to rounding:
      * QPSK: everywhere;
      * amplitude bits of 16-QAM (bits 2, 3) and the last amplitude bits of 64-QAM (bits 4, 5): everywhere / inside the
        decision regions that do not touch the outermost level;
      * sign bits (0, 1): for |u| <= 2a, i.e. while the nearest point of EITHER sign is an innermost one.
    Outside those regions the simplification keeps the inner segment's slope: same sign, smaller magnitude -- it
    under-states the confidence of symbols far outside the constellation and never flips a decision."""
    rng = np.random.default_rng(100 + b)
    K = 4001
    a = {2: 1 / np.sqrt(2), 4: 1 / np.sqrt(10), 6: 1 / np.sqrt(42)}[b]
    sym = (1.2 * (rng.standard_normal((1, 1, K)) + 1j * rng.standard_normal((1, 1, K))) / np.sqrt(2)).astype(np.complex64)
    e = (0.5 + 4 * rng.random((1, K))).astype(np.float32)
    nv = 0.3
    got = oracle.soft_demap(sym, e, b, noise_var=nv)[0, 0].astype(np.float64)          # [K][b]
    pts = ofdm.synth.qam_map_indices(np.arange(1 << b), b)                              # [2^b]
    y = sym[0, 0].astype(np.complex128)
    d2 = np.abs(y[:, None] - pts[None, :]) ** 2                                         # [K][2^b]
    # sum|H|^2 is indexed by FFT bin, the symbols are in ascending frequency: position i <-> bin index (i + (K-1)/2) mod K
    rho = np.roll(e[0].astype(np.float64), -((K - 1) // 2)) / nv
    exact = np.empty((K, b))
    for j in range(b):
        one = ((np.arange(1 << b) >> j) & 1).astype(bool)
        exact[:, j] = rho * (d2[:, one].min(1) - d2[:, ~one].min(1))
    u = np.stack([y.real, y.imag] * (b // 2), axis=1)                                   # axis value bit j looks at
    au = np.abs(u)
    same = np.zeros((K, b), bool)
    same[:, 0:2] = au[:, 0:2] <= 2 * a if b > 2 else True
    if b == 4:
        same[:, 2:4] = True
    if b == 6:
        same[:, 2:4] = (au[:, 2:4] >= 2 * a) & (au[:, 2:4] <= 6 * a)                    # between the two middle levels' outer neighbours
        same[:, 4:6] = au[:, 4:6] <= 8 * a
    scale = np.abs(exact).max()
    assert same.mean() > 0.5
    assert np.abs(got - exact)[same].max() <= 2e-6 * scale, np.abs(got - exact)[same].max() / scale
    # elsewhere: same decision, never over-confident
    nz = np.abs(exact) > 2e-6 * scale
    assert np.array_equal(np.sign(got[nz]), np.sign(exact[nz]))
    assert (np.abs(got) <= np.abs(exact) + 2e-6 * scale).all()   # (absolute: fp32 cancellation next to a threshold)
