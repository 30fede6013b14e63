"""The receive front end that precedes the hot path (rx_and_corr.cpp:64-87,332-393):
PN frame sync + frame stitching.  CPU: the restated oracle on synthetic captures with a known frame position.
GPU: lsmrc_sync_correlate / lsmrc_sync_assemble against the oracle (bit-exact metric and frame), then the whole
capture -> sync -> assemble -> demod chain on the device against the oracle receiver."""
import numpy as np
import pytest

from util import assert_close


def _capture(ofdm, A=4, N=64, C=16, S=6, b=2, offset=137, extra=50, seed=4):
    d = ofdm.synth.make_frames(1, A, N, C, S, b, snr_db=12.0, seed=seed)
    pn = ofdm.synth.make_pn()
    samps = pn.shape[0] + S * (N + C) + extra
    buf1, buf2 = ofdm.synth.make_capture(d["rx"][0], pn, offset, samps=samps, seed=seed)
    return d, pn, buf1, buf2, samps


def test_pn_is_a_maximal_length_sequence(ofdm):
    pn = ofdm.synth.make_pn().real
    assert pn.shape == (255,) and abs(pn.sum()) == 1          # 128 vs 127 balance
    ac = [np.dot(pn, np.roll(pn, k)) for k in range(1, 255)]
    assert set(ac) == {-1.0}                                   # two-valued autocorrelation


@pytest.mark.parametrize("offset", [0, 1, 137, 400])
def test_oracle_sync_finds_the_frame(oracle, ofdm, offset):
    d, pn, buf1, buf2, samps = _capture(ofdm, offset=offset, extra=420)
    off, ch, metric = oracle.sync_correlate(buf1, pn, 0.5)
    assert (off, ch) == (offset, 0) and 0.9 < metric < 1.1
    frame = oracle.sync_assemble(buf1, buf2, off, pn.shape[0])
    S, A, row = d["rx"][0].shape
    slots = oracle.sync_to_slots(frame, S, 64, 16, keep_cp=True)
    assert np.array_equal(slots, d["rx"][0])                  # the frame comes back sample for sample
    stripped = oracle.sync_to_slots(frame, S, 64, 16, keep_cp=False)  # the producer's own layout (CP removed)
    assert np.array_equal(stripped, d["rx"][0][:, :, 16:])
    assert oracle.sync_correlate(buf1, pn, 5.0)[0] == -1       # threshold never reached


@pytest.mark.gpu
@pytest.mark.parametrize("dims", [(4, 64, 16, 6, 2, 137), (8, 1024, 64, 4, 4, 1000), (3, 256, 0, 5, 6, 0)])
def test_gpu_sync_matches_oracle_and_feeds_the_receiver(ofdm, oracle, dims):
    import torch

    A, N, C, S, b, offset = dims
    d, pn, buf1, buf2, samps = _capture(ofdm, A, N, C, S, b, offset=offset, extra=offset + 33)
    dev = torch.device("cuda:0")
    t1 = torch.view_as_real(torch.from_numpy(buf1).to(dev)).contiguous()
    t2 = torch.view_as_real(torch.from_numpy(buf2).to(dev)).contiguous()
    tpn = torch.view_as_real(torch.from_numpy(pn).to(dev)).contiguous()
    metric_all = torch.zeros((A, samps), device=dev)
    rx = torch.empty((S, A, N + C, 2), device=dev)
    comb = torch.empty((1, S - 1, N - 1, 2), device=dev)
    bits = torch.empty((1, S - 1, (b * (N - 1) + 7) // 8), device=dev, dtype=torch.uint8)
    with ofdm.LsMrcReceiver(A, N, C, S, b) as r:
        r.set_pilot(d["pilot_asc"])
        off, ch, metric = r.sync_correlate(t1, A, samps, tpn, pn.shape[0], 0.5, metric_all)
        o_off, o_ch, o_metric, o_all = oracle.sync_correlate(buf1, pn, 0.5, want_all=True)
        assert (off, ch) == (o_off, o_ch) == (offset, 0)
        assert np.float32(metric) == np.float32(o_metric)                       # bit-exact metric at the hit
        n_off = samps - pn.shape[0] + 1
        assert np.array_equal(metric_all.cpu().numpy()[:, :n_off], o_all[:, :n_off])  # ... and everywhere else
        assert r.sync_correlate(t1, A, samps, tpn, pn.shape[0], 50.0)[0] == -1
        r.sync_assemble(t1, t2, samps, off, pn.shape[0], rx)
        r.sync()
        want = oracle.sync_to_slots(oracle.sync_assemble(buf1, buf2, off, pn.shape[0]), S, N, C, keep_cp=True)
        assert np.array_equal(torch.view_as_complex(rx).cpu().numpy(), want)
        # the stitched frame goes straight into the fused receiver without leaving the device
        r.demod_frames_device(rx, 1, comb, bits)
        r.sync()
    ref = oracle.demod_frames(d["rx"], d["pilot_asc"], b, C)
    assert_close(torch.view_as_complex(comb).cpu().numpy(), ref["combined"], "combined after GPU sync")
    assert np.array_equal(bits.cpu().numpy(), ref["bits"])


@pytest.mark.gpu
def test_rx_and_corr_gpu_program(ofdm, oracle, tmp_path):
    """file-driven C++ front end: capture buffers -> sync -> stitch -> demod, all on the device"""
    import os
    import subprocess

    host = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpu-accel-ofdm-ls-mrc_b200", "host")
    ofdm.load_library()
    subprocess.run(["make", "-C", host, "--no-print-directory"], check=True, stdout=subprocess.DEVNULL)
    A, N, C, S, b, offset = 4, 64, 16, 16, 2, 321
    d, pn, buf1, buf2, samps = _capture(ofdm, A, N, C, S, b, offset=offset, extra=400)
    buf1.tofile(tmp_path / "b1.bin")
    buf2.tofile(tmp_path / "b2.bin")
    pn.tofile(tmp_path / "pn.bin")
    d["pilot_asc"].tofile(tmp_path / "Pilots.dat")
    r = subprocess.run([os.path.join(host, "bin", "rx_and_corr_gpu"), "--buf1", "b1.bin", "--buf2", "b2.bin", "--pn", "pn.bin",
                        "--samps", str(samps), "--rows", str(A), "--cols", str(N), "--prefix", str(C), "--syms", str(S),
                        "--qam", str(b), "--thres", "0.5"], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert f'"offset": {offset}' in r.stdout
    ref = oracle.demod_frames(d["rx"], d["pilot_asc"], b, C)
    comb = np.fromfile(tmp_path / "Output_gpu.dat", np.complex64).reshape(S - 1, N - 1)
    assert_close(comb, ref["combined"][0], "rx_and_corr_gpu combined")
    assert np.array_equal(np.fromfile(tmp_path / "Bits_gpu.dat", np.uint8).reshape(S - 1, -1), ref["bits"][0])
