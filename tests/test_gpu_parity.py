"""GPU parity tests proper: the sm_100a path, called through the C ABI, against the CPU
oracle (restated cpuLS.hpp, pinned to the reference build in test_oracle_vs_ref.py) on the
same seeded inputs.  Tolerances from BASELINE.json north_star: H and combined symbols
within 1e-5 relative (fp32), demapped bits bit-exact."""
import numpy as np
import pytest

from util import REL_TOL, assert_bits_match, assert_close, threshold_margin

pytestmark = pytest.mark.gpu

# (A, N, C, S, b, F) -- the BASELINE configs at sizes the oracle finishes in seconds
CASES = {
    "c1": (4, 64, 16, 16, 2, 3),
    "c2_small": (64, 1024, 64, 6, 4, 3),
    "c3_small": (16, 2048, 144, 4, 4, 2),
    "c4_small": (16, 4096, 288, 3, 6, 2),
    "c5": (16, 64, 16, 16, 2, 1),
    "n128": (3, 128, 8, 4, 2, 5),
    "n256": (8, 256, 32, 6, 4, 3),
    "n512": (5, 512, 0, 3, 6, 2),
    "cp0": (4, 64, 0, 16, 2, 2),
    "odd_ant": (7, 1024, 72, 3, 4, 2),
    "one_ant": (1, 1024, 64, 3, 2, 1),
    "one_data_symbol": (6, 1024, 64, 2, 4, 3),      # S-1 = 1: three of the four teams of a CTA idle
    "odd_cp": (4, 64, 7, 5, 2, 3),                  # rows start 8-byte but not 16-byte aligned
    "odd_cp_1024": (3, 1024, 9, 3, 6, 2),
    "many_frames_small": (2, 128, 4, 3, 4, 300),    # several persistent rounds per CTA
    "syms_not_multiple_of_teams": (4, 1024, 64, 8, 2, 2),  # 7 data symbols, 4 teams per CTA
}
SNR = {2: 10.0, 4: 15.0, 6: 20.0}


def _run_case(ofdm, oracle, A, N, C, S, b, F, seed, max_frames=2, n_lanes=2):
    # a single antenna has no diversity: a Rayleigh deep fade makes y/h ill-conditioned for ANY fp32
    # implementation (the oracle included), so that case uses a unit-modulus channel
    d = ofdm.synth.make_frames(F, A, N, C, S, b, snr_db=SNR[b], seed=seed, channel="unit" if A == 1 else "rayleigh")
    ref = oracle.demod_frames(d["rx"], d["pilot_asc"], b, C)
    with ofdm.LsMrcReceiver(A, N, C, S, b, max_frames=max_frames, n_lanes=n_lanes) as rx:
        rx.set_pilot(d["pilot_asc"])
        got = rx.demod_numpy(d["rx"])
        assert rx.launch_count() > 0
    return d, ref, got


@pytest.mark.parametrize("name", list(CASES))
def test_frames_match_oracle(ofdm, oracle, name):
    A, N, C, S, b, F = CASES[name]
    d, ref, got = _run_case(ofdm, oracle, A, N, C, S, b, F, seed=100 + len(name))
    assert_close(got["hconj"], ref["hconj"], f"{name} Hconj")
    assert_close(got["hsqrd"], ref["hsqrd"], f"{name} sum|H|^2")
    assert_close(got["combined"], ref["combined"], f"{name} combined")
    assert np.array_equal(got["bits"], ref["bits"]), f"{name}: demapped bits differ from the oracle"


def test_bits_recover_source_at_high_snr(ofdm):
    A, N, C, S, b, F = 8, 1024, 64, 4, 6, 2
    d = ofdm.synth.make_frames(F, A, N, C, S, b, snr_db=40.0, seed=7)
    with ofdm.LsMrcReceiver(A, N, C, S, b, max_frames=F) as rx:
        rx.set_pilot(d["pilot_asc"])
        got = rx.demod_numpy(d["rx"])
    want = ofdm.synth.pack_bits_rows(d["src_idx"], b)
    assert np.array_equal(got["bits"], want)
    # noiseless-ish channel estimate equals the true channel
    h_est = np.conj(got["hconj"])
    assert np.abs(h_est - d["h_true"]).max() / np.abs(d["h_true"]).max() < 2e-2


@pytest.mark.parametrize("dims", [(4, 64, 16, 16, 2), (6, 1024, 64, 5, 4), (5, 1024, 9, 3, 6), (4, 2048, 144, 3, 4), (8, 256, 32, 4, 6)])
def test_per_symbol_entry_points_match_frame_path(ofdm, oracle, dims):
    A, N, C, S, b = dims
    d = ofdm.synth.make_frames(1, A, N, C, S, b, snr_db=SNR[b], seed=5)
    ref = oracle.demod_frames(d["rx"], d["pilot_asc"], b, C)
    with ofdm.LsMrcReceiver(A, N, C, S, b) as rx:
        rx.set_pilot(d["pilot_asc"])
        with pytest.raises(ofdm.LsmrcError):
            rx.demod_one_symbol(d["rx"][0, 1])  # data symbol before the pilot symbol
        rx.first_vector(d["rx"][0, 0])
        hc, hs = rx.get_channel()
        assert_close(hc, ref["hconj"][0], "firstVector Hconj")
        assert_close(hs, ref["hsqrd"][0], "firstVector sum|H|^2")
        for s in range(1, S):
            comb, bits = rx.demod_one_symbol(np.ascontiguousarray(d["rx"][0, s]))
            assert_close(comb, ref["combined"][0, s - 1], f"demodOneSymbol {s}")
            assert np.array_equal(bits, ref["bits"][0, s - 1])


def test_chunking_and_lanes_do_not_change_results(ofdm, oracle):
    A, N, C, S, b, F = 4, 256, 16, 5, 4, 11
    d = ofdm.synth.make_frames(F, A, N, C, S, b, snr_db=15.0, seed=11)
    outs = []
    for max_frames, lanes in ((1, 1), (3, 2), (16, 3)):
        with ofdm.LsMrcReceiver(A, N, C, S, b, max_frames=max_frames, n_lanes=lanes) as rx:
            rx.set_pilot(d["pilot_asc"])
            outs.append(rx.demod_numpy(d["rx"]))
    for o in outs[1:]:
        for k in ("combined", "bits", "hconj", "hsqrd"):
            assert np.array_equal(o[k], outs[0][k]), k


def test_errors_are_reported(ofdm):
    with pytest.raises(ofdm.LsmrcError):
        ofdm.LsMrcReceiver(4, 100, 0, 4, 2)  # not a supported FFT size
    with pytest.raises(ofdm.LsmrcError):
        ofdm.LsMrcReceiver(4, 64, 0, 4, 3)  # not a QAM order
    with ofdm.LsMrcReceiver(4, 64, 16, 16, 2) as rx:
        d = ofdm.synth.make_frames(1, 4, 64, 16, 16, 2, seed=1)
        with pytest.raises(ofdm.LsmrcError):
            rx.demod_numpy(d["rx"])  # no pilot yet
        assert rx.set_pilot_file("/nonexistent/Pilots.dat") == 1  # reference fallback 0.707+0.707i


@pytest.mark.gpu
@pytest.mark.parametrize("dims", [(16, 64, 16, 16, 2, 1), (4, 64, 16, 16, 2, 2), (24, 64, 16, 5, 6, 1), (60, 64, 16, 4, 2, 1), (16, 128, 32, 9, 4, 1), (32, 128, 32, 3, 2, 2),
                                  (8, 256, 64, 6, 4, 2), (12, 512, 128, 4, 6, 1), (6, 1024, 64, 5, 4, 1),
                                  (3, 2048, 144, 3, 2, 1), (2, 4096, 288, 3, 6, 1)])
def test_one_launch_mode_matches_oracle_and_two_kernel_path(ofdm, oracle, dims):
    """Launch-latency-bound batches (BASELINE c5: one small frame per call) run as ONE fused kernel; it must
    give the oracle's results and agree with the pilot + data kernel pair on the same input."""
    import torch

    A, N, C, S, b, F = dims
    K = N - 1
    d = ofdm.synth.make_frames(F, A, N, C, S, b, snr_db=18.0, seed=5, channel="unit" if A == 1 else "rayleigh")
    ref = oracle.demod_frames(d["rx"], d["pilot_asc"], b, C)
    dev = torch.device("cuda:0")
    rx = torch.view_as_real(torch.from_numpy(d["rx"]).to(dev)).contiguous()
    out = {}
    for mode in (1, 0):
        comb = torch.zeros((F, S - 1, K, 2), device=dev)
        bits = torch.zeros((F, S - 1, (K * b + 7) // 8), device=dev, dtype=torch.uint8)
        hconj = torch.zeros((F, A, K, 2), device=dev)
        hsq = torch.zeros((F, K), device=dev)
        with ofdm.LsMrcReceiver(A, N, C, S, b) as r:
            r.set_pilot(d["pilot_asc"])
            r.set_oneshot(mode)
            r.demod_frames_device(rx, F, comb, bits, hconj, hsq)
            r.sync()
            assert (r.oneshot_count() > 0) == (mode == 1), (mode, r.describe_plan())
        out[mode] = (torch.view_as_complex(comb).cpu().numpy(), bits.cpu().numpy(), torch.view_as_complex(hconj).cpu().numpy(),
                     hsq.cpu().numpy())
        assert_close(out[mode][2], ref["hconj"], "Hconj", tol=1e-5)
        assert_close(out[mode][3], ref["hsqrd"], "Hsqrd", tol=1e-5)
        assert_close(out[mode][0], ref["combined"], "combined", tol=1e-5)
        if not np.array_equal(out[mode][1], ref["bits"]):
            n_diff = int(np.unpackbits(out[mode][1] ^ ref["bits"]).sum())
            pytest.fail(f"mode {mode}: {n_diff} demapped bits differ (closest oracle symbol is "
                        f"{threshold_margin(ref['combined'], b):.3e} from a threshold)")
    assert_close(out[1][0], out[0][0], "one-launch vs two-kernel combined", tol=2e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("dims", [(16, 64, 16, 16, 2, 1), (4, 64, 16, 16, 2, 3), (8, 256, 32, 5, 6, 1)])
@pytest.mark.parametrize("how", ["host_alloc", "host_register"])
def test_small_pinned_frames_are_processed_in_place(ofdm, oracle, dims, how):
    """lsmrc_demod_frames_host on a small batch in pinned memory: the one-launch kernel reads the samples from
    and writes every result to the host buffers directly (no staging copies); same results as the oracle."""
    A, N, C, S, b, F = dims
    K = N - 1
    d = ofdm.synth.make_frames(F, A, N, C, S, b, snr_db=SNR[b], seed=23)
    ref = oracle.demod_frames(d["rx"], d["pilot_asc"], b, C)
    with ofdm.LsMrcReceiver(A, N, C, S, b, max_frames=F) as r:
        r.set_pilot(d["pilot_asc"])
        shapes = {"rx": (d["rx"].shape, np.complex64), "comb": ((F, S - 1, K), np.complex64),
                  "bits": ((F, S - 1, (K * b + 7) // 8), np.uint8), "hconj": ((F, A, K), np.complex64), "hsq": ((F, K), np.float32)}
        if how == "host_alloc":
            buf = {k: r.pinned_array(*v) for k, v in shapes.items()}
        else:
            buf = {k: np.zeros(*v) for k, v in shapes.items()}
            for v in buf.values():
                r.host_register(v)
        buf["rx"][...] = d["rx"]
        for k in ("comb", "bits", "hconj", "hsq"):
            buf[k][...] = 0
        r.demod_frames_host(buf["rx"], F, buf["comb"], buf["bits"], buf["hconj"], buf["hsq"])
        assert "in-place-host=1" in r.describe_plan(), r.describe_plan()
        got = {k: np.array(v) for k, v in buf.items()}
        # pageable buffers take the staged path and must agree
        out2 = r.demod_numpy(d["rx"])
        if how == "host_register":
            for v in buf.values():
                r.host_unregister(v)
    assert_close(got["hconj"], ref["hconj"], "Hconj")
    assert_close(got["hsq"], ref["hsqrd"], "sum|H|^2")
    assert_close(got["comb"], ref["combined"], "combined")
    assert_bits_match(got["bits"], ref["bits"], ref["combined"], b, f"{dims} in place", got_combined=got["comb"])
    assert_close(out2["combined"], got["comb"], "staged vs in place", tol=2e-6)


def _to_wire_format(rx):
    """what the radio would have sent for these frames: int16 I/Q at 12 dB below full scale (clipped), and the complex
    float samples UHD's host-side converter makes of them (fc32 = sc16 * 1/32767, rx_and_corr.cpp:283)"""
    scale = np.float32(1.0 / 32767.0)
    peak = np.abs(rx.view(np.float32)).max()
    iq = np.clip(np.rint(rx.view(np.float32).reshape(rx.shape + (2,)) * (8191.0 / peak)), -32768, 32767).astype(np.int16)
    as_float = (iq.astype(np.float32) * scale).reshape(rx.shape + (2,)).view(np.complex64)[..., 0]
    return iq, scale, np.ascontiguousarray(as_float)


@pytest.mark.gpu
@pytest.mark.parametrize("dims", [(4, 64, 16, 16, 2, 5), (8, 256, 32, 5, 6, 3), (16, 1024, 64, 6, 4, 7), (6, 1024, 9, 3, 6, 2),
                                  (8, 2048, 144, 4, 4, 3), (5, 4096, 288, 3, 6, 2)])
def test_wire_format_ingest_equals_host_conversion(ofdm, oracle, dims):
    """lsmrc_demod_frames_host_sc16: int16 I/Q over the link, converted on the device.  The oracle runs on the floats UHD's
    host conversion gives for the same int16 samples (1e-5 / bits exact); and because int16 -> float * scale is exact
    up to one rounding, the complex-float entry point fed with those floats must agree BIT FOR BIT."""
    A, N, C, S, b, F = dims
    d = ofdm.synth.make_frames(F, A, N, C, S, b, snr_db=SNR[b], seed=31)
    iq, scale, rx_f = _to_wire_format(d["rx"])
    ref = oracle.demod_frames(rx_f, d["pilot_asc"], b, C)
    K = N - 1
    with ofdm.LsMrcReceiver(A, N, C, S, b, max_frames=2) as r:       # several chunks over the lanes
        r.set_pilot(d["pilot_asc"])
        r.set_oneshot(0)
        comb = np.empty((F, S - 1, K), np.complex64)
        bits = np.empty((F, S - 1, (K * b + 7) // 8), np.uint8)
        hc = np.empty((F, A, K), np.complex64)
        hs = np.empty((F, K), np.float32)
        r.demod_frames_host_sc16(iq, F, scale, comb, bits, hc, hs)
        same = r.demod_numpy(rx_f)
    assert_close(hc, ref["hconj"], "Hconj")
    assert_close(hs, ref["hsqrd"], "sum|H|^2")
    assert_close(comb, ref["combined"], "combined")
    assert_bits_match(bits, ref["bits"], ref["combined"], b, f"{dims} sc16", got_combined=comb)
    assert np.array_equal(comb.view(np.uint32), same["combined"].view(np.uint32))
    assert np.array_equal(hc.view(np.uint32), same["hconj"].view(np.uint32))
    assert np.array_equal(bits, same["bits"])


@pytest.mark.gpu
def test_wire_format_conversion_alone(ofdm):
    """lsmrc_sc16_to_fc32_device against numpy for even and odd row geometries (vector and scalar kernels)"""
    import torch

    dev = torch.device("cuda:0")
    rng = np.random.default_rng(5)
    with ofdm.LsMrcReceiver(4, 64, 16, 4, 2) as r:
        for rows, n_in, skip, n_out in [(7, 80, 16, 64), (5, 73, 9, 64), (3, 1088, 64, 1024), (2, 65, 0, 65), (4, 80, 0, 0)]:
            iq = rng.integers(-32768, 32768, size=(rows, n_in, 2), dtype=np.int16)
            scale = np.float32(1.0 / 32767.0)
            d_in = torch.from_numpy(iq).to(dev)
            d_out = torch.full((rows, max(n_out, 1), 2), -7.0, device=dev)
            torch.cuda.synchronize()   # the copy and the fill run on torch's stream, the receiver on its own
            r.sc16_to_fc32_device(d_in, rows, n_in, skip, n_out, scale, d_out)
            r.sync()
            got = d_out.cpu().numpy()
            if n_out == 0:
                assert (got == -7.0).all()       # nothing to convert, nothing written
                continue
            want = iq[:, skip:skip + n_out].astype(np.float32) * scale
            assert np.array_equal(got, want), (rows, n_in, skip, n_out)


@pytest.mark.gpu
@pytest.mark.parametrize("dims", [(16, 1024, 64, 9, 4, 40), (4, 64, 16, 16, 2, 300), (8, 2048, 144, 4, 6, 12)])
def test_results_are_bit_repeatable(ofdm, dims):
    """fixed reduction orders everywhere (no float atomics; tickets only decide WHO computes an item): the same
    input gives bit-identical channel estimates, combined symbols and bits on every call"""
    import torch

    A, N, C, S, b, F = dims
    K = N - 1
    dev = torch.device("cuda:0")
    rx = torch.randn((F, S, A, N + C, 2), device=dev)
    outs = []
    with ofdm.LsMrcReceiver(A, N, C, S, b) as r:
        r.set_pilot(ofdm.synth.make_pilot(K, 1))
        for _ in range(4):
            comb = torch.empty((F, S - 1, K, 2), device=dev)
            bits = torch.empty((F, S - 1, (K * b + 7) // 8), device=dev, dtype=torch.uint8)
            hc = torch.empty((F, A, K, 2), device=dev)
            hs = torch.empty((F, K), device=dev)
            r.demod_frames_device(rx, F, comb, bits, hc, hs)
            r.sync()
            outs.append((comb, bits, hc, hs))
    for o in outs[1:]:
        for a_, b_ in zip(outs[0], o):
            assert torch.equal(a_, b_)


@pytest.mark.gpu
@pytest.mark.parametrize("dims", [(9, 1024, 64, 11, 4, 64), (5, 1024, 32, 6, 6, 200), (12, 512, 32, 7, 2, 96), (6, 2048, 144, 5, 4, 40),
                                  # enough (frame, symbol) pairs that 2048/4096 points run their dedicated data kernel (ring of
                                  # channel-row chunks, bulk-copied sample rows), with odd symbol counts and an odd prefix
                                  (3, 2048, 144, 10, 4, 70), (5, 2048, 7, 4, 6, 161), (3, 4096, 288, 6, 6, 60), (2, 4096, 33, 3, 2, 130)])
def test_many_frames_through_the_persistent_kernels(ofdm, oracle, dims):
    """batches large enough that every persistent CTA takes several work items (ticket loop, Hconj ring and bulk-copy
    barriers wrap many times), with symbol counts that leave teams idle in the last group of a frame"""
    import torch

    A, N, C, S, b, F = dims
    K = N - 1
    d = ofdm.synth.make_frames(F, A, N, C, S, b, snr_db=SNR[b], seed=77)
    ref = oracle.demod_frames(d["rx"], d["pilot_asc"], b, C)
    dev = torch.device("cuda:0")
    rx = torch.view_as_real(torch.from_numpy(d["rx"]).to(dev)).contiguous()
    comb = torch.zeros((F, S - 1, K, 2), device=dev)
    bits = torch.zeros((F, S - 1, (K * b + 7) // 8), device=dev, dtype=torch.uint8)
    hc = torch.zeros((F, A, K, 2), device=dev)
    hs = torch.zeros((F, K), device=dev)
    with ofdm.LsMrcReceiver(A, N, C, S, b) as r:
        r.set_pilot(d["pilot_asc"])
        r.set_oneshot(0)     # the kernel pair, whatever the batch size
        for _ in range(2):   # second call: ticket base and barrier phases carry over from the first
            r.demod_frames_device(rx, F, comb, bits, hc, hs)
        r.sync()
        assert r.oneshot_count() == 0
    assert_close(torch.view_as_complex(hc).cpu().numpy(), ref["hconj"], "Hconj")
    assert_close(hs.cpu().numpy(), ref["hsqrd"], "sum|H|^2")
    assert_close(torch.view_as_complex(comb).cpu().numpy(), ref["combined"], "combined")
    got_bits = bits.cpu().numpy()
    if not np.array_equal(got_bits, ref["bits"]):
        n_diff = int(np.unpackbits(got_bits ^ ref["bits"]).sum())
        pytest.fail(f"{n_diff} demapped bits differ (closest oracle symbol is {threshold_margin(ref['combined'], b):.3e} from a threshold)")


# (A, N, C, S, b, F): at least one full wave of data items (444 CTAs x teams), so the 2048/4096-point plans take the
# single launch; the third case has several antenna groups per frame, the fourth an odd prefix (no bulk copies)
ONE_LAUNCH = [(2, 4096, 288, 4, 6, 160), (3, 2048, 144, 8, 4, 140), (32, 4096, 288, 4, 6, 150), (2, 4096, 33, 3, 2, 230),
              (4, 2048, 144, 14, 4, 80)]


@pytest.mark.gpu
@pytest.mark.parametrize("dims", ONE_LAUNCH)
def test_single_launch_frames_equal_the_kernel_pair(ofdm, oracle, dims):
    """2048/4096 points, large batches: pilot and data items come from ONE persistent launch (lsmrc_frames_sh; data items
    wait on per-frame ready flags written by the pilot items).  Same arithmetic in the same order as the kernel pair, so
    the outputs must be bit-identical to it -- and match the oracle; repeated calls reuse the device-side launch number."""
    import torch

    A, N, C, S, b, F = dims
    K = N - 1
    d = ofdm.synth.make_frames(F, A, N, C, S, b, snr_db=SNR[b], seed=91)
    dev = torch.device("cuda:0")
    rx = torch.view_as_real(torch.from_numpy(d["rx"]).to(dev)).contiguous()

    def run(r, want_h):
        comb = torch.zeros((F, S - 1, K, 2), device=dev)
        bits = torch.zeros((F, S - 1, (K * b + 7) // 8), device=dev, dtype=torch.uint8)
        hc = torch.zeros((F, A, K, 2), device=dev) if want_h else None
        hs = torch.zeros((F, K), device=dev) if want_h else None
        torch.cuda.synchronize()   # the fills run on torch's stream, the receiver on its own
        r.demod_frames_device(rx, F, comb, bits, hc, hs)
        r.sync()
        return comb, bits, hc, hs

    with ofdm.LsMrcReceiver(A, N, C, S, b) as r:
        r.set_pilot(d["pilot_asc"])
        one = [run(r, True), run(r, False), run(r, True)]          # three launches: the launch number advances on the device
        assert r.one_launch_frames_count() == 3
        r.set_one_launch_frames(False)
        pair = run(r, True)
        assert r.one_launch_frames_count() == 3
    for got in one:
        for x, y in zip(got, pair):
            if x is not None:
                assert torch.equal(x, y)
    ref = oracle.demod_frames(d["rx"], d["pilot_asc"], b, C)
    comb, bits, hc, hs = one[0]
    assert_close(torch.view_as_complex(hc).cpu().numpy(), ref["hconj"], "Hconj")
    assert_close(hs.cpu().numpy(), ref["hsqrd"], "sum|H|^2")
    assert_close(torch.view_as_complex(comb).cpu().numpy(), ref["combined"], "combined")
    assert_bits_match(bits.cpu().numpy(), ref["bits"], ref["combined"], b, f"{dims} one launch",
                      got_combined=torch.view_as_complex(comb).cpu().numpy())


# ---- BASELINE configs c2, c3, c4 at their FULL dimensions (cpuLS_main.cpp:80-93 is the loop to match) ---------------
FULL = {
    "c2_full": "c2_full_A64_N1024_C64_S101_16qam",
    "c3_full": "c3_full_A128_N2048_C144_S14_16qam",
    "c4_full": "c4_full_A256_N4096_C288_S14_64qam",
}


@pytest.mark.gpu
@pytest.mark.parametrize("path_kind", ["device", "host"])
@pytest.mark.parametrize("name", list(FULL))
def test_full_baseline_dimensions_match_oracle_and_reference_outputs(ofdm, oracle, name, path_kind):
    """One whole frame of each named configuration (64 x 1024 x 101 symbols, 128 x 2048 x 14, 256 x 4096 x 14) through
    lsmrc_demod_frames_device and lsmrc_demod_frames_host: H, sum|H|^2 and combined symbols within 1e-5 (both norms)
    of the oracle, bits identical; additionally against the outputs of the REFERENCE's own build committed in
    tests/golden (which the oracle equals bit for bit, tests/test_oracle_vs_ref.py)."""
    import os

    import torch

    from util import HCONJ_SAMPLE_STRIDE, load_golden

    g = load_golden(os.path.join(os.path.dirname(__file__), "golden", FULL[name] + ".npz"), ofdm)
    A, N, C, S, b, F = (int(v) for v in g["dims"])
    K = N - 1
    ref = oracle.demod_frames(g["rx"], g["pilot_asc"], b, C)
    with ofdm.LsMrcReceiver(A, N, C, S, b, max_frames=1, n_lanes=2) as r:
        r.set_pilot(g["pilot_asc"])
        r.set_oneshot(0)
        if path_kind == "host":
            got = r.demod_numpy(g["rx"])
        else:
            dev = torch.device("cuda:0")
            rx = torch.view_as_real(torch.from_numpy(g["rx"]).to(dev)).contiguous()
            comb = torch.zeros((F, S - 1, K, 2), device=dev)
            bits = torch.zeros((F, S - 1, (K * b + 7) // 8), device=dev, dtype=torch.uint8)
            hc = torch.zeros((F, A, K, 2), device=dev)
            hs = torch.zeros((F, K), device=dev)
            r.demod_frames_device(rx, F, comb, bits, hc, hs)
            r.sync()
            got = {"combined": torch.view_as_complex(comb).cpu().numpy(), "bits": bits.cpu().numpy(),
                   "hconj": torch.view_as_complex(hc).cpu().numpy(), "hsqrd": hs.cpu().numpy()}
        assert r.launch_count() >= 2
    assert_close(got["hconj"], ref["hconj"], f"{name} Hconj")
    assert_close(got["hsqrd"], ref["hsqrd"], f"{name} sum|H|^2")
    assert_close(got["combined"], ref["combined"], f"{name} combined")
    assert_bits_match(got["bits"], ref["bits"], ref["combined"], b, name, got_combined=got["combined"])
    # the reference build's own outputs
    assert_close(got["combined"], g["combined"], f"{name} combined vs reference build")
    assert_close(got["hsqrd"], g["hsqrd"], f"{name} sum|H|^2 vs reference build")
    assert_close(np.ascontiguousarray(got["hconj"].ravel()[::HCONJ_SAMPLE_STRIDE]), g["hconj_sample"], f"{name} Hconj sample vs reference build")
    assert_bits_match(got["bits"], g["bits"], g["combined"], b, name + " vs reference build", got_combined=got["combined"])
    # and the frame decodes to what was transmitted (20 dB / 15 dB with 64+ antennas: error free)
    assert np.array_equal(got["bits"], ofdm.synth.pack_bits_rows(g["src_idx"], b))


@pytest.mark.gpu
def test_caller_stream_orders_the_launches_and_null_returns_to_the_own_stream(ofdm, oracle):
    """lsmrc_set_stream: device-resident calls run on the caller's stream (work enqueued on it before and after is
    ordered with the kernels); NULL switches back to the handle's own stream (include/ofdm_lsmrc.h)."""
    import torch

    A, N, C, S, b, F = 6, 1024, 64, 5, 4, 24
    K = N - 1
    d = ofdm.synth.make_frames(F, A, N, C, S, b, snr_db=15.0, seed=31)
    ref = oracle.demod_frames(d["rx"], d["pilot_asc"], b, C)
    dev = torch.device("cuda:0")
    user = torch.cuda.Stream(dev)
    host_rx = torch.view_as_real(torch.from_numpy(d["rx"])).contiguous().pin_memory()
    with ofdm.LsMrcReceiver(A, N, C, S, b) as r:
        r.set_pilot(d["pilot_asc"])
        r.set_oneshot(0)
        for use_user in (True, False):
            rx = torch.empty_like(host_rx, device=dev)
            comb = torch.zeros((F, S - 1, K, 2), device=dev)
            bits = torch.zeros((F, S - 1, (K * b + 7) // 8), device=dev, dtype=torch.uint8)
            if use_user:
                r.set_stream(user.cuda_stream)
                with torch.cuda.stream(user):
                    rx.copy_(host_rx, non_blocking=True)        # the kernels must wait for this copy ...
                    r.demod_frames_device(rx, F, comb, bits)
                    out = comb.to("cpu", non_blocking=True)     # ... and this copy for the kernels
                user.synchronize()
            else:
                r.set_stream(None)                              # back to the handle's own stream
                rx.copy_(host_rx)
                torch.cuda.synchronize(dev)
                r.demod_frames_device(rx, F, comb, bits)
                r.sync()
                out = comb.cpu()
            assert_close(torch.view_as_complex(out).numpy(), ref["combined"], f"combined (user stream: {use_user})")
            assert np.array_equal(bits.cpu().numpy(), ref["bits"])


@pytest.mark.gpu
@pytest.mark.parametrize("dims", [(6, 1024, 64, 5, 4, 140), (2, 4096, 288, 4, 6, 160), (4, 64, 16, 16, 2, 300)])
def test_device_calls_replay_from_a_cuda_graph(ofdm, dims):
    """The persistent kernels keep their work counters (and the single-launch kernel its launch number and ready flags) on the
    device and re-arm them themselves, so a captured call can be replayed: capture lsmrc_demod_frames_device on the caller's
    stream once, replay it on three different inputs, and compare each replay with an ordinary call on the same input."""
    import torch

    A, N, C, S, b, F = dims
    K = N - 1
    dev = torch.device("cuda:0")
    user = torch.cuda.Stream(dev)
    g = torch.Generator(device=dev).manual_seed(5)
    inputs = [torch.randn((F, S, A, N + C, 2), device=dev, generator=g) for _ in range(3)]
    rx = torch.empty_like(inputs[0])
    comb = torch.zeros((F, S - 1, K, 2), device=dev)
    bits = torch.zeros((F, S - 1, (K * b + 7) // 8), device=dev, dtype=torch.uint8)
    hs = torch.zeros((F, K), device=dev)
    torch.cuda.synchronize(dev)
    with ofdm.LsMrcReceiver(A, N, C, S, b) as r:
        r.set_pilot(ofdm.synth.make_pilot(K, 3))
        r.set_oneshot(0)
        r.set_stream(user.cuda_stream)
        with torch.cuda.stream(user):
            rx.copy_(inputs[0])
            r.demod_frames_device(rx, F, comb, bits, None, hs)        # first call: channel state allocated outside the capture
        user.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=user):
            r.demod_frames_device(rx, F, comb, bits, None, hs)
        for x in inputs:
            with torch.cuda.stream(user):
                rx.copy_(x)
                comb.zero_()
                bits.zero_()
                hs.zero_()
                graph.replay()
            user.synchronize()
            got = (comb.clone(), bits.clone(), hs.clone())
            with torch.cuda.stream(user):
                comb.zero_()
                bits.zero_()
                hs.zero_()
                r.demod_frames_device(rx, F, comb, bits, None, hs)
            user.synchronize()
            assert torch.equal(got[0], comb) and torch.equal(got[1], bits) and torch.equal(got[2], hs)
            assert bool(torch.isfinite(comb).all()) and float(comb.abs().max()) > 0
        r.set_stream(None)


@pytest.mark.gpu
def test_two_receivers_on_one_gpu_from_two_threads(ofdm):
    """One handle per thread is the contract (a handle is single-threaded).  Two receivers -- a 4096-point one on the
    single-launch kernel, whose data items wait for pilot items of the same launch, and a 1024-point one on the kernel pair --
    run concurrently on one GPU from two host threads; both kinds of persistent kernel then share the SMs, neither may stall
    the other for good, and every result must equal the one computed alone."""
    import threading

    import torch

    dev = torch.device("cuda:0")
    cases = [(2, 4096, 288, 4, 6, 160), (6, 1024, 64, 5, 4, 140)]
    state = []
    for A, N, C, S, b, F in cases:
        K = N - 1
        g = torch.Generator(device=dev).manual_seed(N)
        rx = torch.randn((F, S, A, N + C, 2), device=dev, generator=g)
        comb = torch.zeros((F, S - 1, K, 2), device=dev)
        bits = torch.zeros((F, S - 1, (K * b + 7) // 8), device=dev, dtype=torch.uint8)
        r = ofdm.LsMrcReceiver(A, N, C, S, b)
        r.set_pilot(ofdm.synth.make_pilot(K, 2))
        r.set_oneshot(0)
        state.append((r, rx, F, comb, bits))
    torch.cuda.synchronize(dev)
    try:
        alone = []
        for r, rx, F, comb, bits in state:
            r.demod_frames_device(rx, F, comb, bits)
            r.sync()
            alone.append((comb.clone(), bits.clone()))
        errors = []

        def work(i):
            r, rx, F, comb, bits = state[i]
            try:
                for _ in range(12):
                    r.demod_frames_device(rx, F, comb, bits)
                r.sync()
            except Exception as e:  # noqa: BLE001
                errors.append(e)

        threads = [threading.Thread(target=work, args=(i,)) for i in range(len(state))]
        for t in threads:
            t.start()
        for t in threads:
            t.join(timeout=120)
        assert not any(t.is_alive() for t in threads), "a receiver did not finish: the two persistent kernels block each other"
        assert not errors, errors
        for (r, rx, F, comb, bits), (c0, b0) in zip(state, alone):
            assert torch.equal(comb, c0) and torch.equal(bits, b0)
        assert state[0][0].one_launch_frames_count() == 13
    finally:
        for r, *_ in state:
            r.close()
