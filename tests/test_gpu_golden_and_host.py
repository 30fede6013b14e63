"""More -m gpu parity: the committed golden fixtures (outputs of the reference's own CPU code),
the C++ host programs that keep the reference's entry points (ring feeder -> ring -> gpuLS_main /
stream_main), and size-independent properties at BASELINE.json's full dimensions."""
import glob
import os
import subprocess
import uuid

import numpy as np
import pytest

from util import assert_close

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "gpu-accel-ofdm-ls-mrc_b200", "host")
# (the full-dimension fixtures, which carry a seed instead of the input frame, have their own test in test_gpu_parity.py)
GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")) if "_full_" not in p)


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_cuda_path_matches_reference_golden(ofdm, path):
    g = np.load(path)
    A, N, C, S, b, F = (int(v) for v in g["dims"])
    with ofdm.LsMrcReceiver(A, N, C, S, b, max_frames=2, n_lanes=2) as rx:
        if bool(g["fallback_pilot"]):
            assert rx.set_pilot_file("/nonexistent/Pilots.dat") == 1   # cpuLS.hpp:85-88 fallback
        else:
            rx.set_pilot(g["pilot_asc"])
        got = rx.demod_numpy(g["rx"])
    assert_close(got["hconj"], g["hconj"], "Hconj vs reference build")
    assert_close(got["hsqrd"], g["hsqrd"], "sum|H|^2 vs reference build")
    assert_close(got["combined"], g["combined"], "combined vs reference build")
    assert np.array_equal(got["bits"], g["bits"])


@pytest.fixture(scope="module")
def host_bins(ofdm):
    ofdm.load_library()
    subprocess.run(["make", "-C", HOST, "--no-print-directory"], check=True, stdout=subprocess.DEVNULL)
    return os.path.join(HOST, "bin")


def _run_ring(host_bins, tmp_path, consumer, extra, d, A, N, C, S, b, F, ring, feeder_extra=()):
    shm = "/lsmrc_" + uuid.uuid4().hex[:8]
    rx_file = tmp_path / "rx.bin"
    d["rx"].tofile(rx_file)
    pil = tmp_path / "Pilots.dat"
    d["pilot_asc"].tofile(pil)
    dims = ["--rows", str(A), "--cols", str(N), "--prefix", str(C), "--syms", str(S), "--ring", str(ring), "--shm", shm]
    feeder = subprocess.Popen([os.path.join(host_bins, "ring_feeder"), "--file", str(rx_file), "--frames", str(F)] + dims + list(feeder_extra))
    try:
        r = subprocess.run([os.path.join(host_bins, consumer), "--qam", str(b), "--frames", str(F), "--pilots", str(pil)] + dims + extra,
                           cwd=tmp_path, capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stdout + r.stderr
        feeder.wait(timeout=60)
    finally:
        if feeder.poll() is None:
            feeder.kill()
        if os.path.exists("/dev/shm" + shm):
            os.unlink("/dev/shm" + shm)
    K = N - 1
    comb = np.fromfile(tmp_path / "Output_gpu.dat", np.complex64).reshape(F, S - 1, K)
    bits = np.fromfile(tmp_path / "Bits_gpu.dat", np.uint8).reshape(F, S - 1, -1)
    return comb, bits, r.stdout


@pytest.mark.parametrize("consumer,extra", [("gpuLS_main", []), ("gpuLS_main", ["--frame-mode"]), ("stream_main", []),
                                            ("stream_main", ["--batch", "1", "--lanes", "2"]), ("stream_main", ["--batch", "3", "--lanes", "5"])])
def test_ring_fed_cpp_consumers_match_oracle(ofdm, oracle, host_bins, tmp_path, consumer, extra):
    A, N, C, S, b, F = 4, 64, 16, 16, 2, 7            # config c1 through the ring
    d = ofdm.synth.make_frames(F, A, N, C, S, b, snr_db=10.0, seed=1235)
    ref = oracle.demod_frames(d["rx"], d["pilot_asc"], b, C)
    ring = S + 1 if (consumer == "gpuLS_main" and not extra) else 4 * S + 1
    comb, bits, out = _run_ring(host_bins, tmp_path, consumer, extra, d, A, N, C, S, b, F, ring)
    assert_close(comb, ref["combined"], f"{consumer} combined")
    assert np.array_equal(bits, ref["bits"])
    if consumer == "stream_main":
        # small frames in the pinned ring are read in place by the one-launch kernel (ring wrap included); frames
        # that are already waiting go out in one submission, so there are at most F submissions, all in place
        import re
        m_ = re.search(r"calls=(\d+) in-place-host=(\d+)", out)
        assert m_ and 1 <= int(m_.group(1)) <= F and m_.group(1) == m_.group(2), out
        if "--batch" in extra and extra[extra.index("--batch") + 1] == "1":
            assert int(m_.group(1)) == F, out


def test_stream_main_config3_dims(ofdm, oracle, host_bins, tmp_path):
    A, N, C, S, b, F = 16, 2048, 144, 14, 4, 5        # config c3 shape with fewer antennas
    d = ofdm.synth.make_frames(F, A, N, C, S, b, snr_db=15.0, seed=1237)
    ref = oracle.demod_frames(d["rx"], d["pilot_asc"], b, C)
    comb, bits, out = _run_ring(host_bins, tmp_path, "stream_main", [], d, A, N, C, S, b, F, 4 * S + 1)
    assert_close(comb, ref["combined"], "stream_main combined")
    assert np.array_equal(bits, ref["bits"])
    assert "antenna_samples_per_s" in out


# ---- full BASELINE dimensions: properties that do not need the (slow) oracle --------------------
@pytest.mark.parametrize("name,frames", [("c2", 3), ("c3", 2), ("c4", 1), ("c5", 4)])
def test_full_size_round_trip_and_linearity(ofdm, name, frames):
    import torch

    cfg = ofdm.CONFIGS[name]
    dev = torch.device("cuda:0")
    rx, pilot_asc, src = ofdm.synth.make_frames_torch(frames, cfg, dev, chunk=1)
    K, S = cfg.K, cfg.n_sym
    comb = torch.empty((frames, S - 1, K, 2), device=dev)
    bits = torch.empty((frames, S - 1, cfg.bits_row_bytes), device=dev, dtype=torch.uint8)
    hc = torch.empty((frames, cfg.n_ant, K, 2), device=dev)
    hs = torch.empty((frames, K), device=dev)
    with ofdm.LsMrcReceiver.from_config(cfg) as r:
        r.set_pilot(pilot_asc)
        r.demod_frames_device(torch.view_as_real(rx), frames, comb, bits, hc, hs)
        r.sync()
        # encode -> channel -> decode round trip: decoded bits are the transmitted bits
        want = ofdm.synth.pack_bits_rows(src.cpu().numpy(), cfg.qam_bits)
        assert np.array_equal(bits.cpu().numpy(), want)
        # sum|H|^2 is the energy of the stored estimate
        e = (hc ** 2).sum(dim=(1, 3))
        assert torch.allclose(e, hs, rtol=1e-4)
        # linearity in the data symbols: scaling every data symbol by g scales the combined output by g,
        # the channel estimate (pilot symbol untouched) stays put
        g = 0.5
        rx2 = rx.clone()
        rx2[:, 1:] *= g
        comb2 = torch.empty_like(comb)
        r.demod_frames_device(torch.view_as_real(rx2), frames, comb2, None)
        r.sync()
        assert torch.allclose(comb2, comb * g, rtol=1e-5, atol=1e-6)


def test_standalone_steps_reproduce_the_chain(ofdm, oracle):
    """The individually callable steps (gpuLS.cuh:87-99 wrappers -> lsmrc_stage_*) chained by hand give the
    oracle's Hconj, sum|H|^2 and combined symbols; the batched FFT alone matches numpy."""
    import torch

    A, N, C, S, b = 8, 1024, 64, 4, 4
    K = N - 1
    d = ofdm.synth.make_frames(1, A, N, C, S, b, snr_db=15.0, seed=33)
    ref = oracle.demod_frames(d["rx"], d["pilot_asc"], b, C)
    dev = torch.device("cuda:0")
    rx = torch.view_as_real(torch.from_numpy(d["rx"][0]).to(dev)).contiguous()            # [S][A][N+C][2]
    y = torch.empty((S, A, N, 2), device=dev)
    with ofdm.LsMrcReceiver(A, N, C, S, b) as r:
        r.set_pilot(d["pilot_asc"])
        r.stage("drop_prefix", y, rx, S * A)
        r.sync()
        assert torch.equal(y, rx[:, :, C:, :])
        r.stage("fft", y, S * A)
        r.sync()
        want = np.fft.fft(d["rx"][0][:, :, C:].astype(np.complex128), axis=-1)
        got = torch.view_as_complex(y).cpu().numpy()
        assert np.abs(got - want).max() / np.abs(want).max() < 2e-6
        hconj = torch.empty((A, K, 2), device=dev)
        hsq = torch.empty((K,), device=dev)
        r.stage("find_hs", y[0], hconj, None)
        r.stage("find_hsqrd", hconj, hsq)
        yf = torch.empty((S - 1, A, K, 2), device=dev)
        r.stage("mult_conj", y[1:].contiguous(), hconj, yf, S - 1)
        comb = torch.empty((S - 1, K, 2), device=dev)
        r.stage("combine", yf, hsq, comb, S - 1)
        out = torch.empty_like(comb)
        r.stage("shift_rows", comb, out, S - 1)
        r.sync()
    assert_close(torch.view_as_complex(hconj).cpu().numpy(), ref["hconj"][0], "stage Hconj")
    assert_close(hsq.cpu().numpy(), ref["hsqrd"][0], "stage sum|H|^2")
    assert_close(torch.view_as_complex(out).cpu().numpy(), ref["combined"][0], "stage combined")


def test_decoded_bits_travel_back_over_the_return_ring(ofdm, oracle, host_bins, tmp_path):
    """input ring -> stream_main (GPU) -> return ring (ShMemBitsBuff) -> bits_sink: the downstream process
    receives exactly the oracle's packed bits, frame by frame"""
    A, N, C, S, b, F = 4, 64, 16, 16, 2, 9
    d = ofdm.synth.make_frames(F, A, N, C, S, b, snr_db=10.0, seed=1239)
    ref = oracle.demod_frames(d["rx"], d["pilot_asc"], b, C)
    frame_bytes = ref["bits"].shape[1] * ref["bits"].shape[2]
    name = "/lsmrc_bits_" + uuid.uuid4().hex[:8]
    out = tmp_path / "Bits_ring.dat"
    sink = subprocess.Popen([os.path.join(host_bins, "bits_sink"), "--shm", name, "--frame-bytes", str(frame_bytes), "--slots", "4",
                             "--frames", str(F), "--out", str(out)], stdout=subprocess.PIPE, text=True)
    try:
        comb, bits, _ = _run_ring(host_bins, tmp_path, "stream_main", ["--bits-ring", name, "--bits-slots", "4"], d, A, N, C, S, b, F,
                                  4 * S + 1)
        sink.wait(timeout=60)
    finally:
        if sink.poll() is None:
            sink.kill()
        if os.path.exists("/dev/shm" + name):
            os.unlink("/dev/shm" + name)
    assert sink.returncode == 0
    got = np.fromfile(out, np.uint8).reshape(ref["bits"].shape)
    assert np.array_equal(got, ref["bits"])
    assert np.array_equal(bits, ref["bits"])


def test_latency_harness_reports_every_path(host_bins):
    """host/latency_main (what bench.py's latency leg runs): every path present, the one-launch policy launches one
    kernel per frame and is not slower than the kernel pair"""
    import json

    r = subprocess.run([os.path.join(host_bins, "latency_main"), "--launches", "300", "--warmup", "100"], capture_output=True, text=True,
                       timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    d = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("device_one_launch", "host_one_launch_in_place", "host_one_launch_staged", "device_two_kernels", "host_two_kernels",
                "floor_trivial_kernel"):
        assert d[key]["p50_us"] > 0 and d[key]["p99_us"] >= d[key]["p50_us"], key
    assert d["device_one_launch"]["kernels_per_frame"] == 1 and d["device_two_kernels"]["kernels_per_frame"] == 2
    assert d["device_one_launch"]["p50_us"] < d["device_two_kernels"]["p50_us"]
    assert d["host_one_launch_in_place"]["p50_us"] < d["host_two_kernels"]["p50_us"]


# ---- the reference's OWN GPU driver, compiled unmodified against the drop-in headers -----------------------------
@pytest.mark.parametrize("dims", [(4, 64, 16, 16, 2), (8, 1024, 64, 6, 4)], ids=lambda d: "A%d_N%d_C%d_S%d" % d[:4])
def test_reference_gpu_driver_unmodified_runs_on_the_facade(ofdm, oracle, host_bins, tmp_path, dims):
    """oracle/_ref/gpuLS_main_ref_* is /root/reference/gpuLS_main.cu, byte for byte, built by oracle/Makefile against
    host/gpuLS.cuh + host/ShMemSymBuff_cucomplex.hpp and linked with the C-ABI library.  Fed through the ring by
    ring_feeder it must write the oracle's combined symbols to Output_gpu.dat (gpuLS_main.cu:104-126) and the five
    timers to time_gpu.dat (printTimes / storeTimes, :139-140)."""
    A, N, C, S, b = dims
    exe = os.path.join(ROOT, "oracle", "_ref", "gpuLS_main_ref_A%d_N%d_C%d_S%d" % (A, N, C, S))
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/gpuLS_main_ref_* not built (needs /root/reference and nvcc at build time)")
    K = N - 1
    d = ofdm.synth.make_frames(1, A, N, C, S, b, snr_db=15.0, seed=77)
    ref = oracle.demod_frames(d["rx"], d["pilot_asc"], b, C)
    shm = "/blah"                                     # hard-coded in the reference (shmemID)
    if os.path.exists("/dev/shm" + shm):
        os.unlink("/dev/shm" + shm)
    d["rx"].tofile(tmp_path / "rx.bin")
    d["pilot_asc"].tofile(tmp_path / "Pilots.dat")    # fileNameForX
    dimargs = ["--rows", str(A), "--cols", str(N), "--prefix", str(C), "--syms", str(S), "--ring", str(S), "--shm", shm]
    feeder = subprocess.Popen([os.path.join(host_bins, "ring_feeder"), "--file", str(tmp_path / "rx.bin"), "--frames", "1"] + dimargs)
    try:
        r = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stdout + r.stderr
        feeder.wait(timeout=60)
    finally:
        if feeder.poll() is None:
            feeder.kill()
        if os.path.exists("/dev/shm" + shm):
            os.unlink("/dev/shm" + shm)
    comb = np.fromfile(tmp_path / "Output_gpu.dat", np.complex64).reshape(1, S - 1, K)
    assert_close(comb, ref["combined"], "reference gpuLS_main.cu on the facade: combined")
    bits = np.stack([oracle.demap_row(comb[0, s], b)[0] for s in range(S - 1)])[None]
    assert np.array_equal(bits, ref["bits"])
    times = np.fromfile(tmp_path / "time_gpu.dat", np.float32)
    assert times.shape == (5,) and np.isfinite(times).all() and (times >= 0).all(), times
    assert "ChanEst" in r.stdout                      # printTimes(true) table, ShMemSymBuff_gpu.hpp:153-162


def test_stream_main_full_config3_with_trace(ofdm, oracle, host_bins, tmp_path):
    """BASELINE config c3 at its full shape (128 antennas x 2048 points x 14 symbols, 31 MB per frame) through the ring
    on three lanes, with the per-submission timeline: results equal the oracle's, and the H2D copy of a frame runs
    while an earlier frame is still in the kernels / on its way back (the overlap north_star asks for)."""
    A, N, C, S, b, F = 128, 2048, 144, 14, 4, 9
    d = ofdm.synth.make_frames(F, A, N, C, S, b, snr_db=15.0, seed=1237)
    ref = oracle.demod_frames(d["rx"], d["pilot_asc"], b, C)
    # (the producer fills slots with 8 threads: a single memcpy thread is slower than the GPU side and nothing would overlap)
    comb, bits, out = _run_ring(host_bins, tmp_path, "stream_main", ["--lanes", "3", "--trace", str(tmp_path / "trace.csv")],
                                d, A, N, C, S, b, F, 4 * S + 1, feeder_extra=["--threads", "8"])
    assert_close(comb, ref["combined"], "stream_main combined (full c3)")
    assert np.array_equal(bits, ref["bits"])
    t = np.loadtxt(tmp_path / "trace.csv", delimiter=",", skiprows=1)
    assert t.shape == (F, 7)
    t = t[np.argsort(t[:, 0])]
    enq, h2d, ker, done = t[:, 3], t[:, 4], t[:, 5], t[:, 6]
    assert (enq <= h2d).all() and (h2d <= ker).all() and (ker <= done).all()
    # submission i+1's copy starts before submission i has delivered its results (how often depends on how fast the
    # producer refills the ring on this host; at least once, the pipeline must not be serial by construction)
    overlapped = int((enq[1:] < done[:-1]).sum())
    assert overlapped >= 1, (overlapped, t)


def test_stream_main_two_gpus_one_worker_per_gpu(ofdm, oracle, host_bins, tmp_path):
    """stream_main --gpus 2: one ring, one worker thread and one receiver handle per GPU inside one process
    (SURVEY 8e); each GPU's output equals the oracle on its own frames."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    A, N, C, S, b, F = 8, 1024, 64, 6, 4, 5
    shm = "/lsmrc_" + uuid.uuid4().hex[:8]
    ds, feeders = [], []
    pil = None
    try:
        for g in range(2):
            d = ofdm.synth.make_frames(F, A, N, C, S, b, snr_db=15.0, seed=500 + g, pilot_asc=pil)
            pil = d["pilot_asc"]
            ds.append(d)
            f = tmp_path / f"rx_{g}.bin"
            d["rx"].tofile(f)
            feeders.append(subprocess.Popen([os.path.join(host_bins, "ring_feeder"), "--file", str(f), "--frames", str(F), "--rows", str(A),
                                             "--cols", str(N), "--prefix", str(C), "--syms", str(S), "--ring", str(4 * S + 1),
                                             "--shm", f"{shm}_{g}"]))
        pil.tofile(tmp_path / "Pilots.dat")
        r = subprocess.run([os.path.join(host_bins, "stream_main"), "--gpus", "2", "--qam", str(b), "--frames", str(F), "--pilots",
                            str(tmp_path / "Pilots.dat"), "--rows", str(A), "--cols", str(N), "--prefix", str(C), "--syms", str(S),
                            "--ring", str(4 * S + 1), "--shm", shm], cwd=tmp_path, capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stdout + r.stderr
        for f in feeders:
            f.wait(timeout=60)
    finally:
        for f in feeders:
            if f.poll() is None:
                f.kill()
        for g in range(2):
            if os.path.exists(f"/dev/shm{shm}_{g}"):
                os.unlink(f"/dev/shm{shm}_{g}")
    assert '"gpus": 2' in r.stdout
    for g in range(2):
        ref = oracle.demod_frames(ds[g]["rx"], pil, b, C)
        comb = np.fromfile(tmp_path / f"Output_gpu_{g}.dat", np.complex64).reshape(F, S - 1, N - 1)
        bits = np.fromfile(tmp_path / f"Bits_gpu_{g}.dat", np.uint8).reshape(F, S - 1, -1)
        assert_close(comb, ref["combined"], f"GPU {g} combined")
        assert np.array_equal(bits, ref["bits"])
