"""CPU-only tests of the C++ host side that keeps the reference's entry points
(gpu-accel-ofdm-ls-mrc_b200/host): the shared-memory ring's layout/protocol and that the
facade headers and drivers compile against the C ABI with plain g++ (no CUDA headers)."""
import os
import subprocess
import time
import uuid

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "gpu-accel-ofdm-ls-mrc_b200", "host")


@pytest.fixture(scope="module")
def ring_test_bin(tmp_path_factory):
    out = tmp_path_factory.mktemp("ring") / "ring_test"
    subprocess.run(["g++", "-std=c++14", "-O1", "-Wall", "-pthread", "-I", HOST, "-o", str(out),
                    os.path.join(ROOT, "tests", "cpp", "ring_test.cpp"), "-lrt"], check=True)
    return str(out)


def test_ring_selftest(ring_test_bin):
    name = "/lsmrc_test_" + uuid.uuid4().hex[:8]
    r = subprocess.run([ring_test_bin, "selftest", name], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    assert "ring selftest ok" in r.stdout


def test_ring_reads_what_the_reference_producer_writes(ring_test_bin):
    writer = os.path.join(ROOT, "oracle", "_ref", "ring_writer_A4_N64_C16_S16")
    if not os.path.exists(writer):
        pytest.skip("oracle/_ref/ring_writer_* not built (needs /root/reference)")
    count = 40  # more than two laps of the 16-slot ring
    if os.path.exists("/dev/shm/blah"):
        os.unlink("/dev/shm/blah")
    # reader first: it creates the (zeroed) segment and waits for the master to initialise it
    rd = subprocess.Popen([ring_test_bin, "read", "/blah", "4", "64", "16", "16", str(count)],
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    time.sleep(0.05)
    w = subprocess.Popen([writer, str(count)])
    try:
        out, err = rd.communicate(timeout=60)
    finally:
        w.wait(timeout=30)
        if os.path.exists("/dev/shm/blah"):
            os.unlink("/dev/shm/blah")

    class r:  # noqa: N801
        returncode, stdout, stderr = rd.returncode, out, err
    assert r.returncode == 0, r.stderr
    assert f"ring read ok ({count} symbols)" in r.stdout


@pytest.mark.parametrize("threads,ring", [(4, 9), (3, 5), (1, 6)])
def test_feeder_with_several_producer_threads_keeps_order(ring_test_bin, tmp_path, threads, ring):
    """host/ring_feeder --threads T (T slots in flight, published in order) -> ring -> reader: every slot arrives once, in
    order, intact, over several laps of a ring that is shorter than a frame plus the threads in flight"""
    import numpy as np
    assert subprocess.run(["make", "-C", HOST, "--no-print-directory", "bin/ring_feeder"], capture_output=True).returncode == 0
    A, N, C, S, frames, repeat = 3, 64, 16, 7, 2, 5
    rng = np.random.default_rng(11)
    rx = rng.standard_normal((frames, S, A, N + C, 2)).astype(np.float32)
    rx.tofile(tmp_path / "rx.bin")
    name = "/lsmrc_feed_" + uuid.uuid4().hex[:8]
    count = frames * S * repeat
    rd = subprocess.Popen([ring_test_bin, "dump", name, str(A), str(N), str(C), str(ring), str(count), str(tmp_path / "dump.bin")],
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    try:
        time.sleep(0.05)
        fd = subprocess.run([os.path.join(HOST, "bin", "ring_feeder"), "--file", str(tmp_path / "rx.bin"), "--rows", str(A), "--cols", str(N),
                             "--prefix", str(C), "--syms", str(S), "--ring", str(ring), "--shm", name, "--repeat", str(repeat),
                             "--threads", str(threads)], capture_output=True, text=True, timeout=60)
        out, err = rd.communicate(timeout=60)
    finally:
        if rd.poll() is None:
            rd.kill()
        if os.path.exists("/dev/shm" + name):
            os.unlink("/dev/shm" + name)
    assert fd.returncode == 0, fd.stderr
    assert rd.returncode == 0, err
    got = np.fromfile(tmp_path / "dump.bin", np.float32).reshape(repeat, frames, S, A, N + C, 2)
    assert np.array_equal(got, np.broadcast_to(rx, got.shape))


def test_host_programs_build_without_cuda_headers(ofdm):
    ofdm.load_library()  # make sure the .so the programs link exists
    r = subprocess.run(["make", "-C", HOST, "--no-print-directory"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for exe in ("gpuLS_main", "ring_feeder", "stream_main", "rx_and_corr_gpu", "latency_main", "bits_sink"):
        assert os.path.exists(os.path.join(HOST, "bin", exe))
    # the facade keeps the reference's names (gpuLS.cuh:72-113, gpuLS_main.cu:104-141)
    text = open(os.path.join(HOST, "gpuLS.hpp")).read()
    for name in ("class gpuLS", "matrix_readX", "copyPilotToGPU", "shiftOneRowCPU", "ShiftOneRow", "DropPrefix",
                 "FindLeastSquaresGPU", "FindHsqrdforMRC", "MultiplyWithChannelConj", "CombineForMRC", "batchedFFT",
                 "firstVector", "demodOneSymbol", "demodOneFrame", "demodOneFrameCUDA", "demodOptimized", "demodCuBlas"):
        assert name in text, name
    ring = open(os.path.join(HOST, "ShMemSymBuff.hpp")).read()
    for name in ("readNextSymbol", "readLastSymbol", "readNextSymbolCUDA", "readLastSymbolCUDA", "writeNextSymbolWithWait",
                 "writeNextSymbolNoWait", "setBuffLen", "printTimes", "storeTimes", "createStream", "destroyStream"):
        assert name in ring, name


def test_return_ring_carries_frames_of_bits_between_processes(ring_test_bin, ofdm, tmp_path):
    """ShMemBitsBuff: writer (master) -> shared memory -> host/bits_sink (slave), more frames than slots"""
    ofdm.load_library()
    assert subprocess.run(["make", "-C", HOST, "--no-print-directory", "bin/bits_sink"], capture_output=True).returncode == 0
    name = "/lsmrc_bits_" + uuid.uuid4().hex[:8]
    frame_bytes, slots, frames = 3835, 4, 23          # c1-like: 15 rows of 16 bytes would be 240; odd size on purpose
    out = tmp_path / "bits.dat"
    # reader first: it creates the zeroed segment and waits for the master to initialise it
    sink = subprocess.Popen([os.path.join(HOST, "bin", "bits_sink"), "--shm", name, "--frame-bytes", str(frame_bytes),
                             "--slots", str(slots), "--frames", str(frames), "--out", str(out)], stdout=subprocess.PIPE, text=True)
    try:
        time.sleep(0.05)
        w = subprocess.run([ring_test_bin, "bitswrite", name, str(frame_bytes), str(slots), str(frames)],
                           capture_output=True, text=True, timeout=60)
        assert w.returncode == 0, w.stderr
        sink.wait(timeout=60)
    finally:
        if sink.poll() is None:
            sink.kill()
        if os.path.exists("/dev/shm" + name):
            os.unlink("/dev/shm" + name)
    assert sink.returncode == 0
    import numpy as np
    got = np.fromfile(out, np.uint8).reshape(frames, frame_bytes)
    i, j = np.meshgrid(np.arange(frames), np.arange(frame_bytes), indexing="ij")
    assert np.array_equal(got, ((i * 131 + j * 7) & 255).astype(np.uint8))
