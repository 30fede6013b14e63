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


def test_host_programs_build_without_cuda_headers(ofdm):
    ofdm.load_library()  # make sure the .so the programs link exists
    r = subprocess.run(["make", "-C", HOST, "--no-print-directory"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for exe in ("gpuLS_main", "ring_feeder", "stream_main", "rx_and_corr_gpu"):
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
