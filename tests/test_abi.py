"""The C-ABI library loads without a GPU and exports every symbol include/ofdm_lsmrc.h declares;
compute entry points fail loudly (no CPU fallback) when there is no device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ofdm_lsmrc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lsmrc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(ofdm):
    lib = ofdm.load_library()
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/ofdm_lsmrc.h but not exported"
    assert sorted(ofdm.ABI) == names, "binding.ABI and the header disagree"
    assert lib.lsmrc_abi_version() == 1


def test_geometry_helpers_need_no_gpu(ofdm):
    lib = ofdm.load_library()
    assert lib.lsmrc_bits_row_bytes(1024, 4) == 512
    assert lib.lsmrc_bits_row_bytes(64, 2) == 16
    for n in (64, 128, 256, 512, 1024, 2048, 4096):
        assert lib.lsmrc_supported_fft_size(n) == 1
    assert lib.lsmrc_supported_fft_size(100) == 0 and lib.lsmrc_supported_fft_size(8192) == 0
    assert lib.lsmrc_error_name(-5) == b"LSMRC_ERR_NO_DEVICE"
    cfg = ofdm.pkg.binding.LsmrcConfig(64, 1024, 64, 101, 4, 1, 0, 1)
    assert lib.lsmrc_rx_frame_elems(ctypes.byref(cfg)) == 101 * 64 * 1088


def test_no_cpu_fallback(ofdm):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ofdm.LsmrcError) as e:
        ofdm.LsMrcReceiver(4, 64, 16, 16, 2)
    assert e.value.code == -5  # LSMRC_ERR_NO_DEVICE


def test_product_does_not_link_the_oracle(ofdm):
    import subprocess

    out = subprocess.run(["nm", "-D", "--defined-only", ofdm.build.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle_" not in out and "fftwf_" not in out
    for src in ("csrc/lsmrc_capi.cu", "csrc/lsmrc_kernels.cuh", "binding.py", "host/gpuLS.hpp"):
        text = open(os.path.join(ROOT, "gpu-accel-ofdm-ls-mrc_b200", src)).read()
        assert not re.search(r"#\s*include[^\n]*(oracle|cufft|fftw)", text), src
        assert "liboracle" not in text and "cufftExec" not in text and "oracle_py" not in text, src
    ldd = subprocess.run(["ldd", ofdm.build.LIB_PATH], capture_output=True, text=True).stdout
    assert "cufft" not in ldd and "oracle" not in ldd


def test_configs_match_baseline_table(ofdm):
    c = ofdm.CONFIGS
    assert c["c1"].antenna_samples_per_frame == 5120 and c["c2"].antenna_samples_per_frame == 7032832
    assert c["c3"].antenna_samples_per_frame == 3928064 and c["c4"].antenna_samples_per_frame == 15712256
    assert c["c5"].antenna_samples_per_frame == 20480
    # SURVEY 8d byte formula (+4K for the sum|H|^2 row and per-row padded bit packing)
    assert abs(c["c2"].algorithmic_bytes_per_frame - 54346414) < 5000
    assert abs(c["c4"].algorithmic_bytes_per_frame - 126292879) < 20000


def test_reference_gpu_driver_compiles_unmodified_against_the_drop_in_headers():
    """gpuLS_main.cu of the reference builds, byte for byte, against host/gpuLS.cuh and host/ShMemSymBuff_cucomplex.hpp
    and links with the C-ABI library (recipe: oracle/Makefile, _ref/gpuLS_main_ref_%).  The -m gpu suite runs it."""
    import shutil
    import subprocess

    if not os.path.isdir("/root/reference") or shutil.which("nvcc") is None:
        pytest.skip("needs /root/reference and nvcc (build container)")
    import ofdm_b200

    ofdm_b200.load_library()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "oracle", "_ref", "gpuLS_main_ref_A4_N64_C16_S16")
    if os.path.exists(exe):
        os.unlink(exe)
    r = subprocess.run(["make", "-C", os.path.join(root, "oracle"), "--no-print-directory", "_ref/gpuLS_main_ref_A4_N64_C16_S16"],
                       capture_output=True, text=True)
    assert r.returncode == 0 and os.path.exists(exe), r.stdout + r.stderr
    src = os.path.realpath(os.path.join(root, "oracle", "_ref", "src", "gpuLS_main.cu"))
    assert src == "/root/reference/gpuLS_main.cu"     # a link to the reference's file, not a copy
    nm = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True).stdout
    for sym in ("lsmrc_first_vector", "lsmrc_demod_one_symbol", "lsmrc_set_pilot_file"):
        assert sym in nm, f"the reference driver does not reach the C ABI through {sym}"

