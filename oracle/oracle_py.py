"""oracle/oracle_py.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes view of oracle/liboracle.so (the CPU restatement of the reference's
cpuLS.hpp receive path) plus a runner for the reference-built binaries in
oracle/_ref/.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS: dict[str, ctypes.CDLL] = {}


def build(fast: bool = False, ref: bool = True) -> None:
    """Compile the oracle (and, when /root/reference exists, oracle/_ref)."""
    targets = ["liboracle.so", "liboracle_fast.so"]
    subprocess.run(["make", "-C", _HERE, "--no-print-directory", *targets], check=True,
                   stdout=subprocess.DEVNULL)
    if ref and os.path.isdir("/root/reference"):
        subprocess.run(["make", "-C", _HERE, "--no-print-directory", "ref"], check=True,
                       stdout=subprocess.DEVNULL)


def _lib(fast: bool = False) -> ctypes.CDLL:
    name = "liboracle_fast.so" if fast else "liboracle.so"
    if name not in _LIBS:
        path = os.path.join(_HERE, name)
        if not os.path.exists(path):
            build()
        lib = ctypes.CDLL(path)
        lib.oracle_bits_row_bytes.restype = ctypes.c_size_t
        lib.oracle_bits_row_bytes.argtypes = [ctypes.c_int, ctypes.c_int]
        lib.oracle_demod_frames.restype = ctypes.c_int
        lib.oracle_demod_frames.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 6 + \
            [ctypes.c_void_p] * 4 + [ctypes.c_int]
        lib.oracle_fft_f32.restype = None
        lib.oracle_fft_f32.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        lib.oracle_pilot_to_bin_order.restype = None
        lib.oracle_pilot_to_bin_order.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        lib.oracle_shift_one_row.restype = None
        lib.oracle_shift_one_row.argtypes = [ctypes.c_void_p, ctypes.c_int]
        lib.oracle_demap_row.restype = None
        lib.oracle_demap_row.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_void_p, ctypes.c_void_p]
        lib.oracle_soft_demap_row.restype = None
        lib.oracle_soft_demap_row.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                              ctypes.c_void_p]
        lib.oracle_noise_var_frame.restype = ctypes.c_double
        lib.oracle_noise_var_frame.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        lib.oracle_zf_create.restype = ctypes.c_int
        lib.oracle_zf_create.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        lib.oracle_zf_apply.restype = None
        lib.oracle_zf_apply.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        lib.oracle_sync_correlate.restype = ctypes.c_int
        lib.oracle_sync_correlate.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                              ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        lib.oracle_sync_assemble.restype = None
        lib.oracle_sync_assemble.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_void_p]
        lib.oracle_sync_to_slots.restype = None
        lib.oracle_sync_to_slots.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 6 + [ctypes.c_void_p]
        _LIBS[name] = lib
    return _LIBS[name]


def bits_row_bytes(K: int, qam_bits: int) -> int:
    return (K * qam_bits + 7) // 8


def fft_f32(x: np.ndarray, sign: int = -1) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.complex64)
    out = np.empty_like(x)
    _lib().oracle_fft_f32(x.shape[-1], x.ctypes.data, out.ctypes.data, sign)
    return out


def pilot_to_bin_order(p_asc: np.ndarray) -> np.ndarray:
    p_asc = np.ascontiguousarray(p_asc, dtype=np.complex64)
    out = np.empty_like(p_asc)
    _lib().oracle_pilot_to_bin_order(p_asc.ctypes.data, out.ctypes.data, p_asc.shape[0])
    return out


def shift_one_row(row: np.ndarray) -> np.ndarray:
    out = np.array(row, dtype=np.complex64, copy=True, order="C")
    _lib().oracle_shift_one_row(out.ctypes.data, out.shape[0])
    return out


def demap_row(sym: np.ndarray, qam_bits: int):
    sym = np.ascontiguousarray(sym, dtype=np.complex64)
    K = sym.shape[0]
    packed = np.zeros(bits_row_bytes(K, qam_bits), dtype=np.uint8)
    idx = np.zeros(K, dtype=np.uint8)
    _lib().oracle_demap_row(sym.ctypes.data, K, qam_bits, packed.ctypes.data, idx.ctypes.data)
    return packed, idx


def demod_frames(rx: np.ndarray, pilot_asc: np.ndarray, qam_bits: int, cp: int,
                 n_threads: int = 1, fast: bool = False, want_bits: bool = True):
    """rx [F][S][A][N+C] complex64 -> dict(hconj [F,A,K], hsqrd [F,K], combined [F,S-1,K], bits)."""
    rx = np.ascontiguousarray(rx, dtype=np.complex64)
    F, S, A, NC = rx.shape
    N = NC - cp
    K = N - 1
    pilot_asc = np.ascontiguousarray(pilot_asc, dtype=np.complex64)
    assert pilot_asc.shape == (K,)
    hconj = np.empty((F, A, K), np.complex64)
    hsqrd = np.empty((F, K), np.float32)
    comb = np.empty((F, S - 1, K), np.complex64)
    bits = np.zeros((F, S - 1, bits_row_bytes(K, qam_bits)), np.uint8)
    rc = _lib(fast).oracle_demod_frames(rx.ctypes.data, pilot_asc.ctypes.data, F, S, A, N, cp,
                                        qam_bits, hconj.ctypes.data, hsqrd.ctypes.data,
                                        comb.ctypes.data, bits.ctypes.data if want_bits else None,
                                        n_threads)
    if rc != 0:
        raise ValueError(f"oracle_demod_frames rc={rc}")
    return {"hconj": hconj, "hsqrd": hsqrd, "combined": comb, "bits": bits}


def ref_case_name(A: int, N: int, C: int, S: int) -> str:
    return f"A{A}_N{N}_C{C}_S{S}"


def ref_binary(A: int, N: int, C: int, S: int):
    p = os.path.join(_HERE, "_ref", "cpuls_ref_" + ref_case_name(A, N, C, S))
    return p if os.path.exists(p) else None


def run_reference(rx: np.ndarray, pilot_asc, cp: int):
    """Run the reference's own cpuLS.hpp code (oracle/_ref binary) on rx; None if not built."""
    rx = np.ascontiguousarray(rx, dtype=np.complex64)
    F, S, A, NC = rx.shape
    N = NC - cp
    K = N - 1
    exe = ref_binary(A, N, cp, S)
    if exe is None:
        return None
    with tempfile.TemporaryDirectory(prefix="cpuls_ref_") as d:
        rx.tofile(os.path.join(d, "rx.bin"))
        args = [exe, d, os.path.join(d, "rx.bin"), str(F), os.path.join(d, "out")]
        if pilot_asc is not None:
            np.ascontiguousarray(pilot_asc, dtype=np.complex64).tofile(os.path.join(d, "pil.bin"))
            args.append(os.path.join(d, "pil.bin"))
        subprocess.run(args, check=True, timeout=300, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL)
        hconj = np.fromfile(os.path.join(d, "out.hconj"), np.complex64).reshape(F, A, K)
        hsqrd = np.fromfile(os.path.join(d, "out.hsqrd"), np.float32).reshape(F, K)
        comb = np.fromfile(os.path.join(d, "out.comb"), np.complex64).reshape(F, S - 1, K)
    return {"hconj": hconj, "hsqrd": hsqrd, "combined": comb}


# ---- receive front end (rxsync_oracle.c) ------------------------------------------------------------
def sync_correlate(buf: np.ndarray, pn: np.ndarray, thres: float, want_all: bool = False):
    """buf [A][samps] c64, pn [L] c64 -> (offset or -1, channel, metric[, metric_all [A][samps]])"""
    buf = np.ascontiguousarray(buf, np.complex64)
    pn = np.ascontiguousarray(pn, np.complex64)
    A, samps = buf.shape
    ch, m = ctypes.c_int(-1), ctypes.c_float(0)
    allm = np.zeros((A, samps), np.float32) if want_all else None
    off = _lib().oracle_sync_correlate(buf.ctypes.data, A, samps, pn.ctypes.data, pn.shape[0], thres, ctypes.byref(ch),
                                       ctypes.byref(m), allm.ctypes.data if want_all else None)
    return (off, ch.value, m.value, allm) if want_all else (off, ch.value, m.value)


def sync_assemble(buf1: np.ndarray, buf2: np.ndarray, off: int, pn_len: int) -> np.ndarray:
    buf1 = np.ascontiguousarray(buf1, np.complex64)
    buf2 = np.ascontiguousarray(buf2, np.complex64)
    A, samps = buf1.shape
    out = np.empty((A, samps - pn_len), np.complex64)
    _lib().oracle_sync_assemble(buf1.ctypes.data, buf2.ctypes.data, A, samps, off, pn_len, out.ctypes.data)
    return out


def sync_to_slots(copy_buff: np.ndarray, S: int, N: int, cp: int, keep_cp: bool) -> np.ndarray:
    copy_buff = np.ascontiguousarray(copy_buff, np.complex64)
    A, per = copy_buff.shape
    out = np.empty((S, A, N + (cp if keep_cp else 0)), np.complex64)
    _lib().oracle_sync_to_slots(copy_buff.ctypes.data, A, per, S, N, cp, int(keep_cp), out.ctypes.data)
    return out


def time_reference(rx: np.ndarray, pilot_asc, cp: int):
    """Run the reference build on rx and return (frames, seconds of its frame loop) or None if no binary
    exists for these dimensions.  Single process, single thread: that is what the reference is."""
    import json

    rx = np.ascontiguousarray(rx, dtype=np.complex64)
    F, S, A, NC = rx.shape
    exe = ref_binary(A, NC - cp, cp, S)
    if exe is None:
        return None
    d = tempfile.mkdtemp(prefix="cpuls_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        rx.tofile(os.path.join(d, "rx.bin"))
        args = [exe, d, os.path.join(d, "rx.bin"), str(F), os.path.join(d, "out")]
        if pilot_asc is not None:
            np.ascontiguousarray(pilot_asc, dtype=np.complex64).tofile(os.path.join(d, "pil.bin"))
            args.append(os.path.join(d, "pil.bin"))
        r = subprocess.run(args, check=True, timeout=600, capture_output=True, text=True)
        t = json.loads(r.stdout.strip().splitlines()[-1])
        return int(t["frames"]), float(t["seconds"])
    finally:
        import shutil

        shutil.rmtree(d, ignore_errors=True)


def soft_demap(combined: np.ndarray, hsqrd: np.ndarray, qam_bits: int, noise_var: float) -> np.ndarray:
    """combined [F,S-1,K], hsqrd [F,K] -> max-log LLRs [F,S-1,K,b] (LLR > 0 <=> bit 0)"""
    combined = np.ascontiguousarray(combined, np.complex64)
    hsqrd = np.ascontiguousarray(hsqrd, np.float32)
    F, D, K = combined.shape
    out = np.empty((F, D, K, qam_bits), np.float32)
    for f in range(F):
        for s in range(D):
            _lib().oracle_soft_demap_row(combined[f, s].ctypes.data, hsqrd[f].ctypes.data, K, qam_bits, noise_var,
                                         out[f, s].ctypes.data)
    return out


def noise_var(combined: np.ndarray, hsqrd: np.ndarray, qam_bits: int) -> np.ndarray:
    """decision-directed noise-variance estimate per frame: combined [F,S-1,K], hsqrd [F,K] -> [F] float64"""
    combined = np.ascontiguousarray(combined, np.complex64)
    hsqrd = np.ascontiguousarray(hsqrd, np.float32)
    F, D, K = combined.shape
    return np.array([_lib().oracle_noise_var_frame(combined[f].ctypes.data, hsqrd[f].ctypes.data, K, D, qam_bits)
                     for f in range(F)])


# ---- multi-user zero forcing (zf_oracle.c; cpuLS.hpp:400-463) --------------------------------------------
def zf_create(X: np.ndarray):
    """X [U,A,K] complex64 -> (Hzf [K,U,A] complex64: per subcarrier the A x U matrix, column-major; n_singular)"""
    X = np.ascontiguousarray(X, np.complex64)
    U, A, K = X.shape
    H = np.empty((K, U, A), np.complex64)
    bad = _lib().oracle_zf_create(X.ctypes.data, H.ctypes.data, A, K, U)
    return H, bad


def zf_apply(Hzf: np.ndarray, Xd: np.ndarray) -> np.ndarray:
    """Hzf [K,U,A], Xd [U,K] -> HX [A,K]"""
    Hzf = np.ascontiguousarray(Hzf, np.complex64)
    Xd = np.ascontiguousarray(Xd, np.complex64)
    K, U, A = Hzf.shape
    out = np.empty((A, K), np.complex64)
    _lib().oracle_zf_apply(Hzf.ctypes.data, Xd.ctypes.data, out.ctypes.data, A, K, U)
    return out
