"""Synthetic uplink frames (the reference ships no data and no generator).

Per frame: seeded QAM source bits -> Gray-mapped unit-power constellation on bins
1..N-1 (bin 0 = DC left empty, as the receiver drops it: cpuLS.hpp:292,355), a fixed
unit-modulus QPSK pilot as symbol 0, block-fading i.i.d. Rayleigh channel per antenna
and subcarrier, IFFT, cyclic prefix, AWGN at the configured per-antenna SNR.
Layout [F][S][A][N+C] complex64 -- ring-slot order (ShMemSymBuff.hpp:92-106).

numpy version for the CPU-sized parity tests, torch version to build full-size
batches directly in HBM.  Used to feed the receiver; not part of the receive path.
"""
from __future__ import annotations

import numpy as np

_QAM_SCALE = {2: np.sqrt(2.0), 4: np.sqrt(10.0), 6: np.sqrt(42.0)}


def qam_map_indices(idx: np.ndarray, qam_bits: int) -> np.ndarray:
    """3GPP TS 38.211 5.1.3-5.1.5 mapping; idx holds b bits, bit j = b_j (LSB first)."""
    idx = np.asarray(idx).astype(np.int64)
    b = [(idx >> j) & 1 for j in range(qam_bits)]
    if qam_bits == 2:
        re = 1 - 2 * b[0]
        im = 1 - 2 * b[1]
    elif qam_bits == 4:
        re = (1 - 2 * b[0]) * (2 - (1 - 2 * b[2]))
        im = (1 - 2 * b[1]) * (2 - (1 - 2 * b[3]))
    elif qam_bits == 6:
        re = (1 - 2 * b[0]) * (4 - (1 - 2 * b[2]) * (2 - (1 - 2 * b[4])))
        im = (1 - 2 * b[1]) * (4 - (1 - 2 * b[3]) * (2 - (1 - 2 * b[5])))
    else:
        raise ValueError("qam_bits must be 2, 4 or 6")
    return (re + 1j * im) / _QAM_SCALE[qam_bits]


def make_pilot(K: int, seed: int) -> np.ndarray:
    """Unit-modulus QPSK pilot, ascending-frequency order (the Pilots.dat order)."""
    rng = np.random.default_rng(seed ^ 0x5EED)
    return np.exp(1j * (np.pi / 4) * (2 * rng.integers(0, 4, K) + 1)).astype(np.complex64)


def asc_to_bin(v_asc: np.ndarray) -> np.ndarray:
    """ascending frequency [bin N/2..N-1, 1..N/2-1] -> FFT-bin order [1..N-1] (last axis)."""
    K = v_asc.shape[-1]
    return np.roll(v_asc, -((K + 1) // 2), axis=-1)


def bin_to_asc(v_bin: np.ndarray) -> np.ndarray:
    K = v_bin.shape[-1]
    return np.roll(v_bin, (K + 1) // 2, axis=-1)


def pack_bits_rows(idx: np.ndarray, qam_bits: int) -> np.ndarray:
    """idx [..., K] symbol indices -> packed rows [..., ceil(K*b/8)], LSB first."""
    idx = np.asarray(idx, dtype=np.uint8)
    K = idx.shape[-1]
    bits = ((idx[..., None] >> np.arange(qam_bits, dtype=np.uint8)) & 1).reshape(*idx.shape[:-1], K * qam_bits)
    return np.packbits(bits, axis=-1, bitorder="little")


def make_frames(n_frames, n_ant, fft_size, cp_len, n_sym, qam_bits, snr_db=None, seed=0,
                channel="rayleigh", pilot_asc=None):
    """Returns dict(rx [F,S,A,N+C] c64, pilot_asc [K] c64, src_idx [F,S-1,K] u8 (ascending
    frequency), h_true [F,A,K] c128 (bin order))."""
    F, A, N, C, S = n_frames, n_ant, fft_size, cp_len, n_sym
    K = N - 1
    rng = np.random.default_rng(seed)
    pilot_asc = make_pilot(K, seed) if pilot_asc is None else np.asarray(pilot_asc, np.complex64)
    src_idx = rng.integers(0, 1 << qam_bits, size=(F, S - 1, K), dtype=np.uint8)
    data_asc = qam_map_indices(src_idx, qam_bits)
    tx_bin = np.zeros((F, S, N), np.complex128)
    tx_bin[:, 0, 1:] = asc_to_bin(pilot_asc.astype(np.complex128))
    tx_bin[:, 1:, 1:] = asc_to_bin(data_asc)
    if channel == "identity":
        h = np.ones((F, A, K), np.complex128)
    elif channel == "unit":  # unit modulus, random phase: no deep fades (for single-antenna cases)
        h = np.exp(2j * np.pi * rng.random((F, A, K)))
    else:
        h = (rng.standard_normal((F, A, K)) + 1j * rng.standard_normal((F, A, K))) / np.sqrt(2.0)
    hfull = np.zeros((F, A, N), np.complex128)
    hfull[:, :, 1:] = h
    yf = tx_bin[:, :, None, :] * hfull[:, None, :, :]          # [F,S,A,N]
    yt = np.fft.ifft(yf, axis=-1)                               # unnormalised forward FFT undoes this
    if snr_db is not None:
        sig_pow = K / (N * N)                                   # per time sample, unit-power bins
        sigma = np.sqrt(sig_pow / (10.0 ** (snr_db / 10.0)) / 2.0)
        yt = yt + sigma * (rng.standard_normal(yt.shape) + 1j * rng.standard_normal(yt.shape))
    rx = np.concatenate([yt[..., N - C:], yt], axis=-1) if C > 0 else yt
    return {"rx": np.ascontiguousarray(rx.astype(np.complex64)), "pilot_asc": pilot_asc,
            "src_idx": src_idx, "h_true": h}


def make_frames_torch(n_frames, cfg, device, seed=None, snr_db="cfg", chunk=64):
    """Full-size batch built directly on `device` with torch (generator only; the receive
    path never uses torch.fft).  Returns (rx [F,S,A,N+C] complex64, pilot_asc numpy [K],
    src_idx [F,S-1,K] uint8 tensor in ascending-frequency order)."""
    import torch

    A, N, C, S, b = cfg.n_ant, cfg.fft_size, cfg.cp_len, cfg.n_sym, cfg.qam_bits
    K = N - 1
    seed = cfg.seed if seed is None else seed
    snr = cfg.snr_db if snr_db == "cfg" else snr_db
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    pilot_asc = make_pilot(K, seed)
    pil_bin = torch.from_numpy(asc_to_bin(pilot_asc)).to(device)
    lut = torch.from_numpy(qam_map_indices(np.arange(1 << b), b).astype(np.complex64)).to(device)
    rx = torch.empty((n_frames, S, A, N + C), dtype=torch.complex64, device=device)
    src = torch.empty((n_frames, S - 1, K), dtype=torch.uint8, device=device)
    shift = (K + 1) // 2
    for f0 in range(0, n_frames, chunk):
        nf = min(chunk, n_frames - f0)
        idx = torch.randint(0, 1 << b, (nf, S - 1, K), generator=g, device=device, dtype=torch.uint8)
        src[f0:f0 + nf] = idx
        tx = torch.zeros((nf, S, N), dtype=torch.complex64, device=device)
        tx[:, 0, 1:] = pil_bin
        tx[:, 1:, 1:] = torch.roll(lut[idx.long()], -shift, dims=-1)
        h = torch.zeros((nf, A, N), dtype=torch.complex64, device=device)
        hr = torch.randn((nf, A, K, 2), generator=g, device=device) * (0.5 ** 0.5)
        h[:, :, 1:] = torch.view_as_complex(hr)
        yt = torch.fft.ifft(tx[:, :, None, :] * h[:, None, :, :], dim=-1)
        if snr is not None:
            sigma = (K / (N * N) / (10.0 ** (snr / 10.0)) / 2.0) ** 0.5
            yt = yt + sigma * torch.view_as_complex(torch.randn((*yt.shape, 2), generator=g, device=device))
        rx[f0:f0 + nf, :, :, C:] = yt
        if C > 0:
            rx[f0:f0 + nf, :, :, :C] = yt[..., N - C:]
        del tx, h, hr, yt, idx
    return rx, pilot_asc, src


def make_pn(length=255, seed=0x1D):
    """+-1 maximal-length sequence (8-bit LFSR x^8+x^6+x^5+x^4+1 -> period 255), as complex64.  Stands in for
    the reference's PNSeq_255_MaxLenSeq.dat (rx_and_corr.cpp:228), which is not shipped."""
    reg = seed & 0xFF or 1
    out = np.empty(length, np.float32)
    for i in range(length):
        out[i] = 1.0 if (reg & 1) else -1.0
        fb = ((reg >> 0) ^ (reg >> 2) ^ (reg >> 3) ^ (reg >> 4)) & 1
        reg = (reg >> 1) | (fb << 7)
    return out.astype(np.complex64)


def make_capture(rx_frame, pn, offset, samps=None, noise=0.02, seed=0):
    """Wrap one frame [S][A][N+C] into the two per-channel capture buffers the reference's receive loop sees
    (rx_and_corr.cpp:305-312): noise, then conj(pn) starting at `offset` of buffer 1 (the correlator multiplies by
    pn without conjugating, :348), then the frame, whose tail wraps into buffer 2.  Returns (buf1, buf2) [A][samps]."""
    S, A, row = rx_frame.shape
    L = pn.shape[0]
    frame = np.ascontiguousarray(np.transpose(rx_frame, (1, 0, 2))).reshape(A, S * row)   # [A][S*row]
    samps = samps or (L + S * row)
    assert samps - L >= S * row and 0 <= offset and offset + L <= samps
    rng = np.random.default_rng(seed)
    scale = float(np.abs(frame).mean()) or 1.0
    def noise_buf():
        return (noise * scale * (rng.standard_normal((A, samps)) + 1j * rng.standard_normal((A, samps)))).astype(np.complex64)
    buf1, buf2 = noise_buf(), noise_buf()
    buf1[:, offset:offset + L] = np.conj(pn)[None, :]
    n_first = samps - offset - L
    stream = np.concatenate([frame, noise_buf()[:, :max(0, samps - L - S * row)]], axis=1)  # samps-L per channel
    buf1[:, offset + L:] = stream[:, :n_first]
    buf2[:, :offset] = stream[:, n_first:n_first + offset]
    return buf1, buf2
