"""Named workloads of BASELINE.json, frozen as concrete dimensions.

A = RX antennas (numOfRows), N = FFT size (dimension), C = cyclic prefix (prefix),
S = symbols per frame (lenOfBuffer; symbol 0 is the pilot), b = bits per QAM symbol.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class RxConfig:
    name: str
    n_ant: int
    fft_size: int
    cp_len: int
    n_sym: int
    qam_bits: int
    snr_db: float
    seed: int
    note: str = ""

    @property
    def K(self) -> int:
        return self.fft_size - 1

    @property
    def slot_elems(self) -> int:
        return self.n_ant * (self.fft_size + self.cp_len)

    @property
    def frame_elems(self) -> int:
        return self.n_sym * self.slot_elems

    @property
    def antenna_samples_per_frame(self) -> int:
        """One antenna-sample = one complex64 time sample from one antenna, CP included."""
        return self.frame_elems

    @property
    def bits_row_bytes(self) -> int:
        return (self.K * self.qam_bits + 7) // 8

    @property
    def rx_bytes_per_frame(self) -> int:
        return 8 * self.frame_elems

    @property
    def algorithmic_bytes_per_frame(self) -> int:
        """Compulsory HBM traffic of a fully fused receiver (the CP is never read):
        input 8*A*S*N + Hconj out 8*A*K + sum|H|^2 out 4*K + combined out 8*(S-1)*K
        + packed bits (S-1)*ceil(K*b/8)."""
        A, N, S, K = self.n_ant, self.fft_size, self.n_sym, self.K
        return 8 * A * S * N + 8 * A * K + 4 * K + 8 * (S - 1) * K + (S - 1) * self.bits_row_bytes


CONFIGS = {
    # cpuLS_main reference case: 64-pt FFT, 16-sample CP, 4 antennas, 1 pilot + 15 data, QPSK
    "c1": RxConfig("c1", 4, 64, 16, 16, 2, 10.0, 1235, "cpuLS_main reference (ring plumbing)"),
    # headline: 1024-pt FFT, 64 antennas, 1 pilot + 100 data symbols, 16-QAM
    "c2": RxConfig("c2", 64, 1024, 64, 101, 4, 15.0, 1236, "10k-frame batch, single B200"),
    # massive MIMO slots streamed through the ring
    "c3": RxConfig("c3", 128, 2048, 144, 14, 4, 15.0, 1237, "ring-streamed 14-symbol slots"),
    # scaling run
    "c4": RxConfig("c4", 256, 4096, 288, 14, 6, 20.0, 1238, "frame batches sharded over 1/2/4/8 GPUs"),
    # latency bound
    "c5": RxConfig("c5", 16, 64, 16, 16, 2, 10.0, 1239, "one frame per launch, p50/p99"),
}
