"""tests/golden/make_golden.py -- regenerates the committed golden fixtures.

Run in the build container (needs /root/reference to have been compiled by `make -C oracle ref`):
    python tests/golden/make_golden.py
Each fixture holds seeded synthetic input frames and the outputs of the REFERENCE's own CPU code
(cpuLS.hpp compiled from /root/reference with the FFT/BLAS header shims, driven by
oracle/ref_driver.cpp), plus the hard-demapped bits of those outputs.  The fixtures travel to
the GPU box; the reference does not.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ofdm_b200 as m  # noqa: E402
from oracle import oracle_py  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
# name -> (A, N, C, S, qam_bits, frames, seed, snr_db, pilot) ; pilot None = reference fallback 0.707+0.707i
CASES = {
    "c1_A4_N64_C16_S16_qpsk": (4, 64, 16, 16, 2, 2, 1235, 10.0, "qpsk"),
    "c1_fallback_pilot": (4, 64, 16, 16, 2, 1, 77, 10.0, None),
    "A8_N256_C32_S6_16qam": (8, 256, 32, 6, 4, 2, 5, 15.0, "qpsk"),
    "A16_N1024_C64_S5_16qam": (16, 1024, 64, 5, 4, 1, 9, 15.0, "qpsk"),
}


# Full BASELINE dimensions (configs c2, c3, c4; one frame each).  The input frame is 31-126 MB, so these fixtures hold
# the generator's seed plus a SHA-256 of the generated samples instead of the samples, the reference's combined
# symbols, sum|H|^2 and bits in full, and H as a SHA-256 plus a strided sample (H alone is 8.4 MB at c4).
# name -> (A, N, C, S, qam_bits, seed, snr_db)
FULL_CASES = {
    "c2_full_A64_N1024_C64_S101_16qam": (64, 1024, 64, 101, 4, 1236, 15.0),
    "c3_full_A128_N2048_C144_S14_16qam": (128, 2048, 144, 14, 4, 1237, 15.0),
    "c4_full_A256_N4096_C288_S14_64qam": (256, 4096, 288, 14, 6, 1238, 20.0),
}
HCONJ_SAMPLE_STRIDE = 97


def sha256(a):
    import hashlib

    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8).tobytes()).hexdigest()


def make_full():
    for name, (A, N, C, S, b, seed, snr) in FULL_CASES.items():
        d = m.synth.make_frames(1, A, N, C, S, b, snr_db=snr, seed=seed)
        ref = oracle_py.run_reference(d["rx"], d["pilot_asc"], C)
        assert ref is not None, "oracle/_ref binary missing for " + name
        bits = np.stack([np.stack([oracle_py.demap_row(ref["combined"][f, s], b)[0] for s in range(S - 1)]) for f in range(1)])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), rx_seed=np.array(seed), snr_db=np.array(snr),
                            rx_sha256=np.array(sha256(d["rx"])), pilot_asc=d["pilot_asc"], hsqrd=ref["hsqrd"],
                            combined=ref["combined"], bits=bits, hconj_sha256=np.array(sha256(ref["hconj"])),
                            hconj_sample=ref["hconj"].ravel()[::HCONJ_SAMPLE_STRIDE].copy(),
                            dims=np.array([A, N, C, S, b, 1]))
        print(name, "ok", os.path.getsize(os.path.join(HERE, name + ".npz")) // 1024, "KiB")


def main():
    oracle_py.build(ref=True)
    make_full()
    for name, (A, N, C, S, b, F, seed, snr, pil) in CASES.items():
        d = m.synth.make_frames(F, A, N, C, S, b, snr_db=snr, seed=seed)
        K = N - 1
        if pil is None:
            # the frames must be built on the fallback pilot the reference substitutes for a missing Pilots.dat
            fb = np.full(K, 0.707 + 0.707j, np.complex64)
            d = m.synth.make_frames(F, A, N, C, S, b, snr_db=snr, seed=seed, pilot_asc=fb)
        ref = oracle_py.run_reference(d["rx"], None if pil is None else d["pilot_asc"], C)
        assert ref is not None, "oracle/_ref binary missing for " + name
        bits = np.stack([np.stack([oracle_py.demap_row(ref["combined"][f, s], b)[0] for s in range(S - 1)]) for f in range(F)])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), rx=d["rx"], pilot_asc=d["pilot_asc"],
                            hconj=ref["hconj"], hsqrd=ref["hsqrd"], combined=ref["combined"], bits=bits,
                            src_idx=d["src_idx"], dims=np.array([A, N, C, S, b, F]), fallback_pilot=np.array(pil is None))
        print(name, "ok", os.path.getsize(os.path.join(HERE, name + ".npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
