"""Comparison helpers shared by the parity tests."""
import numpy as np

# north_star: H estimates and combined symbols within 1e-5 relative error (fp32), bits exact.
REL_TOL = 1e-5


def rel_errors(got, ref):
    """(max-norm, 2-norm) error relative to the reference's max / 2-norm, per SURVEY Appendix A.11."""
    got = np.asarray(got)
    ref = np.asarray(ref)
    d = got.astype(np.complex128) - ref.astype(np.complex128)
    mx = float(np.abs(d).max() / max(np.abs(ref).max(), 1e-30))
    l2 = float(np.linalg.norm(d.ravel()) / max(np.linalg.norm(ref.astype(np.complex128).ravel()), 1e-30))
    return mx, l2


def assert_close(got, ref, what, tol=REL_TOL):
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    assert np.isfinite(np.asarray(got).view(np.float32)).all(), f"{what}: non-finite values"
    mx, l2 = rel_errors(got, ref)
    assert mx <= tol and l2 <= tol, f"{what}: max-rel {mx:.3e}, l2-rel {l2:.3e} > {tol:g}"
    return mx, l2


def threshold_margin(sym, qam_bits):
    """smallest distance of any combined symbol component to a demap decision threshold"""
    s = np.asarray(sym)
    comps = np.concatenate([s.real.ravel(), s.imag.ravel()]).astype(np.float64)
    d = np.abs(comps)
    if qam_bits == 4:
        d = np.minimum(d, np.abs(np.abs(comps) - 2 / np.sqrt(10)))
    elif qam_bits == 6:
        for t in (2 / np.sqrt(42), 4 / np.sqrt(42), 6 / np.sqrt(42)):
            d = np.minimum(d, np.abs(np.abs(comps) - t))
    return float(d.min())


def symbol_margins(sym, qam_bits):
    """per-symbol distance of the closer component to its nearest demap decision threshold (same shape as sym)"""
    s = np.asarray(sym)

    def axis(u):
        u = np.abs(u.astype(np.float64))
        d = u.copy()                      # threshold at 0
        if qam_bits == 4:
            d = np.minimum(d, np.abs(u - 2 / np.sqrt(10)))
        elif qam_bits == 6:
            for t in (2 / np.sqrt(42), 4 / np.sqrt(42), 6 / np.sqrt(42)):
                d = np.minimum(d, np.abs(u - t))
        return d

    return np.minimum(axis(s.real), axis(s.imag))


def assert_bits_match(got_bits, ref_bits, ref_combined, qam_bits, what, got_combined=None, tol=REL_TOL):
    """Demapped bits must equal the oracle's.  The only admissible difference is a symbol whose ORACLE value lies
    closer to a decision threshold than the two fp32 evaluations lie apart (|got - ref| when the GPU's combined
    symbols are given, else the 1e-5 parity tolerance): the two may then legitimately fall on different sides.
    Every differing symbol is checked individually and reported."""
    got_bits = np.asarray(got_bits)
    ref_bits = np.asarray(ref_bits)
    assert got_bits.shape == ref_bits.shape, f"{what}: bits shape {got_bits.shape} vs {ref_bits.shape}"
    if np.array_equal(got_bits, ref_bits):
        return 0
    K = ref_combined.shape[-1]
    diff = np.unpackbits(got_bits ^ ref_bits, axis=-1, bitorder="little")[..., : K * qam_bits]
    bad = diff.reshape(*diff.shape[:-1], K, qam_bits).any(-1)
    margins = symbol_margins(ref_combined, qam_bits)
    scale = float(np.abs(ref_combined).max())
    if got_combined is not None:
        d = np.asarray(got_combined).astype(np.complex128) - np.asarray(ref_combined).astype(np.complex128)
        allowed = np.minimum(np.maximum(np.abs(d.real), np.abs(d.imag)), tol * scale)
    else:
        allowed = np.full(margins.shape, tol * scale)
    where = np.argwhere(bad)
    worst = [(tuple(int(i) for i in w), float(margins[tuple(w)])) for w in where[:8]]
    unexplained = bad & (margins > allowed)
    assert not unexplained.any(), (f"{what}: {int(bad.sum())} demapped symbols differ, {int(unexplained.sum())} of them are "
                                   f"not threshold ties (oracle margin > |got - ref|); first (index, oracle margin): {worst}")
    return int(bad.sum())


def load_golden(path, ofdm):
    """Golden fixture -> dict.  Full-dimension fixtures store the generator seed instead of the (31-126 MB) input
    frame: the frame is regenerated here and checked against the stored SHA-256, so a drifting generator cannot
    silently change what the fixture pins."""
    import hashlib

    g = dict(np.load(path))
    A, N, C, S, b, F = (int(v) for v in g["dims"])
    if "rx" not in g:
        d = ofdm.synth.make_frames(F, A, N, C, S, b, snr_db=float(g["snr_db"]), seed=int(g["rx_seed"]))
        digest = hashlib.sha256(np.ascontiguousarray(d["rx"]).view(np.uint8).tobytes()).hexdigest()
        assert digest == str(g["rx_sha256"]), "synthetic generator no longer reproduces the fixture's input frame"
        assert np.array_equal(d["pilot_asc"], g["pilot_asc"])
        g["rx"] = d["rx"]
        g["src_idx"] = d["src_idx"]
    return g


HCONJ_SAMPLE_STRIDE = 97  # tests/golden/make_golden.py
