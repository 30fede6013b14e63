"""tools/sanitize_probe.py -- small launches of every plan and launch policy, for compute-sanitizer runs where the tool is available
(memcheck / racecheck / synccheck): `compute-sanitizer --tool racecheck python tools/sanitize_probe.py`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ofdm_b200 as m  # noqa: E402

dev = torch.device("cuda:0")
CASES = [(4, 64, 16, 6, 2, 3), (5, 128, 8, 4, 4, 2), (8, 256, 32, 5, 6, 2), (6, 512, 32, 4, 4, 2), (8, 1024, 64, 6, 4, 3),
         (7, 1024, 9, 3, 6, 2), (4, 2048, 144, 3, 4, 2), (3, 4096, 288, 3, 6, 1)]
for (A, N, C, S, b, F) in CASES:
    K = N - 1
    rx = torch.randn((F, S, A, N + C, 2), device=dev)
    comb = torch.empty((F, S - 1, K, 2), device=dev)
    bits = torch.empty((F, S - 1, (K * b + 7) // 8), device=dev, dtype=torch.uint8)
    hc = torch.empty((F, A, K, 2), device=dev)
    hs = torch.empty((F, K), device=dev)
    llr = torch.empty((F, S - 1, K, b), device=dev)
    for policy in (2, 0):
        with m.LsMrcReceiver(A, N, C, S, b) as r:
            r.set_pilot(m.synth.make_pilot(K, 1))
            r.set_oneshot(policy)
            r.demod_frames_device(rx, F, comb, bits, hc, hs)
            r.demod_frames_device_soft(rx, F, comb, llr, 0.5, bits)
            r.sync()
            print(f"A={A} N={N} C={C} S={S} b={b} F={F} policy={policy}: ok, finite={bool(torch.isfinite(comb).all())}", flush=True)
# a batch large enough for persistent CTAs to take several work items (ticket loop, ring and bulk-copy phases wrap)
A, N, C, S, b, F = 6, 1024, 64, 9, 4, 160
rx = torch.randn((F, S, A, N + C, 2), device=dev)
comb = torch.empty((F, S - 1, N - 1, 2), device=dev)
bits = torch.empty((F, S - 1, ((N - 1) * b + 7) // 8), device=dev, dtype=torch.uint8)
with m.LsMrcReceiver(A, N, C, S, b) as r:
    r.set_pilot(m.synth.make_pilot(N - 1, 1))
    r.demod_frames_device(rx, F, comb, bits)
    r.sync()
print("large batch ok", bool(torch.isfinite(comb).all()))
