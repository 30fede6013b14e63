"""tools/fused_probe.py -- developer probe: single-launch frames kernel vs the kernel pair, many repetitions; reports which outputs,
frames and antennas differ."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ofdm_b200 as m  # noqa: E402

A, N, C, S, b, F = (int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "32,4096,288,4,6,150").split(","))
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
K = N - 1
dev = torch.device("cuda:0")
rx = torch.randn((F, S, A, N + C, 2), device=dev)


def run(r):
    comb = torch.zeros((F, S - 1, K, 2), device=dev)
    bits = torch.zeros((F, S - 1, (K * b + 7) // 8), device=dev, dtype=torch.uint8)
    hc = torch.zeros((F, A, K, 2), device=dev)
    hs = torch.zeros((F, K), device=dev)
    torch.cuda.synchronize()   # the fills run on torch's stream, the receiver on its own
    r.demod_frames_device(rx, F, comb, bits, hc, hs)
    r.sync()
    return {"comb": comb, "bits": bits, "hconj": hc, "hsqrd": hs}


for lead in ["-"]:
    with m.LsMrcReceiver(A, N, C, S, b) as r:
        r.set_pilot(m.synth.make_pilot(K, 1))
        r.set_one_launch_frames(False)
        pair = run(r)
        r.set_one_launch_frames(True)
        bad = 0
        for it in range(reps):
            one = run(r)
            for k in one:
                if not torch.equal(one[k], pair[k]):
                    bad += 1
                    diff = (one[k] != pair[k])
                    frames = diff.reshape(F, -1).any(1).nonzero().flatten().tolist()
                    rel = float((one[k].float() - pair[k].float()).abs().max() / pair[k].float().abs().max())
                    msg = f"rep {it}: {k} differs in {int(diff.sum())} elements (max rel {rel:.2e}), frames {frames[:12]}"
                    if k == "hconj":
                        ants = diff.reshape(F, A, -1).any(2).any(0).nonzero().flatten().tolist()
                        msg += f" antennas {ants[:16]}"
                    if k == "comb":
                        syms = diff.reshape(F, S - 1, -1).any(2).any(0).nonzero().flatten().tolist()
                        cols = diff.reshape(-1, K, 2).any(2).any(0).nonzero().flatten().tolist()
                        msg += f" symbols {syms} columns {cols[:8]}..{cols[-4:]} ({len(cols)})"
                    print(msg)
        print(f"{reps} repetitions, {r.one_launch_frames_count()} single launches, {bad} mismatching outputs", flush=True)
