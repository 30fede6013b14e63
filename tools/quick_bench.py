"""tools/quick_bench.py -- developer timing loop (not the contract bench): device-resident
frames through lsmrc_demod_frames_device, CUDA-event kernel times, algorithmic GB/s."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ofdm_b200 as m

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c2")
ap.add_argument("--frames", type=int, default=64)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--nsym", type=int, default=0)
ap.add_argument("--dims", default="", help="A,N,C,S,b instead of a named config")
args = ap.parse_args()
if args.dims:
    A_, N_, C_, S_, b_ = (int(x) for x in args.dims.split(","))
    cfg = m.RxConfig("custom", A_, N_, C_, S_, b_, 15.0, 1, "custom dims")
else:
    cfg = m.CONFIGS[args.config]
if args.nsym:
    import dataclasses
    cfg = dataclasses.replace(cfg, n_sym=args.nsym)
F = args.frames
dev = torch.device("cuda:0")
rx = torch.randn((F, cfg.n_sym, cfg.n_ant, cfg.fft_size + cfg.cp_len, 2), device=dev, dtype=torch.float32)
comb = torch.empty((F, cfg.n_sym - 1, cfg.K, 2), device=dev, dtype=torch.float32)
bits = torch.empty((F, cfg.n_sym - 1, cfg.bits_row_bytes), device=dev, dtype=torch.uint8)
r = m.LsMrcReceiver.from_config(cfg, max_frames=1)
r.set_pilot(m.synth.make_pilot(cfg.K, 1))
r.set_timing(True)
if os.environ.get("LSMRC_ONE_LAUNCH") == "0":
    r.set_one_launch_frames(False)
print(r.describe_plan())
for it in range(args.iters):
    r.demod_frames_device(rx, F, comb, bits)
    p_ms, d_ms = r.last_kernel_ms()
    tot = p_ms + d_ms
    by = cfg.algorithmic_bytes_per_frame * F
    print(f"iter {it}: pilot {p_ms:.3f} ms data {d_ms:.3f} ms  -> {by / tot / 1e6:.1f} GB/s algorithmic, "
          f"{cfg.antenna_samples_per_frame * F / tot / 1e6:.2f} G antenna-samples/s")
