"""Single-frame latency of the device-resident call for a config, one-launch policy vs the kernel pair.
usage: python tools/latency_probe.py [--config c5] [--launches 2000]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import ofdm_b200 as m  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c5")
ap.add_argument("--launches", type=int, default=2000)
args = ap.parse_args()
cfg = m.CONFIGS[args.config]
dev = torch.device("cuda", 0)
rx, pilot_asc, _ = m.synth.make_frames_torch(1, cfg, dev)
comb = torch.empty((1, cfg.n_sym - 1, cfg.K, 2), device=dev)
bits = torch.empty((1, cfg.n_sym - 1, cfg.bits_row_bytes), device=dev, dtype=torch.uint8)
rxf = torch.view_as_real(rx)
for oneshot in (1, 0, 1, 0):
    with m.LsMrcReceiver.from_config(cfg, device=0) as r:
        r.set_pilot(pilot_asc)
        r.set_oneshot(oneshot)
        for _ in range(300):
            r.demod_frames_device(rxf, 1, comb, bits)
        r.sync()
        lat = np.empty(args.launches)
        for i in range(args.launches):
            t0 = time.perf_counter()
            r.demod_frames_device(rxf, 1, comb, bits)
            r.sync()
            lat[i] = time.perf_counter() - t0
        r.set_timing(True)
        dv = []
        for _ in range(200):
            r.demod_frames_device(rxf, 1, comb, bits)
            a, b = r.last_kernel_ms()
            dv.append((a * 1e3, b * 1e3))
        dv = np.array(dv)
        print(f"{args.config} oneshot={oneshot}: host p50 {np.percentile(lat, 50) * 1e6:.2f} us  p99 {np.percentile(lat, 99) * 1e6:.2f} us"
              f"  | events: pilot {np.median(dv[:, 0]):.2f} us, data/fused {np.median(dv[:, 1]):.2f} us  [{r.describe_plan()}]", flush=True)
