"""tools/ring_probe.py -- c3 frames through the shm ring with different producer thread counts (developer probe)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import ofdm_b200 as m

for th in (int(x) for x in (sys.argv[1:] or ["4", "8", "12"])):
    r = bench.ring_stream_leg(m, n_frames=384, feeder_threads=th)
    print(th, {k: r.get(k) for k in ("seconds", "h2d_gbs", "seconds_from_first_submission", "h2d_gbs_from_first_submission", "error")})
    print("   ", (r.get("overlap") or ["", ""])[1:4])
