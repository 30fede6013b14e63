"""tools/ring_probe.py -- c3 frames through the shm ring with different producer thread counts (developer probe).
usage: ring_probe.py [threads ...]; a trailing 'x' (e.g. 8x) runs the producer with --first-lap-only: it writes every slot
once and from then on only publishes, which shows the consumer's ceiling without producer traffic in host memory."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import ofdm_b200 as m

for spec in (sys.argv[1:] or ["4", "8", "12"]):
    th = int(spec.rstrip("x"))
    extra = ("--first-lap-only",) if spec.endswith("x") else ()
    r = bench.ring_stream_leg(m, n_frames=768, feeder_threads=th, feeder_args=extra)
    print(spec, {k: r.get(k) for k in ("seconds", "h2d_gbs", "seconds_from_first_submission", "h2d_gbs_from_first_submission", "error")})
    print("   ", (r.get("overlap") or ["", ""])[0:4])
