"""tools/trace_summary.py <trace.csv> -- how far the three phases of consecutive ring submissions overlap.

Input: the CSV `stream_main --trace` writes (lsmrc_ring_trace: per submission the CUDA-event times, in ms since the
first submission, at which the lane's stream reached the submission, its H2D copy had finished, its kernels had
finished, and its results were on the host).  Output: per-phase totals, and the share of the wall time during which
phases of DIFFERENT submissions ran at the same time."""
import csv
import sys

rows = [r for r in csv.DictReader(open(sys.argv[1]))]
subs = sorted(({k: float(v) for k, v in r.items()} for r in rows), key=lambda r: r["submission"])
ph = {"h2d": [], "kernels": [], "d2h": []}
for r in subs:
    ph["h2d"].append((r["t_enqueued_ms"], r["t_h2d_done_ms"]))
    ph["kernels"].append((r["t_h2d_done_ms"], r["t_kernels_done_ms"]))
    ph["d2h"].append((r["t_kernels_done_ms"], r["t_results_on_host_ms"]))
t_end = max(r["t_results_on_host_ms"] for r in subs)
t_beg = min(r["t_enqueued_ms"] for r in subs)
wall = t_end - t_beg
# sweep: at every elementary interval count active phases by kind
edges = sorted({t for iv in ph.values() for a, b in iv for t in (a, b)})
busy = {k: 0.0 for k in ph}
both = {"h2d+kernels": 0.0, "kernels+d2h": 0.0, "h2d+d2h": 0.0, "all three": 0.0, "idle": 0.0}
for a, b in zip(edges, edges[1:]):
    mid, dt = 0.5 * (a + b), b - a
    act = {k: sum(1 for lo, hi in iv if lo <= mid < hi) for k, iv in ph.items()}
    for k in ph:
        if act[k]:
            busy[k] += dt
    if act["h2d"] and act["kernels"]:
        both["h2d+kernels"] += dt
    if act["kernels"] and act["d2h"]:
        both["kernels+d2h"] += dt
    if act["h2d"] and act["d2h"]:
        both["h2d+d2h"] += dt
    if all(act.values()):
        both["all three"] += dt
    if not any(act.values()):
        both["idle"] += dt
n = len(subs)
frames = sum(int(r["frames"]) for r in subs)
print(f"{n} submissions ({frames} frames) over {wall:.3f} ms on {len({int(r['lane']) for r in subs})} lanes")
for k, iv in ph.items():
    d = [hi - lo for lo, hi in iv]
    print(f"  {k:8s} mean {sum(d) / n:8.3f} ms per submission, busy {busy[k]:9.3f} ms = {100 * busy[k] / wall:5.1f} % of the wall time")
print(f"  sum of the three phases if run back to back: {sum(hi - lo for iv in ph.values() for lo, hi in iv):.3f} ms "
      f"= {sum(hi - lo for iv in ph.values() for lo, hi in iv) / wall:.2f} x the wall time")
for k, v in both.items():
    print(f"  {k:12s} {v:9.3f} ms = {100 * v / wall:5.1f} % of the wall time")
