"""tools/ncu_hotspots.py <report> <section-index> [top-n] -- top stall instructions (SASS) of one profiled launch."""
import csv
import subprocess
import sys

rep, sec = sys.argv[1], int(sys.argv[2])
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(sec),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
block = rows[starts[0]:starts[1]]
print(block[0][1][:120])
hdr = block[1]
data = [r for r in block[2:] if len(r) == len(hdr)]
ia, isamp = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_')]
tot = sum(int(r[isamp]) for r in data)
print('total samples', tot, 'instructions', len(data))
agg = {}
for r in data:
    for i in stall_cols:
        if r[i].isdigit():
            agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i])
print({k[6:]: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]})
items = sorted(((int(r[isamp]), i) for i, r in enumerate(data)), reverse=True)[:topn]
for s, i in items:
    r = data[i]
    reasons = sorted(((int(r[c]), hdr[c][6:]) for c in stall_cols if r[c].isdigit() and int(r[c]) > 0), reverse=True)[:2]
    print(f"{s:6d} {100 * s / tot:5.1f}%  #{i:5d} {r[ia].strip():60s} {reasons}")
