"""tools/ncu_opmix.py <report> <launch-index> -- executed warp-instructions per SASS opcode of one profiled launch."""
import csv
import subprocess
import sys
from collections import Counter

rep, sec = sys.argv[1], int(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(sec),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
block = rows[starts[0]:starts[1]]
hdr = block[1]
data = [r for r in block[2:] if len(r) == len(hdr)]
ia = hdr.index('Source')
ie = hdr.index('# Instructions Executed') if '# Instructions Executed' in hdr else [i for i, h in enumerate(hdr) if 'Instructions Executed' in h][0]
c = Counter()
for r in data:
    op = r[ia].strip().split()
    if not op:
        continue
    name = op[1] if op[0].startswith('@') else op[0]
    c[name.split('.')[0]] += int(r[ie] or 0)
tot = sum(c.values())
print(block[0][1][:100], "total warp-instr", tot)
for k, v in c.most_common(28):
    print(f"{k:12s} {v:14d} {100 * v / tot:5.1f}%")
