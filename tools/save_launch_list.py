"""tools/save_launch_list.py <tag> -- profiles/<tag>_launch_list.md from gpurun_out/launches.csv (the ncu launch list of the bench
command, tools/gpu/r02_final_prof.sh): per-kernel totals, and for this library's kernels their share of the step they belong to."""
import collections
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
rows = [r for r in csv.reader(open(os.path.join(ROOT, "gpurun_out", "launches.csv"))) if len(r) > 10]
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    a = agg.setdefault(r[ik], [0, 0.0])
    a[0] += 1
    a[1] += float(r[iv].replace(",", ""))
tot = sum(a[1] for a in agg.values())
with open(os.path.join(ROOT, "profiles", f"{tag}_launch_list.md"), "w") as f:
    f.write(f"# {tag}: ncu launch list of `python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras`\n\n"
            "`ncu --metrics gpu__time_duration.sum --clock-control none -c 3000` after the plain run of the same command exited 0\n"
            "(cold-cache, serialised: compare shares, not absolutes).  torch kernels are the synthetic-input generator, outside\n"
            "the timed regions.  Inside them only this library's kernels launch: `lsmrc_kernel<Plan<1024,…>, 0|1>` = pilot | data\n"
            "kernel of the c2 region (`value`); `lsmrc_frames_sh<Plan<4096,…>>` = the `scaling_c4` block (pilot items, then data\n"
            "items, in one persistent launch).\n\n")
    f.write("| launches | total us | share of all | avg us | kernel |\n|---|---|---|---|---|\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| {n} | {t / 1e3:.1f} | {100 * t / tot:.1f}% | {t / n / 1e3:.1f} | `{k[:110]}` |\n")
    f.write("\n## our kernels: share of each step\n\n| kernel | launches | avg us | share of its step |\n|---|---|---|---|\n")
    ours = {k: v for k, v in agg.items() if "lsmrc" in k}
    step_of = lambda k: "c2" if "Plan<1024" in k else "c4"
    step_tot = collections.Counter()
    for k, (n, t) in ours.items():
        step_tot[step_of(k)] += t
    for k, (n, t) in ours.items():
        f.write(f"| `{k[:100]}` | {n} | {t / n / 1e3:.1f} | {100 * t / step_tot[step_of(k)]:.1f}% ({step_of(k)}) |\n")
print(open(os.path.join(ROOT, "profiles", f"{tag}_launch_list.md")).read()[-900:])
