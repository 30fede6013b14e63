"""tools/save_profiles_r02.py -- condense the round-2 ncu captures in gpurun_out/ into profiles/ (tracked):
per config c2/c3/c4 the --set full summary of the pilot and data kernel, their opcode mix and stall hot spots
(profiles/r02_ncu_<cfg>.txt), and the measured DRAM bytes per launch (profiles/traffic.json)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go, out = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
FRAMES = {"c2": 256, "c3": 384, "c4": 192}
traffic = {}


def run(tool, *args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool), *args], capture_output=True, text=True).stdout


def dram(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h = rows[0]
    res = []
    for r in rows[2:]:
        d = dict(zip(h, r))
        unit = dict(zip(h, rows[1]))
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            v = float(d[k])
            u = unit[k].lower()
            tot += v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
        res.append((d["Kernel Name"], tot, float(d["gpu__time_duration.sum"])))
    return res


for cfg, frames in FRAMES.items():
    rep = os.path.join(go, f"r02_final_{cfg}.ncu-rep")
    if not os.path.exists(rep):
        continue
    k = dram(rep)
    single = len(k) == 1  # 2048/4096 points: pilot and data items in one persistent launch (lsmrc_frames_sh)
    with open(os.path.join(out, f"r02_ncu_{cfg}.txt"), "w") as f:
        f.write(f"round 2, config {cfg}: ncu --set full --clock-control none --import-source on -k regex:lsmrc_ -s {1 if single else 2} "
                f"-c {1 if single else 2} python tools/quick_bench.py --config {cfg} --frames {frames} --iters 2\n"
                + ("(one launch: pilot items, then data items, lsmrc_frames_sh" if single else "(launch 0 = pilot kernel, launch 1 = data kernel")
                + "; numbers under the profiler are not bench values)\n\n")
        f.write(run("ncu_summary.py", rep))
        for sec in ((0,) if single else (1, 0)):
            f.write("\n---- executed warp-instructions per opcode ----\n" + run("ncu_opmix.py", rep, str(sec)))
        f.write("\n---- top stall instructions of the " + ("" if single else "data ") + "kernel (SASS) ----\n"
                + run("ncu_hotspots.py", rep, "0" if single else "1", "25"))
    if single:
        traffic[cfg] = {"frames": frames, "kernel": k[0][0][:60], "kernel_dram_bytes": k[0][1],
                        "source": f"profiles/r02_ncu_{cfg}.txt (dram__bytes_read.sum + dram__bytes_write.sum, the single launch)"}
    else:
        traffic[cfg] = {"frames": frames, "pilot_kernel": k[0][0][:60], "pilot_kernel_dram_bytes": k[0][1],
                        "data_kernel": k[1][0][:60], "data_kernel_dram_bytes": k[1][1],
                        "source": f"profiles/r02_ncu_{cfg}.txt (dram__bytes_read.sum + dram__bytes_write.sum, one launch each)"}
json.dump(traffic, open(os.path.join(out, "traffic.json"), "w"), indent=1)
print(json.dumps(traffic, indent=1))
