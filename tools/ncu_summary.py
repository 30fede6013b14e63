"""tools/ncu_summary.py <report.ncu-rep> [kernel-index] -- print the metrics we track from an ncu report."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.avg.per_second',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("====", d['Kernel Name'][:90], d.get('Grid Size'), d.get('Block Size'))
    for k in KEYS:
        if k in d:
            print(f"  {k:80s} {d[k]} {units[hdr.index(k)]}")
    for k in hdr:
        if 'issue_stalled' in k and k.endswith('per_issue_active.ratio'):
            try:
                v = float(d[k])
            except ValueError:
                continue
            if v > 0.08:
                print(f"  stall {k.split('issue_stalled_')[1].split('_per_issue')[0]:40s} {v:.2f}")
