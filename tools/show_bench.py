"""tools/show_bench.py <bench.json> -- the figures of a bench.py line that the design notes quote."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = lambda v, n=3: round(v, n) if isinstance(v, float) else v
print("value %.4g  ms/step %.3f  roofline frac %.3f (whole path %.0f GB/s)  launches %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["whole_path_achieved"], d["gpu_launches"]))
print("clocks", d["clocks"])
if d.get("sustained"):
    print("sustained frac %.3f value %.4g clocks %s" % (d["sustained"]["frac"], d["sustained"]["value"], d["sustained"]["clocks"]))
if d.get("e2e"):
    print("e2e", {k: r(v) for k, v in d["e2e"].items() if k not in ("api", "h2d_copy", "h2d_ceiling_what")})
if d.get("cpu_baseline"):
    c = d["cpu_baseline"]
    print("cpu port %.4g (%d cores)  strong %.4g  reference 1 core %s" % (c["value"], c["cores"], c["strong"]["value"], c["reference_1core"] and "%.4g" % c["reference_1core"]["value"]))
if d.get("scaling_c4"):
    print("c4", {k: r(v) for k, v in d["scaling_c4"].items() if k not in ("plan", "workload")})
if d.get("other_configs"):
    print("others", {k: (r(v["frac_of_hbm_peak"]), r(v["ms"])) for k, v in d["other_configs"].items()})
if d.get("latency"):
    print("latency p50/p99", d["latency"].get("p50_us"), d["latency"].get("p99_us"), "in place", d["latency"].get("host_buffers_in_place"))
if d.get("ring_stream"):
    rs = d["ring_stream"]
    print("ring c3", {k: r(v) for k, v in rs.items() if k not in ("workload", "note", "c1")}, "c1", {k: r(v) for k, v in rs.get("c1", {}).items() if k not in ("workload", "note")})
print("parity", d.get("parity"), "affinity", d.get("affinity"))
