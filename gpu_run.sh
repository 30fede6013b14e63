#!/bin/bash
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; tail -2 gpurun_out/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json')); r=json.load(open('gpurun_out/bench_ref.json'))
print('value %.3e ms/step %.3f frac %.3f | e2e %.3e | cpu port %.3e (%d cores) ref1 %.3e | ref arm %.3e kind %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['cpu_baseline']['reference_1core']['value'], r['value'], r['cpu_baseline']['kind']))
print('sustained frac %.3f, latency p50 %.1f us, ring %.1f GB/s, frontend corr %.2f ms' % (d['sustained']['frac'], d['latency']['p50_us'], d['ring_stream']['h2d_gbs'], d['frontend']['correlate_ms']))
print({k: round(v['frac_of_hbm_peak'],3) for k,v in d['other_configs'].items()})
PY
