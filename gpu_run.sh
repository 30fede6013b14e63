#!/bin/bash
mkdir -p gpurun_out
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value %.3e  ms/step %.3f  frac %.3f  achieved %.0f GB/s  e2e %.3e  cpu %.3e' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['achieved'], d['e2e']['value'], d['cpu_baseline']['value']))
print('clocks', d['clocks'])
print('sustained', d['sustained'])
PY
