#!/bin/bash
# team decoder: parity subset + headline bench.  usage: bash tools/gpu_team.sh <tag>
tag=${1:-r02x}
mkdir -p gpurun_out/$tag
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_rnn.py -q -x -k "bf16 or tensorcore or 64_rows or graph or persistent" > gpurun_out/$tag/pytest_team.log 2>&1
echo "team pytest rc=$?" > gpurun_out/$tag/rc.txt
tail -8 gpurun_out/$tag/pytest_team.log
timeout 600 python bench.py --only-headline --no-cpu-baseline --steps 10 > gpurun_out/$tag/bench_headline.json 2> gpurun_out/$tag/bench.err
echo "bench rc=$?" >> gpurun_out/$tag/rc.txt
tail -c 600 gpurun_out/$tag/bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/$tag/bench_headline.json'))
print({k:d[k] for k in ['value','ms_per_step','gpu_launches_per_step']}, 'e2e', d['e2e']['ms_per_step'], 'head', d['head_only']['ms_per_step'])
r=d['roofline']; print(r['kernel'], r['frac'], r['avg_launch_us'])
for k,v in r.get('phases_us_per_step',{}).items(): print('  ',k,v)
for k,v in d['roofline_other'].items(): print(k, round(v['frac'],3), round(v['avg_launch_us'],1), v.get('launches_per_step'))
PY
cat gpurun_out/$tag/rc.txt
