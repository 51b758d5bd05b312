#!/bin/bash
# full validation: GPU test suite, smoke(), default bench (all configs, CPU baseline), reference arm (short)
O=gpurun_out/${1:-full}; mkdir -p $O
S=$(date +%s)
timeout 1200 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1
echo "pytest rc=$? secs=$(( $(date +%s) - S ))" > $O/rc.txt
tail -4 $O/pytest_gpu.log
S=$(date +%s)
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
echo "smoke rc=$? secs=$(( $(date +%s) - S ))" >> $O/rc.txt
tail -4 $O/smoke.log
S=$(date +%s)
timeout 1500 python bench.py > $O/bench.json 2> $O/bench.err
echo "bench rc=$? secs=$(( $(date +%s) - S ))" >> $O/rc.txt
tail -c 300 $O/bench.err
python - <<P
import json
d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1])
print('bench', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'head', d['head_only']['ms_per_step'], d['head_only']['eager_median_ms'])
print('roofline', {k: d['roofline'][k] for k in ('kernel','frac','avg_launch_us','traffic')})
for k,v in d['roofline_other'].items(): print(' ', k, round(v['frac'],3), round(v['avg_launch_us'],1), v.get('launches_per_step'))
print('cpu', d['cpu_baseline'])
for k,v in d['other_configs'].items(): print(k, round(v['ms_per_step'],2), round(v['e2e']['ms_per_step'],2), round(v['head_only']['ms_per_step'],2), v['cpu_baseline'] and round(v['cpu_baseline']['value'],2))
print('fp32', d['fp32_mode']['ms_per_step'], 'forecast', d['v4_forecast']['ms_per_batch'])
print('clocks', d['clocks'])
P
cat $O/rc.txt
