#!/bin/bash
O=gpurun_out/r02tg; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_rnn.py tests/test_gpu_fullsize.py -q -x > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -2 $O/pytest.log
timeout 400 python bench.py --only-headline --no-cpu-baseline --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err
python - <<P
import json
d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1])
print('bench', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['head_only']['ms_per_step'])
t=d['roofline_other']['tilegrad_kernel']; print('tilegrad', t['avg_launch_us'], t['frac'])
P
