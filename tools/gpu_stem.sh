#!/bin/bash
O=gpurun_out/${1:-stem}; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_trunk.py -q -x -k "stem" > $O/pytest_stem.log 2>&1
echo "stem pytest rc=$?" > $O/rc.txt
tail -5 $O/pytest_stem.log
timeout 300 python tools/stem_probe.py > $O/stem_probe.txt 2>&1; cat $O/stem_probe.txt
timeout 600 python bench.py --only-headline --no-cpu-baseline --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err
echo "bench rc=$?" >> $O/rc.txt
python - <<P
import json
d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1])
print('bench', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['head_only']['ms_per_step'])
P
cat $O/rc.txt
