#!/bin/bash
# stem convolution on tcgen05: tests, then headline bench with / without it
O=gpurun_out/r02w; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_trunk.py -q -x -k "stem" > $O/pytest_stem.log 2>&1
echo "stem pytest rc=$?" > $O/rc.txt
tail -25 $O/pytest_stem.log
timeout 600 python -m pytest tests/test_gpu_trunk.py tests/test_gpu_ddp_nccl.py -q > $O/pytest_trunk.log 2>&1
echo "trunk pytest rc=$?" >> $O/rc.txt
tail -5 $O/pytest_trunk.log
for v in 1 0; do
  V2F_STEM_CONV=$v timeout 600 python bench.py --only-headline --no-cpu-baseline --steps 20 --warmup 5 > $O/bench_stem$v.json 2> $O/bench_stem$v.err
  echo "bench stem=$v rc=$?" >> $O/rc.txt
  python - <<P
import json
d=json.loads(open('$O/bench_stem$v.json').read().strip().splitlines()[-1])
print('stem=$v', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['head_only']['ms_per_step'])
P
done
cat $O/rc.txt
