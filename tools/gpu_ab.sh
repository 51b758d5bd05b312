#!/bin/bash
# A/B of one environment switch on the headline bench after the trunk / rnn tests: tools/gpu_ab.sh <outdir> <VAR>
O=gpurun_out/${1:-ab}; VAR=${2:-V2F_BN_PDL}; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_trunk.py tests/test_gpu_rnn.py -q -x > $O/pytest.log 2>&1
echo "pytest rc=$?" > $O/rc.txt; tail -3 $O/pytest.log
for v in 1 0; do
  env $VAR=$v timeout 600 python bench.py --only-headline --no-cpu-baseline --steps 20 --warmup 5 > $O/bench_$v.json 2> $O/bench_$v.err
  echo "bench $VAR=$v rc=$?" >> $O/rc.txt
  python - <<P
import json
d=json.loads(open('$O/bench_$v.json').read().strip().splitlines()[-1])
print('$VAR=$v', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['head_only']['ms_per_step'], 'gemm_tc', d['roofline_other'].get('gemm_tc_kernel',{}).get('achieved'), d['roofline_other'].get('gemm_tc_kernel',{}).get('ms_per_step'))
P
done
cat $O/rc.txt
