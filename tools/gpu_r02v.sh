#!/bin/bash
# 2 GPUs: NCCL tests (fail-fast workers), then the DEFAULT bench under torchrun at N=2 (all configs), timed
O=gpurun_out/r02v; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_ddp_nccl.py -q -s -x > $O/pytest_nccl.log 2>&1
echo "nccl pytest rc=$?" > $O/rc.txt
grep -E "^[01] \{|passed|failed|Error" $O/pytest_nccl.log | cut -c1-1500 | tail -12
S=$(date +%s)
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err
echo "bench n2 rc=$? secs=$(( $(date +%s) - S ))" >> $O/rc.txt
tail -c 600 $O/bench_n2.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02v/bench_n2.json').read().strip().splitlines()[-1])
print('n2', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['head_only'])
for k,v in d['other_configs'].items(): print(k, v['ms_per_step'], v['e2e']['ms_per_step'], v['head_only'])
print(d['fp32_mode']); print(d['v4_forecast'])
P
cat $O/rc.txt
