#!/bin/bash
# N=2 headline under different NCCL settings (protocol / CTA count): which one leaves the backward alone
O=gpurun_out/${1:-ncclenv}; mkdir -p $O
i=0
for envs in "X=1" "NCCL_PROTO=Simple" "NCCL_PROTO=Simple NCCL_MAX_CTAS=4" "NCCL_MAX_CTAS=4" "NCCL_PROTO=Simple NCCL_MAX_CTAS=8 NCCL_MIN_CTAS=8"; do
  i=$((i+1))
  env $envs timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29520+i)) bench.py --gpus 2 --only-headline --no-cpu-baseline --steps 20 --warmup 5 > $O/bench_$i.json 2> $O/bench_$i.err
  rc=$?
  python - <<P
import json
try:
    d=json.loads(open('$O/bench_$i.json').read().strip().splitlines()[-1])
    print('$envs', 'rc=$rc', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3))
except Exception as e:
    print('$envs', 'rc=$rc', 'ERR', e)
P
done
