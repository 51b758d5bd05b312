#!/bin/bash
# 2 GPUs: NCCL tests, overlap timeline, headline bench at N=2.  Hard timeouts: a hang costs 2x GPU-minutes.
O=gpurun_out/${1:-n2}; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_ddp_nccl.py -q -x > $O/pytest_nccl.log 2>&1
echo "nccl pytest rc=$?" > $O/rc.txt; tail -3 $O/pytest_nccl.log | cut -c1-300
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/overlap_trace.py > $O/nccl_overlap.md 2> $O/overlap.err
echo "overlap rc=$?" >> $O/rc.txt; head -12 $O/nccl_overlap.md | cut -c1-200
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --only-headline --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err
echo "bench n2 rc=$?" >> $O/rc.txt
python - <<P
import json
d=json.loads(open('$O/bench_n2.json').read().strip().splitlines()[-1])
print('n2', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['head_only']['ms_per_step'])
P
cat $O/rc.txt
