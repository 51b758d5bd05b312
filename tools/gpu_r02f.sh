#!/bin/bash
# 2 GPUs: NCCL gradient equality test; bench at N=2 with the all-reduce inside / after the graph
mkdir -p gpurun_out/r02f
timeout 900 python -m pytest tests/test_gpu_ddp_nccl.py -q -s > gpurun_out/r02f/pytest_nccl.log 2>&1
echo "nccl pytest rc=$?" > gpurun_out/r02f/rc.txt
tail -15 gpurun_out/r02f/pytest_nccl.log
for mode in "" "--no-graph-nccl"; do
  tag=$( [ -z "$mode" ] && echo ingraph || echo after )
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --only-headline $mode > gpurun_out/r02f/bench_n2_$tag.json 2> gpurun_out/r02f/bench_n2_$tag.err
  echo "bench $tag rc=$?" >> gpurun_out/r02f/rc.txt
  tail -c 400 gpurun_out/r02f/bench_n2_$tag.err
  python -c "
import json,sys
d=json.load(open('gpurun_out/r02f/bench_n2_$tag.json'))
print('$tag', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])
"
done
timeout 600 python bench.py --only-headline --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r02f/bench_n1.json 2> gpurun_out/r02f/bench_n1.err
python -c "
import json
d=json.load(open('gpurun_out/r02f/bench_n1.json'))
print('n1', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])
"
cat gpurun_out/r02f/rc.txt
