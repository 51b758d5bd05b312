#!/bin/bash
# N GPUs: the DEFAULT bench line (all configs) under torchrun, as the driver's scaling run launches it
N=${2:-4}; O=gpurun_out/${1:-n4}; mkdir -p $O
S=$(date +%s)
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 20 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err
echo "bench n$N rc=$? secs=$(( $(date +%s) - S ))" > $O/rc.txt
tail -c 400 $O/bench_n$N.err
python - <<P
import json
d=json.loads(open('$O/bench_n$N.json').read().strip().splitlines()[-1])
print('n$N', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['head_only']['ms_per_step'])
for k,v in d['other_configs'].items(): print(k, round(v['value']), round(v['ms_per_step'],2), round(v['e2e']['ms_per_step'],2))
print('fp32', d['fp32_mode']['ms_per_step'], 'forecast', d['v4_forecast']['value'])
P
S=$(date +%s)
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $O/ref_n$N.json 2> $O/ref_n$N.err
echo "ref n$N rc=$? secs=$(( $(date +%s) - S ))" >> $O/rc.txt
cat $O/ref_n$N.json | cut -c1-300
cat $O/rc.txt
