#!/bin/bash
mkdir -p gpurun_out/r02t
for dbg in 0 3; do
V2F_TEAM_DEBUG=$dbg timeout 600 python bench.py --only-headline --no-cpu-baseline --steps 5 > gpurun_out/r02t/bench_dbg$dbg.json 2> gpurun_out/r02t/bench_dbg$dbg.err
python - <<PY
import json
d=json.load(open('gpurun_out/r02t/bench_dbg$dbg.json'))
r=d['roofline']; print('dbg $dbg', r['avg_launch_us'], {k:v['work'] for k,v in r['phases_us_per_step'].items() if k.startswith('P2')})
PY
done
