#!/bin/bash
mkdir -p gpurun_out/r02n
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/head_step.py --steps 1 --batch 8 > gpurun_out/r02n/san.log 2>&1
echo "rc=$?"
grep -v "^=========     at\|^=========         Host Frame\|^=========                in" gpurun_out/r02n/san.log | head -60
