#!/bin/bash
# team decoder forward + backward parity; ncu full capture of both
mkdir -p gpurun_out/r02e
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_rnn.py -q -x -k "bf16 or tensorcore or 64_rows or graph or persistent" > gpurun_out/r02e/pytest_team.log 2>&1
echo "team pytest rc=$?" > gpurun_out/r02e/rc.txt
tail -25 gpurun_out/r02e/pytest_team.log
timeout 600 python bench.py --only-headline --no-cpu-baseline --steps 10 > gpurun_out/r02e/bench_headline.json 2> gpurun_out/r02e/bench.err
echo "bench rc=$?" >> gpurun_out/r02e/rc.txt
tail -c 600 gpurun_out/r02e/bench.err
timeout 900 python -m pytest tests/test_gpu_training_parity.py -q -s -k "default_dims" > gpurun_out/r02e/pytest_train.log 2>&1
echo "train pytest rc=$?" >> gpurun_out/r02e/rc.txt
grep -E "MAE|passed|failed" gpurun_out/r02e/pytest_train.log | tail -8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_team -s 2 -c 2 -o gpurun_out/r02e/team_full python tools/head_step.py --steps 2 > gpurun_out/r02e/ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/r02e/rc.txt
tail -3 gpurun_out/r02e/ncu.log
cat gpurun_out/r02e/rc.txt
