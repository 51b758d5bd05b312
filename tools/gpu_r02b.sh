#!/bin/bash
# row-team decoder bring-up: targeted parity, then head timing
mkdir -p gpurun_out/r02d
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_rnn.py tests/test_gpu_data.py -q -x -k "bf16 or tensorcore or 64_rows or graph or persistent or device_normalized" > gpurun_out/r02d/pytest_team.log 2>&1
echo "team pytest rc=$?" > gpurun_out/r02d/rc.txt
tail -25 gpurun_out/r02d/pytest_team.log
V2F_TEAM_DECODE=0 timeout 600 python -m pytest tests/test_gpu_fullsize.py -q -x -k "bf16" > gpurun_out/r02d/pytest_noteam.log 2>&1
echo "no-team pytest rc=$?" >> gpurun_out/r02d/rc.txt
tail -3 gpurun_out/r02d/pytest_noteam.log
timeout 600 python bench.py --only-headline --no-cpu-baseline --steps 10 > gpurun_out/r02d/bench_headline.json 2> gpurun_out/r02d/bench.err
echo "bench rc=$?" >> gpurun_out/r02d/rc.txt
tail -c 600 gpurun_out/r02d/bench.err
timeout 900 python -m pytest tests/test_gpu_training_parity.py -q -s -k "gtm or v4 or default_dims" > gpurun_out/r02d/pytest_train.log 2>&1
echo "train pytest rc=$?" >> gpurun_out/r02d/rc.txt
grep -E "MAE|passed|failed" gpurun_out/r02d/pytest_train.log | tail -20
cat gpurun_out/r02d/rc.txt
