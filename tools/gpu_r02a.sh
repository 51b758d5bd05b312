#!/bin/bash
# first GPU pass of round 2: tests, smoke, bench (1 GPU)
mkdir -p gpurun_out/r02a
python -m pytest tests -m gpu -q --maxfail=30 --deselect tests/test_gpu_ddp_nccl.py > gpurun_out/r02a/pytest.log 2>&1
echo "pytest rc=$?" > gpurun_out/r02a/rc.txt
tail -40 gpurun_out/r02a/pytest.log
python __graft_entry__.py --smoke > gpurun_out/r02a/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r02a/rc.txt
tail -5 gpurun_out/r02a/smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02a/bench.json 2> gpurun_out/r02a/bench.err
echo "bench rc=$?" >> gpurun_out/r02a/rc.txt
tail -c 800 gpurun_out/r02a/bench.err
cat gpurun_out/r02a/rc.txt
