#!/bin/bash
O=gpurun_out/${1:-nccl}; mkdir -p $O
export V2F_NCCL_TEST_TRACE=$PWD/$O
timeout 200 python -m pytest tests/test_gpu_ddp_nccl.py -q -x -k sharded > $O/pytest_nccl.log 2>&1
echo "nccl pytest rc=$?" > $O/rc.txt; tail -40 $O/pytest_nccl.log | cut -c1-400
cat $O/nccl_test_rank*.log; cat $O/rc.txt
