#!/bin/bash
export B200REC_SYNTH_CACHE=/dev/shm
O=gpurun_out
mkdir -p $O
( timeout 400 python -m pytest tests/test_multi_gpu.py -x -q -m gpu ) > $O/c18_pytest_mgpu.log 2>&1; echo "pytest rc=$?" >> $O/c18_pytest_mgpu.log
tail -n 12 $O/c18_pytest_mgpu.log
