#!/bin/bash
# round 2, GPU session 9: K1-D v4 (prefetch), device glibc replay, sharded SLIM
export B200REC_SYNTH_CACHE=/dev/shm
O=gpurun_out
mkdir -p $O
( time timeout 300 python -m pytest tests/test_similarity_gpu.py tests/test_golden_gpu.py -x -q -m gpu ) > $O/c9_sim_tests.log 2>&1
echo "sim rc=$?" >> $O/c9_sim_tests.log
( timeout 150 python tools/dev_sim_bench.py C5 binary 4 ) > $O/c9_sim_c5.log 2>&1
( time timeout 300 python -m pytest tests/test_mf_gpu.py tests/test_slim_gpu.py -x -q -m gpu ) > $O/c9_mf_slim_tests.log 2>&1
echo "mf/slim rc=$?" >> $O/c9_mf_slim_tests.log
( timeout 200 python tools/dev_mf_bench.py C5 128 3 ) > $O/c9_mf_c5.log 2>&1
( timeout 300 python -m pytest tests/test_scale_parity_gpu.py -x -q -m gpu -k "c5 or c3" ) > $O/c9_scale_tests.log 2>&1
echo "scale rc=$?" >> $O/c9_scale_tests.log
for f in $O/c9_*.log; do echo "== $f"; tail -n 10 $f; done
