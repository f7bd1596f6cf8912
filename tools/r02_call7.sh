#!/bin/bash
# round 2, GPU session 7: K1-D v3 (bank-spread layout, next-item prefetch), sharded SLIM fix
export B200REC_SYNTH_CACHE=/dev/shm
O=gpurun_out
mkdir -p $O
( time timeout 300 python -m pytest tests/test_similarity_gpu.py tests/test_golden_gpu.py -x -q -m gpu ) > $O/c7_sim_tests.log 2>&1
echo "sim rc=$?" >> $O/c7_sim_tests.log
( timeout 150 python tools/dev_sim_bench.py C5 binary 4 ) > $O/c7_sim_c5.log 2>&1
( time timeout 300 python -m pytest tests/test_slim_gpu.py -x -q -m gpu ) > $O/c7_slim_tests.log 2>&1
echo "slim rc=$?" >> $O/c7_slim_tests.log
( timeout 300 python -m pytest tests/test_scale_parity_gpu.py -x -q -m gpu -k "c5" ) > $O/c7_scale_tests.log 2>&1
echo "scale rc=$?" >> $O/c7_scale_tests.log
for f in $O/c7_*.log; do echo "== $f"; tail -n 10 $f; done
