#!/bin/bash
# round 2, GPU session 13: IALS v2 (slim diagonal block), K1-D v6 (atomic-free sweep)
export B200REC_SYNTH_CACHE=/dev/shm
O=gpurun_out
mkdir -p $O
( time timeout 300 python -m pytest tests/test_ials.py -x -q -m gpu ) > $O/c13_ials_tests.log 2>&1
echo "ials rc=$?" >> $O/c13_ials_tests.log
( timeout 200 python tools/dev_ials_bench.py C4 256 2 ) > $O/c13_ials_v2_256.log 2>&1
( B200REC_IALS_V2=1 timeout 200 python tools/dev_ials_bench.py C4 128 2 ) > $O/c13_ials_v2_128.log 2>&1
( time timeout 300 python -m pytest tests/test_similarity_gpu.py tests/test_golden_gpu.py -x -q -m gpu ) > $O/c13_sim_tests.log 2>&1
echo "sim rc=$?" >> $O/c13_sim_tests.log
( timeout 150 python tools/dev_sim_bench.py C5 binary 4 ) > $O/c13_sim_c5.log 2>&1
( timeout 300 python -m pytest tests/test_scale_parity_gpu.py -x -q -m gpu -k "c5 or c4" ) > $O/c13_scale.log 2>&1
echo "scale rc=$?" >> $O/c13_scale.log
for f in $O/c13_*.log; do echo "== $f"; tail -n 8 $f; done
