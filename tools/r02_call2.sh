#!/bin/bash
# round 2, second GPU session: suite with the bitmap kernel (K1-C) + its C5 timing / phases
export B200REC_SYNTH_CACHE=/dev/shm
O=gpurun_out
mkdir -p $O
( time timeout 300 python -m pytest tests/test_similarity_gpu.py -x -q -m gpu --durations=8 ) > $O/c2_sim_tests.log 2>&1
echo "sim rc=$?" >> $O/c2_sim_tests.log
( timeout 300 python tools/dev_sim_bench.py C5 binary 3 ) > $O/c2_sim_c5.log 2>&1
( time timeout 900 python -m pytest tests -x -q -m gpu --durations=8 ) > $O/c2_tests.log 2>&1
echo "suite rc=$?" >> $O/c2_tests.log
( B200REC_K1C=0 timeout 300 python tools/dev_sim_bench.py C5 binary 2 ) > $O/c2_sim_c5_window.log 2>&1
for f in $O/c2_*.log; do echo "== $f"; tail -n 4 $f; done
