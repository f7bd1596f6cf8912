#!/bin/bash
# round 2, third GPU session: nibble-counter kernel K1-D (parity + C5 timing / phases), dataflow mini-batch MF kernel
export B200REC_SYNTH_CACHE=/dev/shm
O=gpurun_out
mkdir -p $O
( time timeout 420 python -m pytest tests/test_similarity_gpu.py -x -q -m gpu --durations=8 ) > $O/c3_sim_tests.log 2>&1
echo "sim rc=$?" >> $O/c3_sim_tests.log
( timeout 200 python tools/dev_sim_bench.py C5 binary 3 ) > $O/c3_sim_c5.log 2>&1
( time timeout 300 python -m pytest tests/test_mf_gpu.py -x -q -m gpu --durations=5 ) > $O/c3_mf_tests.log 2>&1
echo "mf rc=$?" >> $O/c3_mf_tests.log
( timeout 200 python tools/dev_mf_bench.py C5 ) > $O/c3_mf_c5.log 2>&1
( B200REC_MF_DATAFLOW=0 timeout 200 python tools/dev_mf_bench.py C5 ) > $O/c3_mf_c5_coop.log 2>&1
for f in $O/c3_*.log; do echo "== $f"; tail -n 12 $f; done
