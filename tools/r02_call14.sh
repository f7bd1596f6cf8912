#!/bin/bash
# round 2, GPU session 14: K1-D v6 (atomic-free sweep) after the warp-uniform loop fix -- short timeouts
export B200REC_SYNTH_CACHE=/dev/shm
O=gpurun_out
mkdir -p $O
( time timeout 90 python -m pytest tests/test_similarity_gpu.py -x -q -m gpu -k "k1c" ) > $O/c14_k1c_tests.log 2>&1
rc=$?; echo "k1c rc=$rc" >> $O/c14_k1c_tests.log
if [ $rc -eq 0 ]; then
  ( timeout 100 python tools/dev_sim_bench.py C5 binary 4 ) > $O/c14_sim_c5.log 2>&1
  ( time timeout 120 python -m pytest tests/test_similarity_gpu.py tests/test_golden_gpu.py -x -q -m gpu ) > $O/c14_sim_tests.log 2>&1
  echo "sim rc=$?" >> $O/c14_sim_tests.log
  ( timeout 120 python -m pytest tests/test_scale_parity_gpu.py -x -q -m gpu -k "c5" ) > $O/c14_scale.log 2>&1
  echo "scale rc=$?" >> $O/c14_scale.log
fi
for f in $O/c14_*.log; do echo "== $f"; tail -n 8 $f; done
