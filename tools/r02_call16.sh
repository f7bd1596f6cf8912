#!/bin/bash
# round 2, GPU session 16: IALS v2 with 16 warps per CTA
export B200REC_SYNTH_CACHE=/dev/shm
O=gpurun_out
mkdir -p $O
( time timeout 200 python -m pytest tests/test_ials.py -x -q -m gpu ) > $O/c16_ials_tests.log 2>&1
rc=$?; echo "ials rc=$rc" >> $O/c16_ials_tests.log
if [ $rc -eq 0 ]; then
  ( timeout 150 python -m pytest tests/test_scale_parity_gpu.py -x -q -m gpu -k "c4" ) > $O/c16_scale_c4.log 2>&1
  echo "scale c4 rc=$?" >> $O/c16_scale_c4.log
  ( timeout 150 python tools/dev_ials_bench.py C4 256 2 ) > $O/c16_ials_v2_256.log 2>&1
  ( timeout 150 python tools/dev_ials_bench.py C4 128 2 ) > $O/c16_ials_v2_128.log 2>&1
fi
for f in $O/c16_*.log; do echo "== $f"; tail -n 6 $f; done
