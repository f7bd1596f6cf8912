#!/bin/bash
# round 2, GPU session 12: IALS v2 with the inverse off the critical path
export B200REC_SYNTH_CACHE=/dev/shm
O=gpurun_out
mkdir -p $O
( time timeout 300 python -m pytest tests/test_ials.py -x -q -m gpu ) > $O/c12_ials_tests.log 2>&1
echo "ials rc=$?" >> $O/c12_ials_tests.log
( timeout 200 python -m pytest tests/test_scale_parity_gpu.py -x -q -m gpu -k "c4" ) > $O/c12_scale_c4.log 2>&1
echo "scale c4 rc=$?" >> $O/c12_scale_c4.log
( timeout 200 python tools/dev_ials_bench.py C4 256 2 ) > $O/c12_ials_v2_256.log 2>&1
( B200REC_IALS_V2=1 timeout 200 python tools/dev_ials_bench.py C4 128 2 ) > $O/c12_ials_v2_128.log 2>&1
for f in $O/c12_*.log; do echo "== $f"; tail -n 6 $f; done
