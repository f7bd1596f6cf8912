#!/bin/bash
# round 2, GPU session 11: whole GPU suite, IALS v2 after the first fixes + ncu of its user half epoch
export B200REC_SYNTH_CACHE=/dev/shm
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests -x -q -m gpu --durations=10 ) > $O/c11_suite.log 2>&1
echo "suite rc=$?" >> $O/c11_suite.log
( timeout 200 python tools/dev_ials_bench.py C4 256 2 ) > $O/c11_ials_v2_256.log 2>&1
( B200REC_IALS_V2=1 timeout 200 python tools/dev_ials_bench.py C4 128 2 ) > $O/c11_ials_v2_128.log 2>&1
( timeout 400 ncu --set full --clock-control none --import-source on -k regex:ials_rows_v2 -c 1 -f -o $O/prof_ials_v2_c4 python tools/dev_ials_bench.py C4 256 1 ) > $O/c11_ncu.log 2>&1
echo "ncu rc=$?" >> $O/c11_ncu.log
for f in $O/c11_*.log; do echo "== $f"; tail -n 14 $f; done
ls -la $O/*.ncu-rep
