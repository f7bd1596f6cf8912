#!/bin/bash
# AsySVD kernel v2+ (lane-held ids, two-ahead prefetch, Adam powers on thread 0) with phase timers
export B200REC_SYNTH_CACHE=/dev/shm
O=gpurun_out
mkdir -p $O
( timeout 200 python -m pytest tests/test_next_rows_gpu.py -x -q -m gpu -k "asysvd" ) > $O/c24_tests.log 2>&1; rc=$?; echo "tests rc=$rc"
tail -n 12 $O/c24_tests.log | cut -c1-300
if [ $rc -eq 0 ]; then
( timeout 200 python tools/next_rows_bench.py --only-asy ) > $O/c24_asy.jsonl 2> $O/c24_asy.err; echo "asy rc=$?"
cat $O/c24_asy.jsonl | cut -c1-400; tail -n 3 $O/c24_asy.err | cut -c1-300
( B200REC_ASY_PROF=1 timeout 200 python tools/next_rows_bench.py --only-asy ) > $O/c24_asy_prof.txt 2>&1; echo "prof rc=$?"
grep -B1 "phase cycles" $O/c24_asy_prof.txt | cut -c1-300
fi
