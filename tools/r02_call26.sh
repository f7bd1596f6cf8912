#!/bin/bash
# SLIM-BPR sequential kernel with the next sample / the adaptive state prefetched: parity, timing, phase counters
export B200REC_SYNTH_CACHE=/dev/shm
O=gpurun_out
mkdir -p $O
( timeout 200 python -m pytest tests/test_slim_gpu.py tests/test_next_rows_gpu.py -x -q -m gpu -k "not elasticnet and not asysvd" ) > $O/c26_tests.log 2>&1; rc=$?; echo "tests rc=$rc"
tail -n 8 $O/c26_tests.log | cut -c1-300
( timeout 100 python tools/dev_slim_bench.py ) > $O/c26_slim.jsonl 2>&1; echo "slim rc=$?"
cat $O/c26_slim.jsonl | cut -c1-300
( B200REC_SLIM_PROF=1 timeout 100 python tools/dev_slim_bench.py ) > $O/c26_slim_prof.txt 2>&1; echo "prof rc=$?"
grep "phase cycles" $O/c26_slim_prof.txt | sort | uniq -c | sort -rn | head -8 | cut -c1-300
