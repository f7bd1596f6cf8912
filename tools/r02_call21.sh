#!/bin/bash
# validation + timings of the SURVEY 8(f).4 trainers (tree-sparse SLIM, AsySVD, SLIM ElasticNet)
export B200REC_SYNTH_CACHE=/dev/shm
O=gpurun_out
mkdir -p $O
( timeout 420 python -m pytest tests/test_next_rows_gpu.py -q -m gpu --durations=8 ) > $O/c21_tests.log 2>&1; rc=$?; echo "tests rc=$rc"
tail -n 40 $O/c21_tests.log | cut -c1-400
( timeout 150 python -m pytest tests/test_slim_gpu.py tests/test_mf_gpu.py -x -q -m gpu ) > $O/c21_regress.log 2>&1; echo "regress rc=$?"
tail -n 4 $O/c21_regress.log | cut -c1-300
( timeout 360 python tools/next_rows_bench.py ) > $O/c21_next_rows.jsonl 2> $O/c21_next_rows.err; echo "bench rc=$?"
cat $O/c21_next_rows.jsonl | cut -c1-500; tail -n 5 $O/c21_next_rows.err | cut -c1-300
