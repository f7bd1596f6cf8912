#!/bin/bash
# AsySVD kernel v4 (float4, several rows per warp instruction) + SLIM ElasticNet kernel v2 (3 barriers per step, row prefetch)
export B200REC_SYNTH_CACHE=/dev/shm
O=gpurun_out
mkdir -p $O
( timeout 300 python -m pytest tests/test_next_rows_gpu.py -x -q -m gpu ) > $O/c25_tests.log 2>&1; rc=$?; echo "tests rc=$rc"
tail -n 12 $O/c25_tests.log | cut -c1-300
if [ $rc -eq 0 ]; then
( timeout 150 python tools/next_rows_bench.py --only-asy ) > $O/c25_asy.jsonl 2> $O/c25_asy.err; echo "asy rc=$?"
cat $O/c25_asy.jsonl | cut -c1-400; tail -n 3 $O/c25_asy.err | cut -c1-300
( B200REC_ASY_PROF=1 timeout 150 python tools/next_rows_bench.py --only-asy ) > $O/c25_asy_prof.txt 2>&1; echo "prof rc=$?"
grep "phase cycles" $O/c25_asy_prof.txt | cut -c1-300
( timeout 200 python tools/next_rows_bench.py --no-c4 ) > $O/c25_next_rows.jsonl 2> $O/c25_next_rows.err; echo "bench rc=$?"
cat $O/c25_next_rows.jsonl | cut -c1-500; tail -n 3 $O/c25_next_rows.err | cut -c1-300
fi
