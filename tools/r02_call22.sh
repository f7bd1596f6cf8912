#!/bin/bash
# AsySVD kernel v2 (32 warps, 4 rows in flight, next sample prefetched): parity + timing; regressions of the files touched
export B200REC_SYNTH_CACHE=/dev/shm
O=gpurun_out
mkdir -p $O
( timeout 300 python -m pytest tests/test_next_rows_gpu.py tests/test_slim_gpu.py -x -q -m gpu ) > $O/c22_tests.log 2>&1; echo "tests rc=$?"
tail -n 6 $O/c22_tests.log | cut -c1-300
( timeout 240 python tools/next_rows_bench.py --only-asy ) > $O/c22_asy.jsonl 2> $O/c22_asy.err; echo "asy rc=$?"
cat $O/c22_asy.jsonl | cut -c1-400; tail -n 3 $O/c22_asy.err | cut -c1-300
