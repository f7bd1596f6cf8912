#!/bin/bash
# AsySVD kernel v3 (shared-memory stash of the profile rows + cp.async state prefetch): parity + timing
export B200REC_SYNTH_CACHE=/dev/shm
O=gpurun_out
mkdir -p $O
( timeout 200 python -m pytest tests/test_next_rows_gpu.py -x -q -m gpu -k "asysvd" ) > $O/c23_tests.log 2>&1; rc=$?; echo "tests rc=$rc"
tail -n 12 $O/c23_tests.log | cut -c1-300
if [ $rc -eq 0 ]; then
( timeout 200 python tools/next_rows_bench.py --only-asy ) > $O/c23_asy.jsonl 2> $O/c23_asy.err; echo "asy rc=$?"
cat $O/c23_asy.jsonl | cut -c1-400; tail -n 3 $O/c23_asy.err | cut -c1-300
fi
