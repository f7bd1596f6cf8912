#!/bin/bash
# round 2, last single-GPU session: smoke, the whole GPU suite under -x, the default bench line
export B200REC_SYNTH_CACHE=/dev/shm
O=gpurun_out
mkdir -p $O
( time timeout 480 python -m pytest tests -x -q -m gpu --durations=6 ) > $O/final_suite.log 2>&1; echo "suite rc=$?" >> $O/final_suite.log
( timeout 120 python __graft_entry__.py smoke ) > $O/final_smoke.log 2>&1; echo "smoke rc=$?" >> $O/final_smoke.log
( timeout 300 python bench.py --no-tensor > $O/final_bench_n1.json ) 2> $O/final_bench_n1.err; echo "bench rc=$?" >> $O/final_bench_n1.err
for f in $O/final_*.log; do echo "== $f"; tail -n 6 $f | cut -c1-300; done
tail -n 2 $O/final_bench_n1.err | cut -c1-300; tail -c 700 $O/final_bench_n1.json; echo
