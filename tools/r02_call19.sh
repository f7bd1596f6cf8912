#!/bin/bash
# round 2, GPU session 19: C5 sweep on one GPU (configs[4]), check of the clean-ups, bench step timing with a reused output table
export B200REC_SYNTH_CACHE=/dev/shm
O=gpurun_out
mkdir -p $O
( time timeout 200 python -m pytest tests/test_ials.py tests/test_similarity_gpu.py -x -q -m gpu ) > $O/c19_tests.log 2>&1; echo "tests rc=$?" >> $O/c19_tests.log
( timeout 500 python tools/c5_sweep.py ) > $O/c19_c5_sweep.jsonl 2> $O/c19_c5_sweep.err; echo "sweep rc=$?" >> $O/c19_c5_sweep.err
( timeout 300 python bench.py --steps 10 --warmup 3 --no-tensor --no-bpr --no-cpu-baseline > $O/c19_bench_quick.json ) 2> $O/c19_bench_quick.err; echo "bench rc=$?" >> $O/c19_bench_quick.err
tail -n 5 $O/c19_tests.log; cat $O/c19_c5_sweep.jsonl | cut -c1-330; tail -n 5 $O/c19_c5_sweep.err; head -c 700 $O/c19_bench_quick.json; echo; tail -3 $O/c19_bench_quick.err
