#!/bin/bash
# round 2, fourth GPU session: K1-D after the race fix (tests, C5 timing + phases, ncu), scale parity tests, hot path (iii) timings
export B200REC_SYNTH_CACHE=/dev/shm
O=gpurun_out
mkdir -p $O
( time timeout 300 python -m pytest tests/test_similarity_gpu.py -x -q -m gpu ) > $O/c4_sim_tests.log 2>&1
echo "sim rc=$?" >> $O/c4_sim_tests.log
( timeout 150 python tools/dev_sim_bench.py C5 binary 4 ) > $O/c4_sim_c5.log 2>&1
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:sim_k1d_kernel -c 1 -f -o $O/prof_k1d_c5 python tools/dev_sim_bench.py C5 binary 1 ) > $O/c4_ncu.log 2>&1
echo "ncu rc=$?" >> $O/c4_ncu.log
( time timeout 600 python -m pytest tests/test_scale_parity_gpu.py -x -q -m gpu --durations=8 ) > $O/c4_scale_tests.log 2>&1
echo "scale rc=$?" >> $O/c4_scale_tests.log
( B200REC_IALS_TC=1 timeout 400 python -m pytest tests/test_ials.py -q -m gpu ) > $O/c4_ialstc_tests.log 2>&1
echo "ialstc rc=$?" >> $O/c4_ialstc_tests.log
( timeout 200 python tools/dev_gemm_bench.py ) > $O/c4_gemm_bench.log 2>&1
( timeout 200 python tools/dev_ease_bench.py C4 ) > $O/c4_ease_c4.log 2>&1
( timeout 200 python tools/dev_ials_bench.py C4 128 2 ) > $O/c4_ials_fp64_128.log 2>&1
( B200REC_IALS_TC=1 timeout 200 python tools/dev_ials_bench.py C4 128 2 ) > $O/c4_ials_tc_128.log 2>&1
( timeout 200 python tools/dev_ials_bench.py C4 256 1 ) > $O/c4_ials_fp64_256.log 2>&1
for f in $O/c4_*.log; do echo "== $f"; tail -n 8 $f; done
ls -la $O
