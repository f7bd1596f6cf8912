#!/bin/bash
# round 2, first GPU session: full suite, then the three opt-in kernels (parity + timing)
export B200REC_SYNTH_CACHE=/dev/shm
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/c1_smi.txt
( time timeout 900 python -m pytest tests -x -q -m gpu --durations=12 ) > $O/c1_tests.log 2>&1
echo "suite rc=$?" >> $O/c1_tests.log
( B200REC_K1B=1 timeout 600 python -m pytest tests/test_similarity_gpu.py tests/test_golden_gpu.py tests/test_z_similarity_extra_gpu.py -q -m gpu ) > $O/c1_k1b_tests.log 2>&1
echo "k1b rc=$?" >> $O/c1_k1b_tests.log
( timeout 400 python tools/dev_sim_bench.py C5 binary 3 ) > $O/c1_sim_default.log 2>&1
( B200REC_K1B=1 timeout 400 python tools/dev_sim_bench.py C5 binary 3 ) > $O/c1_sim_k1b.log 2>&1
( B200REC_TEST_GEMM2=1 B200REC_GEMM=2 timeout 600 python -m pytest tests/test_ease_gpu.py -q -m gpu ) > $O/c1_gemm2_tests.log 2>&1
echo "gemm2 rc=$?" >> $O/c1_gemm2_tests.log
( timeout 300 python tools/dev_gemm_bench.py ) > $O/c1_gemm_bench.log 2>&1
( B200REC_IALS_TC=1 timeout 600 python -m pytest tests/test_ials.py -q -m gpu ) > $O/c1_ialstc_tests.log 2>&1
echo "ialstc rc=$?" >> $O/c1_ialstc_tests.log
( timeout 300 python tools/dev_ials_bench.py C4 128 2 ) > $O/c1_ials_fp64.log 2>&1
( B200REC_IALS_TC=1 timeout 300 python tools/dev_ials_bench.py C4 128 2 ) > $O/c1_ials_tc.log 2>&1
tail -3 $O/c1_*.log
