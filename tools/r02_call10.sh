#!/bin/bash
# round 2, GPU session 10: IALS v2 (tensor-core Gram + blocked fp32 Cholesky + fp64 refinement), K1-D v2 restored, full bench
export B200REC_SYNTH_CACHE=/dev/shm
O=gpurun_out
mkdir -p $O
( time timeout 400 python -m pytest tests/test_ials.py -x -q -m gpu ) > $O/c10_ials_tests.log 2>&1
echo "ials rc=$?" >> $O/c10_ials_tests.log
( timeout 200 python -m pytest tests/test_scale_parity_gpu.py -x -q -m gpu -k "c4" ) > $O/c10_scale_c4.log 2>&1
echo "scale c4 rc=$?" >> $O/c10_scale_c4.log
( timeout 200 python tools/dev_ials_bench.py C4 256 2 ) > $O/c10_ials_v2_256.log 2>&1
( timeout 200 python tools/dev_ials_bench.py C4 128 2 ) > $O/c10_ials_v2_128.log 2>&1
( time timeout 200 python -m pytest tests/test_similarity_gpu.py -x -q -m gpu ) > $O/c10_sim_tests.log 2>&1
echo "sim rc=$?" >> $O/c10_sim_tests.log
( timeout 120 python tools/dev_mf_bench.py C5 128 3 ) > $O/c10_mf_c5.log 2>&1
( timeout 900 python bench.py --steps 5 --warmup 3 > $O/c10_bench_n1.json ) 2> $O/c10_bench_n1.err
echo "bench rc=$?" >> $O/c10_bench_n1.err
for f in $O/c10_*.log; do echo "== $f"; tail -n 8 $f; done
tail -c 3000 $O/c10_bench_n1.json; echo; tail -n 15 $O/c10_bench_n1.err
