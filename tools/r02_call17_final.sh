#!/bin/bash
# round 2, final single-GPU session: smoke, whole suite, bench (both arms), launch list + ncu captures for profiles/
export B200REC_SYNTH_CACHE=/dev/shm
O=gpurun_out
mkdir -p $O
( timeout 300 python __graft_entry__.py smoke ) > $O/c17_smoke.log 2>&1; echo "smoke rc=$?" >> $O/c17_smoke.log
( time timeout 600 python -m pytest tests -x -q -m gpu --durations=6 ) > $O/c17_suite.log 2>&1; echo "suite rc=$?" >> $O/c17_suite.log
( timeout 900 python bench.py --steps 10 --warmup 3 > $O/c17_bench_n1.json ) 2> $O/c17_bench_n1.err; echo "bench rc=$?" >> $O/c17_bench_n1.err
( timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/c17_bench_ref.json ) 2> $O/c17_bench_ref.err; echo "ref rc=$?" >> $O/c17_bench_ref.err
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/c17_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-tensor --e2e-steps 3 ) > $O/c17_launchlist.log 2>&1; echo "launch list rc=$?" >> $O/c17_launchlist.log
( timeout 200 ncu --set full --clock-control none --import-source on -k regex:sim_k1d_kernel -c 1 -f -o $O/prof_k1d_final python tools/dev_sim_bench.py C5 binary 1 ) > $O/c17_ncu_k1d.log 2>&1; echo "ncu k1d rc=$?" >> $O/c17_ncu_k1d.log
( timeout 200 ncu --set full --clock-control none -k regex:tc2_gemm --launch-skip 8 -c 1 -f -o $O/prof_gemm2 python tools/dev_gemm_bench.py ) > $O/c17_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?" >> $O/c17_ncu_gemm.log
( timeout 200 ncu --set full --clock-control none -k regex:mf_dataflow -c 1 -f -o $O/prof_mf_dataflow python tools/dev_mf_bench.py C5 128 1 ) > $O/c17_ncu_mf.log 2>&1; echo "ncu mf rc=$?" >> $O/c17_ncu_mf.log
( timeout 300 ncu --section SourceCounters --section WarpStateStats --section SpeedOfLight --section LaunchStats --section Occupancy --clock-control none --import-source on -k regex:ials_rows_v2 -c 1 -f -o $O/prof_ials_v2_final python tools/dev_ials_bench.py C4 256 1 ) > $O/c17_ncu_ials.log 2>&1; echo "ncu ials rc=$?" >> $O/c17_ncu_ials.log
for f in $O/c17_*.log; do echo "== $f"; tail -n 5 $f; done
tail -c 600 $O/c17_bench_n1.json; echo; tail -c 400 $O/c17_bench_ref.json; echo; ls -la $O/*.ncu-rep
