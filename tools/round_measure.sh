#!/bin/bash
# One GPU-box session that produces everything profiles/ cites for a round (run under gpurun from the repo root):
#   tools/round_measure.sh [LIBVARIANT]     (optional: name of a _variants/ library to measure instead of the default)
# Outputs under gpurun_out/: bench_n1.json, bench_ref.json, launches.csv (ncu launch list of the bench command),
# prof_sim_c5.ncu-rep (ncu --set full of one K1 launch).  Numbers printed under ncu are never bench values.
set -u
export B200REC_SYNTH_CACHE=/dev/shm
if [ "${1:-}" != "" ]; then export B200REC_LIB=$PWD/recsys2019_deeplearning_evaluation_b200/_variants/libb200rec_$1.so; fi
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$?"; cat gpurun_out/bench_n1.json | cut -c1-600
if [ "${SKIP_REF:-0}" != "1" ]; then
  python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
  echo "reference rc=$?"; cat gpurun_out/bench_ref.json | cut -c1-300
fi
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/b_ncu.log 2>&1
echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sim_topk_kernel -c 1 -f -o gpurun_out/prof_sim_c5 \
  python tools/dev_sim_bench.py C5 binary 1 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/
