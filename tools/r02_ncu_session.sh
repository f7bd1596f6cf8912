#!/bin/bash
# ncu captures of the two kernels reworked last (numbers under the profiler are not bench values)
export B200REC_SYNTH_CACHE=/dev/shm
O=gpurun_out
mkdir -p $O
( timeout 140 ncu --set full --clock-control none --import-source on -k regex:slim_enet_kernel -c 1 -f -o $O/prof_enet_c2 python tools/dev_enet_bench.py ) > $O/ncu_ncu_enet.log 2>&1; echo "ncu enet rc=$?"
( timeout 100 ncu --set full --clock-control none --import-source on -k regex:slim_sequential_kernel --launch-skip 1 -c 1 -f -o $O/prof_slim_seq_c2 python tools/dev_slim_bench.py ) > $O/ncu_ncu_slim.log 2>&1; echo "ncu slim rc=$?"
tail -n 3 $O/ncu_ncu_enet.log $O/ncu_ncu_slim.log | cut -c1-200; ls -la $O/prof_enet_c2.ncu-rep $O/prof_slim_seq_c2.ncu-rep
