#!/bin/bash
# round 2, second 2-GPU session: sharded SLIM / IALS / EASE checks, 2-GPU bench with the C3 legs
export B200REC_SYNTH_CACHE=/dev/shm
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631"
( timeout 200 $TR tools/mgpu_slim_check.py ) > $O/c15_slim.log 2>&1; echo "slim rc=$?" >> $O/c15_slim.log
( timeout 300 $TR tools/mgpu_ials_check.py ) > $O/c15_ials.log 2>&1; echo "ials rc=$?" >> $O/c15_ials.log
( timeout 200 $TR tools/mgpu_ease_check.py ) > $O/c15_ease.log 2>&1; echo "ease rc=$?" >> $O/c15_ease.log
( timeout 300 python -m pytest tests/test_multi_gpu.py -x -q -m gpu ) > $O/c15_pytest_mgpu.log 2>&1; echo "pytest rc=$?" >> $O/c15_pytest_mgpu.log
( timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --no-tensor > $O/c15_bench_n2.json ) 2> $O/c15_bench_n2.err; echo "bench rc=$?" >> $O/c15_bench_n2.err
for f in $O/c15_*.log; do echo "== $f"; grep -v "^\*\*\|OMP_NUM\|^$\|NCCL version" $f | tail -n 6; done
tail -c 1800 $O/c15_bench_n2.json; echo; tail -4 $O/c15_bench_n2.err
