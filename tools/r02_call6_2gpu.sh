#!/bin/bash
# round 2, 2-GPU session: the NCCL / peer-memory paths of dist.py on hardware, then a short 2-GPU bench
export B200REC_SYNTH_CACHE=/dev/shm
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/c6_smi.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621"
( timeout 300 $TR tools/mgpu_check.py ) > $O/c6_sim.log 2>&1; echo "sim rc=$?" >> $O/c6_sim.log
( timeout 300 $TR tools/mgpu_bpr_check.py ) > $O/c6_bpr.log 2>&1; echo "bpr rc=$?" >> $O/c6_bpr.log
( timeout 300 $TR tools/mgpu_slim_check.py ) > $O/c6_slim.log 2>&1; echo "slim rc=$?" >> $O/c6_slim.log
( timeout 400 $TR tools/mgpu_ials_check.py ) > $O/c6_ials.log 2>&1; echo "ials rc=$?" >> $O/c6_ials.log
( timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --no-tensor > $O/c6_bench_n2.json ) 2> $O/c6_bench_n2.err; echo "bench rc=$?" >> $O/c6_bench_n2.err
( timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --no-tensor --no-bpr --no-cpu-baseline --gather nccl > $O/c6_bench_n2_nccl.json ) 2> $O/c6_bench_n2_nccl.err
for f in $O/c6_*.log; do echo "== $f"; tail -n 8 $f; done
tail -c 1500 $O/c6_bench_n2.json; echo; tail -5 $O/c6_bench_n2.err; tail -c 600 $O/c6_bench_n2_nccl.json
