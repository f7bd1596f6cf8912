#!/bin/bash
# 4-GPU check of the bench (peer-store table + BPR delta exchange beyond 2 ranks) and of the K1 checker.
export B200REC_SYNTH_CACHE=/dev/shm
O=gpurun_out
mkdir -p $O
( timeout 220 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 4 --steps 5 --warmup 3 --no-tensor ) > $O/c20_bench_n4.json 2> $O/c20_bench_n4.err; echo "bench4 rc=$?"
tail -n 3 $O/c20_bench_n4.err | cut -c1-300
( timeout 70 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29612 tools/mgpu_check.py ) > $O/c20_mgpu_check.log 2>&1; echo "mgpu_check rc=$?"
tail -n 4 $O/c20_mgpu_check.log | cut -c1-300
