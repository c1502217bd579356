#!/bin/bash
# 2-GPU check of the sharded paths + bench line; run as: gpurun --gpus 2 -- bash tools/run_dist2.sh
set -x
python tools/dist_breakdown.py > gpurun_out/dist_breakdown.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531"
timeout 300 $TR tools/dist_check.py > gpurun_out/dist_check2.log 2>&1; echo "dist_check rc=$?" >> gpurun_out/dist_check2.log
timeout 400 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_2gpu.log
tail -3 gpurun_out/dist_check2.log; tail -2 gpurun_out/bench_2gpu.log | cut -c1-600
