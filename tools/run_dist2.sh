#!/bin/bash
# N-GPU check of the sharded paths + bench line; run as: gpurun --gpus N -- bash tools/run_dist2.sh N
N=${1:-2}
set -x
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; tail -3 gpurun_out/gpu_tests.log; fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
timeout 300 $TR tools/dist_check.py > gpurun_out/dist_check$N.log 2>&1; echo "dist_check rc=$?" >> gpurun_out/dist_check$N.log
timeout 300 $TR tools/dist_phases.py > gpurun_out/dist_phases$N.log 2>&1; echo "rc=$?" >> gpurun_out/dist_phases$N.log
timeout 300 $TR tools/dist_phases.py 504 480 > gpurun_out/dist_phases${N}_504.log 2>&1; echo "rc=$?" >> gpurun_out/dist_phases${N}_504.log
timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_${N}gpu.log
tail -3 gpurun_out/dist_check$N.log; tail -12 gpurun_out/dist_phases$N.log; tail -2 gpurun_out/bench_${N}gpu.log | cut -c1-300
