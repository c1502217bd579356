#!/bin/bash
# round 2, multi-GPU call: N = number of GPUs of this box ($1).  NCCL / peer-memory parity under pytest, per-rank
# phase break-down of the host-buffer entry, bench with the peer-DMA gather and with the NCCL send/recv gather.
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
echo "=== host-path phases"
timeout 300 $TR tools/dist_phases_host.py > gpurun_out/r2m_phases_host$N.log 2>&1; grep -E "^world|^rank|^ +[0-9]+ +[0-9]+|un-instrumented" gpurun_out/r2m_phases_host$N.log
echo "=== device-path phases"
timeout 300 $TR tools/dist_phases.py > gpurun_out/r2m_phases_dev$N.log 2>&1; grep -E "^world|ms  ->|max diff" gpurun_out/r2m_phases_dev$N.log
echo "=== bench --gpus $N (peer gather)"
timeout 600 $TR bench.py --gpus $N --steps 10 > gpurun_out/r2m_bench${N}_peer.json 2> gpurun_out/r2m_bench${N}_peer.err; tail -c 2500 gpurun_out/r2m_bench${N}_peer.json | cut -c1-2500; tail -3 gpurun_out/r2m_bench${N}_peer.err
echo "=== bench --gpus $N --gather rows"
timeout 600 $TR bench.py --gpus $N --steps 10 --gather rows --images 2 > gpurun_out/r2m_bench${N}_rows.json 2> gpurun_out/r2m_bench${N}_rows.err; cut -c1-400 gpurun_out/r2m_bench${N}_rows.json
echo "=== pytest (multi-GPU tests)"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "nccl or another_gpu" 2>&1 | tail -4
