#!/bin/bash
# round 2, multi-GPU call for the host-buffer entry: N = number of GPUs ($1).  Ownership flipped (the earlier rank owns
# the shared grid row), hand-over by peer DMA after the first step.
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "=== pytest (host range / multi-GPU tests)"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "host_range or nccl or another_gpu" 2>&1 | tail -4
echo "=== host-path timeline"
timeout 300 $TR tools/dist_phases_host.py > gpurun_out/r2n_phases_host$N.log 2>&1; grep -E "^world|^rank|^ +[0-9]+ +[0-9]+|total|Error|error" gpurun_out/r2n_phases_host$N.log | head -30
echo "=== bench --gpus $N"
timeout 600 $TR bench.py --gpus $N --steps 10 --images 2 > gpurun_out/r2n_bench${N}.json 2> gpurun_out/r2n_bench${N}.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2n_bench${N}.json").read().strip().splitlines()[-1])
    print("value %.1f  e2e %.1f  ms %.3f  parity %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], json.dumps(d["parity"]["distributed_vs_single"])))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r2n_bench${N}.err").read()[-1500:])
PY
