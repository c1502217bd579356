#!/bin/bash
# 8-GPU A/B of the NUMA binding (process affinity + first touch of the shared host image): host-path timeline and bench e2e
N=${1:-8}
export OMP_NUM_THREADS=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi topo -m 2>/dev/null | head -14
for mode in bind nobind; do
  if [ $mode = nobind ]; then export NIND_NO_NUMA_BIND=1; else unset NIND_NO_NUMA_BIND; fi
  echo "=== host-path timeline ($mode)"
  timeout 300 $TR tools/dist_phases_host.py > gpurun_out/r2p_phases_host${N}_$mode.log 2>&1; grep -E "^world|^rank +crops|^ +[0-9]+ +[0-9]+|total|Error|error" gpurun_out/r2p_phases_host${N}_$mode.log | head -30
done
unset NIND_NO_NUMA_BIND
grep "bound to" gpurun_out/r2p_phases_host${N}_bind.log | head -8
echo "=== bench --gpus $N"
timeout 600 $TR bench.py --gpus $N --steps 10 --images 2 --no-parity > gpurun_out/r2p_bench${N}.json 2> gpurun_out/r2p_bench${N}.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2p_bench${N}.json").read().strip().splitlines()[-1])
print("value %.1f  e2e %.1f  ms %.3f stream %.1f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["throughput_mode"]["value"]))
PY
