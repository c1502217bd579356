#!/bin/bash
# round 2, 8-GPU call: host-path timeline, BASELINE configs 2 / 4 / 5 on N GPUs (default bench with 64 streamed images,
# UNet 45 MP cs 512, crop-size x overlap sweep).
N=${1:-8}
export OMP_NUM_THREADS=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "=== host-path timeline"
timeout 300 $TR tools/dist_phases_host.py > gpurun_out/r2n_phases_host$N.log 2>&1; grep -E "^world|^rank|^ +[0-9]+ +[0-9]+|total|Error|error" gpurun_out/r2n_phases_host$N.log | head -30
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    p = d.get("parity") or {}
    print("value %.1f  e2e %.1f  ms %.3f  stream %s  parity %s %s" % (d["value"], d["e2e"]["value"] if d.get("e2e") else -1, d["ms_per_step"],
          (d.get("throughput_mode") or {}).get("value"), (p.get("distributed_vs_single") or {}).get("pass"), (p.get("vs_oracle") or {}).get("pass")))
except Exception as e:
    print("bench failed:", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
}
echo "=== bench --gpus $N (UtNet cs 248, 64 streamed images)"
timeout 600 $TR bench.py --gpus $N --steps 10 --images 8 > gpurun_out/r2n_bench${N}.json 2> gpurun_out/r2n_bench${N}.err; summ gpurun_out/r2n_bench${N}.json
echo "=== bench --gpus $N --network UNet (45 MP, cs 512)"
timeout 600 $TR bench.py --gpus $N --network UNet --steps 5 --images 2 --parity-crops 2 > gpurun_out/r2n_bench${N}_unet.json 2> gpurun_out/r2n_bench${N}_unet.err; summ gpurun_out/r2n_bench${N}_unet.json
echo "=== bench --gpus $N --sweep"
timeout 600 $TR bench.py --gpus $N --sweep --images 3 > gpurun_out/r2n_sweep${N}.json 2> gpurun_out/r2n_sweep${N}.err; grep -c overlap gpurun_out/r2n_sweep${N}.err; cut -c1-300 gpurun_out/r2n_sweep${N}.json
