#!/bin/bash
for rep in 1 2; do
for bin in probe probe_epi2; do
  for cfg in "9 64 64 8 254 254 0 0 2" "9 64 64 8 254 254 0 0 0" "9 128 64 8 252 252 0 0 0"; do
    echo "$bin $cfg: $(timeout 60 ./tools/$bin conv $cfg | grep -E 'TFLOP|FAIL|PASS|failed' | tr '\n' ' ' | sed 's/checked=[0-9]* //' | cut -c1-150)"
  done
done
done
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --layers --no-cpu-baseline > gpurun_out/bench_v13.json 2> gpurun_out/bench_v13_layers.txt
