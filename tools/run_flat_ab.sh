#!/bin/bash
# flat (1-D) tiles vs 16x8 tiles on the narrow-map layer shapes; args: ... epi n_tile ws ctas cg flat
while read -r name cfg; do
  for flat in 0 1; do
    echo "$name flat $flat: $(timeout 60 ./tools/probe_flat conv $cfg 0 -1 0 0 $flat 2>&1 | grep -E 'TFLOP|FAIL|PASS|failed|error' | tr '\n' ' ' | sed 's/checked=[0-9]* //' | cut -c1-120)"
  done
done <<'CFG'
convs3.0 9 128 256 48 60 60 0 0 0
convs4.0 9 256 512 64 28 28 0 0 0
bottom.0 9 512 1024 96 12 12 0 0 0
bottom.2 9 1024 1024 96 14 14 0 0 0
up1 1 1024 2048 96 12 12 0 0 1
tconvs1.0 9 1024 512 64 28 28 0 0 0
tconvs1.2 9 512 512 64 30 30 0 0 0
up2 1 512 1024 64 28 28 0 0 1
tconvs2.0 9 512 256 48 60 60 0 0 0
tconvs2.2 9 256 256 48 62 62 0 0 0
odd1 9 64 64 3 19 23 0 0 0
odd2 9 128 128 5 11 37 0 0 0
odd3 1 64 256 3 9 13 0 0 1
CFG
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --layers --no-cpu-baseline > gpurun_out/bench_v14.json 2> gpurun_out/bench_v14_layers.txt
