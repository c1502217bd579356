#!/bin/bash
# round 2, GPU call B: TMA-store epilogue + permanent pixel-pair mode: probe correctness sweep, GPU tests, bench A/B
P=./tools/probe
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
run() { echo "--- $*: $(timeout 60 $P "$@" 2>&1 | grep -E 'TFLOP|FAIL|PASS|failed|error|mismatch|pool:|stray' | head -8 | tr '\n' ' ' | sed 's/checked=[0-9]* //' | cut -c1-420)"; }
#   taps cin n B Hs Ws 0 0 epi [n_tile ws ctas cg flat pair pool]
run conv 9 64 64 2 40 40 0 0 0
run conv 9 64 64 2 40 40 0 0 0 0 -1 0 1
run conv 9 64 64 3 41 37 0 0 0
run conv 9 64 64 2 40 40 0 0 0 0 -1 0 0 -1 0 1
run conv 9 64 64 2 40 40 0 0 0 0 -1 0 0 -1 1 0
run conv 9 64 64 2 40 40 0 0 0 0 -1 0 0 -1 1 1
run conv 9 64 64 3 22 38 0 0 0 0 -1 0 0 -1 1 1
run conv 9 128 64 2 40 44 0 0 0 0 -1 0 0 -1 1 0
run conv 9 64 64 2 60 60 0 0 2
run conv 9 64 64 2 60 60 0 0 2 0 -1 0 0 -1 1
run conv 9 128 128 2 40 40 0 0 0
run conv 9 128 128 2 40 40 0 0 0 0 -1 0 0 -1 0 1
run conv 9 64 128 3 34 50 0 0 0 0 -1 0 1 -1 0 1
run conv 9 128 256 2 30 30 0 0 0
run conv 9 256 512 3 30 30 0 0 0
run conv 9 256 512 3 30 30 0 0 0 0 -1 0 0 0
run conv 1 128 256 2 40 40 0 0 1
run conv 1 128 256 2 41 37 0 0 1 0 -1 0 0 0
run conv 1 256 1024 2 28 28 0 0 1
run conv 1 1024 2048 2 28 28 0 0 1 0 -1 0 0 0
run conv 9 8 64 2 40 40 0 0 0
run conv 9 8 64 3 41 37 0 0 0
echo "=== timing, level-1 shapes (B = 32)"
run conv 9 64 64 32 250 250 0 0 0
run conv 9 64 64 32 250 250 0 0 0 0 -1 0 0 -1 1
run conv 9 64 64 32 250 250 0 0 0 0 -1 0 0 -1 1 1
run conv 9 64 64 32 250 250 0 0 0 0 -1 0 0 -1 0 1
run conv 9 128 64 32 252 252 0 0 0
run conv 9 128 64 32 252 252 0 0 0 0 -1 0 0 -1 1
run conv 9 64 64 32 254 254 0 0 2
run conv 9 64 64 32 254 254 0 0 2 0 -1 0 0 -1 1
run conv 9 8 64 32 252 252 0 0 0
run conv 1 128 256 32 126 126 0 0 1
run conv 9 64 128 32 126 126 0 0 0
run conv 9 128 128 32 124 124 0 0 0
run conv 9 256 128 32 128 128 0 0 0
echo "=== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "=== bench (pair64 on = default / off)"
python bench.py --steps 5 --no-cpu-baseline --layers 2> gpurun_out/r2b_layers_pair1.txt | cut -c1-300
python bench.py --steps 5 --no-cpu-baseline --layers --opt pair64=0 2> gpurun_out/r2b_layers_pair0.txt | cut -c1-300
