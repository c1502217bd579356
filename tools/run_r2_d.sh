#!/bin/bash
# round 2, GPU call D: dual MMA issuer + multi-buffered TMA-store staging + 4-set pixel-pair kernel
P=./tools/probe
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
run() { echo "--- $*: $(timeout 60 $P "$@" 2>&1 | grep -E 'TFLOP|FAIL|PASS|failed|error|mismatch|pool:|stray' | head -8 | tr '\n' ' ' | sed 's/checked=[0-9]* //' | cut -c1-420)"; }
#   taps cin n B Hs Ws 0 0 epi [n_tile ws ctas cg flat pair pool dual]
run conv 9 64 64 2 40 40 0 0 0
run conv 9 64 64 2 40 40 0 0 0 0 -1 0 1
run conv 9 64 64 3 41 37 0 0 0
run conv 9 64 64 2 40 40 0 0 0 0 -1 0 0 -1 0 1
run conv 9 64 64 2 40 40 0 0 0 0 -1 0 0 -1 1 0
run conv 9 64 64 2 40 40 0 0 0 0 -1 0 0 -1 1 1
run conv 9 64 64 3 22 38 0 0 0 0 -1 0 0 -1 1 1
run conv 9 64 64 2 60 60 0 0 2
run conv 9 64 64 2 60 60 0 0 2 0 -1 0 0 -1 1
run conv 9 128 64 2 40 44 0 0 0
run conv 9 128 128 2 40 40 0 0 0
run conv 9 128 128 2 40 40 0 0 0 0 -1 0 0 -1 0 1
run conv 9 64 128 3 34 50 0 0 0 0 -1 0 1 -1 0 1
run conv 9 128 256 2 30 30 0 0 0
run conv 9 256 512 3 30 30 0 0 0
run conv 9 256 512 3 30 30 0 0 0 0 -1 0 0 0
run conv 9 512 512 8 60 60 0 0 0
run conv 1 128 256 2 40 40 0 0 1
run conv 1 128 256 2 41 37 0 0 1 0 -1 0 0 0
run conv 1 256 1024 2 28 28 0 0 1
run conv 1 1024 2048 2 28 28 0 0 1 0 -1 0 0 0
run conv 9 8 64 2 40 40 0 0 0
run conv 9 8 64 3 41 37 0 0 0
run conv 9 64 64 1 24 24 0 0 0
run conv 9 64 128 1 60 60 0 0 0
run conv 9 64 128 2 62 62 0 0 0 0 -1 0 1 1
run conv 9 8 64 1 20 20 0 0 0
echo "=== timing (B = 32): dual issuer off / on"
for d in 0 1; do
run conv 9 64 64 32 250 250 0 0 0 0 -1 0 0 -1 0 0 $d
run conv 9 64 64 32 250 250 0 0 0 0 -1 0 0 -1 1 0 $d
run conv 9 64 64 32 250 250 0 0 0 0 -1 0 0 -1 1 1 $d
run conv 9 64 64 32 250 250 0 0 0 0 -1 0 0 -1 0 1 $d
run conv 9 128 64 32 252 252 0 0 0 0 -1 0 0 -1 0 0 $d
run conv 9 64 64 32 254 254 0 0 2 0 -1 0 0 -1 0 0 $d
run conv 9 64 64 32 254 254 0 0 2 0 -1 0 0 -1 1 0 $d
run conv 9 8 64 32 252 252 0 0 0 0 -1 0 0 -1 0 0 $d
run conv 1 128 256 32 126 126 0 0 1 0 -1 0 0 -1 0 0 $d
run conv 9 64 128 32 126 126 0 0 0 0 -1 0 0 -1 0 0 $d
run conv 9 128 128 32 124 124 0 0 0 0 -1 0 0 -1 0 0 $d
run conv 9 256 128 32 128 128 0 0 0 0 -1 0 0 -1 0 0 $d
run conv 9 256 256 32 60 60 0 0 0 0 -1 0 0 -1 0 0 $d
run conv 9 512 512 32 30 30 0 0 0 0 -1 0 0 -1 0 0 $d
done
echo "=== traces"
tr() { echo "--- $*"; NIND_TRACE=1 timeout 60 $P "$@" 2>&1 | grep -E "CONV|TFLOP|^ +[0-9]+ \|" | head -14; }
tr conv 9 64 64 32 250 250 0 0 0
tr conv 9 64 64 32 250 250 0 0 0 0 -1 0 0 -1 1
tr conv 9 128 64 32 252 252 0 0 0
tr conv 9 8 64 32 252 252 0 0 0
echo "=== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "=== bench (default / dual off / pair off)"
python bench.py --steps 5 --no-cpu-baseline --layers 2> gpurun_out/r2d_layers.txt | cut -c1-260
python bench.py --steps 5 --no-cpu-baseline --layers --opt dual_issuer=0 2> gpurun_out/r2d_layers_dual0.txt | cut -c1-260
python bench.py --steps 5 --no-cpu-baseline --layers --opt pair64=0 2> gpurun_out/r2d_layers_pair0.txt | cut -c1-260
python bench.py --steps 5 --no-cpu-baseline --cs 504 --no-parity | cut -c1-260
python bench.py --steps 3 --no-cpu-baseline --network UNet --no-parity | cut -c1-260
