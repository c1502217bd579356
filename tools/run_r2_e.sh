#!/bin/bash
# round 2, GPU call E: dual issuer with the idle lanes parked (syncwarp restored): timing A/B, tests, benches
P=./tools/probe
run() { echo "--- $*: $(timeout 60 $P "$@" 2>&1 | grep -E 'TFLOP|FAIL|PASS|failed|error|mismatch|pool:|stray' | head -8 | tr '\n' ' ' | sed 's/checked=[0-9]* //; s/maxerr.*bad=/bad=/' | cut -c1-200)"; }
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
run conv 1 512 1024 32 60 60 0 0 1 0 -1 0 0 -1 0 0 $d
done
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "=== bench (default / dual off / pair off / cs 504 / UNet)"
python bench.py --steps 5 --no-cpu-baseline --layers 2> gpurun_out/r2e_layers.txt > gpurun_out/r2e_bench.json; cut -c1-200 gpurun_out/r2e_bench.json
python bench.py --steps 5 --no-cpu-baseline --no-parity --layers --opt dual_issuer=0 2> gpurun_out/r2e_layers_dual0.txt | cut -c1-200
python bench.py --steps 5 --no-cpu-baseline --no-parity --layers --opt pair64=0 2> gpurun_out/r2e_layers_pair0.txt | cut -c1-200
python bench.py --steps 5 --no-cpu-baseline --cs 504 > gpurun_out/r2e_bench_504.json; cut -c1-200 gpurun_out/r2e_bench_504.json
python bench.py --steps 3 --no-cpu-baseline --network UNet > gpurun_out/r2e_bench_unet.json; cut -c1-200 gpurun_out/r2e_bench_unet.json
