#!/bin/bash
# round 2, GPU call H: deeper activation ring for the first layer (24 stages), one TMA store per 4 rows in the
# depth-to-space epilogue (folded tensor map), 64->128 on CTA pairs re-tested.  Same-box A/B against tools/probe_base.
run() { P=$1; shift; echo "--- $(basename $P) $*: $(timeout 60 $P "$@" 2>&1 | grep -E 'TFLOP|FAIL|PASS|failed|error|mismatch|pool:|stray' | head -8 | tr '\n' ' ' | sed 's/checked=[0-9]* //; s/maxerr.*bad=/bad=/' | cut -c1-200)"; }
for P in ./tools/probe_base ./tools/probe; do
run $P conv 9 8 64 32 252 252 0 0 0
run $P conv 1 128 256 32 126 126 0 0 1
run $P conv 1 256 512 32 62 62 0 0 1
run $P conv 1 512 1024 32 30 30 0 0 1
run $P conv 1 1024 2048 32 14 14 0 0 1
run $P conv 9 64 128 32 126 126 0 0 0
run $P conv 9 64 128 32 126 126 0 0 0 0 -1 0 2
done
echo "=== correctness on odd shapes"
run ./tools/probe conv 1 128 256 2 41 37 0 0 1
run ./tools/probe conv 1 128 256 3 40 40 0 0 1
run ./tools/probe conv 1 256 1024 2 28 28 0 0 1
run ./tools/probe conv 9 8 64 3 41 37 0 0 0
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "=== bench"
python bench.py --steps 5 --no-cpu-baseline --layers 2> gpurun_out/r2h_layers.txt > gpurun_out/r2h_bench.json; cut -c1-200 gpurun_out/r2h_bench.json
