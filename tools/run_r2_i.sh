#!/bin/bash
# round 2, GPU call I: WIDE epilogue staging (128-byte rows, one TMA store per 64 channels) on the store-bound layers.
run() { P=$1; shift; echo "--- $(basename $P) $*: $(timeout 60 $P "$@" 2>&1 | grep -E 'TFLOP|FAIL|PASS|failed|error|mismatch|pool:|stray' | head -8 | tr '\n' ' ' | sed 's/checked=[0-9]* //; s/maxerr.*bad=/bad=/' | cut -c1-200)"; }
for P in ./tools/probe_base ./tools/probe; do
run $P conv 9 8 64 133 252 252 0 0 0
run $P conv 1 128 256 133 126 126 0 0 1
run $P conv 1 256 512 133 62 62 0 0 1
run $P conv 1 512 1024 133 30 30 0 0 1
run $P conv 1 1024 2048 133 14 14 0 0 1
done
echo "=== wide off / on where it is not the default"
for w in 0 1; do
run ./tools/probe conv 9 128 128 32 124 124 0 0 0 0 -1 0 0 -1 0 0 -1 $w
run ./tools/probe conv 9 256 128 32 128 128 0 0 0 0 -1 0 0 -1 0 0 -1 $w
run ./tools/probe conv 9 128 64 32 252 252 0 0 0 0 -1 0 0 -1 0 0 -1 $w
run ./tools/probe conv 9 256 256 32 60 60 0 0 0 0 -1 0 0 -1 0 0 -1 $w
done
echo "=== correctness on odd shapes"
run ./tools/probe conv 1 128 256 2 41 37 0 0 1
run ./tools/probe conv 1 128 256 3 40 40 0 0 1
run ./tools/probe conv 1 256 1024 2 28 28 0 0 1
run ./tools/probe conv 9 8 64 3 41 37 0 0 0
run ./tools/probe conv 9 128 128 3 41 37 0 0 0 0 -1 0 0 -1 0 0 -1 1
run ./tools/probe conv 1 64 64 3 41 37 0 0 0
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "=== A/B bench"
bash tools/run_ab.sh 2
