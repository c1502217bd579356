#!/bin/bash
# round 2, GPU call F: programmatic dependent launch on/off, full GPU tests
P=./tools/probe
run() { echo "--- $*: $(timeout 60 $P "$@" 2>&1 | grep -E 'TFLOP|FAIL|PASS|failed|error|mismatch|pool:|stray' | head -8 | tr '\n' ' ' | sed 's/checked=[0-9]* //; s/maxerr.*bad=/bad=/' | cut -c1-200)"; }
run conv 9 64 64 32 250 250 0 0 0
run conv 9 64 64 32 250 250 0 0 0 0 -1 0 0 -1 1 1
run conv 9 128 64 32 252 252 0 0 0
run conv 9 8 64 32 252 252 0 0 0
run conv 9 512 512 32 30 30 0 0 0
run conv 1 128 256 32 126 126 0 0 1
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "=== bench (pdl on = default / off), then small-batch forwards"
python bench.py --steps 5 --no-cpu-baseline --no-parity --images 2 | cut -c1-200
python bench.py --steps 5 --no-cpu-baseline --no-parity --images 2 --opt pdl=0 | cut -c1-200
python bench.py --steps 5 --no-cpu-baseline --no-parity --images 2 --batch 28 | cut -c1-200
python bench.py --steps 5 --no-cpu-baseline --no-parity --images 2 --batch 28 --opt pdl=0 | cut -c1-200
python bench.py --steps 5 --no-cpu-baseline --no-parity --images 2 --cs 120 | cut -c1-200
