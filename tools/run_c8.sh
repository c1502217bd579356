#!/bin/bash
P=./tools/probe
run() { echo "--- $*"; timeout 60 $P "$@" | grep -E "^CONV|TFLOP|PASS|FAIL|mismatch|failed" | head -8; }
run conv 9 8 64 2 40 40 0 0 0
run conv 9 8 64 3 41 37 0 0 0
run conv 9 8 64 1 508 508 0 0 0
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --no-cpu-baseline --layers 2>gpurun_out/b248.err | cut -c1-220; grep -E "gather|convs1.0" gpurun_out/b248.err
python bench.py --steps 10 --cs 504 --no-cpu-baseline --layers 2>gpurun_out/b504.err | cut -c1-220; grep -E "gather|convs1.0" gpurun_out/b504.err
