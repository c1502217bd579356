#!/bin/bash
# Runs the stand-alone GPU probe suite (tools/probe) under per-test timeouts.
# Usage (on the GPU box): bash tools/run_probe.sh > gpurun_out/probe.log 2>&1
P=./tools/probe
run() { echo "--- $*"; timeout 60 $P "$@"; echo "    exit=$?"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
# 1. descriptor semantics
run desc 0 1024 0
run desc 8 1024 0
run desc 1 1024 0
run desc 3 1280 0
run desc 11 1280 0
run desc 1 1024 1
run desc 3 1280 1
run desc 11 1280 1
# 2. conv kernel, a_mode 1 (aligned copies) — depends only on basic descriptor semantics
run conv 9 64 64 2 40 40 1 0 0
run conv 1 64 64 2 40 40 0 0 0
# 3. conv kernel, a_mode 0 (row-offset descriptors)
run conv 9 64 64 2 40 40 0 0 0
run conv 9 64 64 2 40 40 0 1 0
# 4. wider / deeper shapes in both modes
for m in 0 1; do
  run conv 9 128 128 2 40 40 $m 0 0
  run conv 9 128 256 2 30 30 $m 0 0
  run conv 9 256 512 3 30 30 $m 0 0
  run conv 9 128 64 4 150 150 $m 0 0
  run conv 9 64 64 2 60 60 $m 0 2
done
run conv 1 128 256 2 40 40 0 0 1
run conv 1 256 1024 2 28 28 0 0 1
run conv 9 64 64 4 150 150 0 0 0 64 0
run conv 9 64 64 1 510 510 0 0 0
run conv 9 128 64 1 508 508 0 0 0
run conv 9 256 128 2 252 252 0 0 0
run conv 9 512 256 4 124 124 0 0 0
run conv 9 1024 512 8 60 60 0 0 0
