#!/bin/bash
# Runs the stand-alone GPU probe suite (tools/probe) under per-test timeouts.
# Usage (on the GPU box): bash tools/run_probe.sh > gpurun_out/probe.log 2>&1
P=./tools/probe
run() { echo "--- $*"; timeout 60 $P "$@"; echo "    exit=$?"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
run desc 0 1024 0
run desc 3 1280 0
run desc 11 1280 0
run conv 9 64 64 2 40 40 0 0 0
run conv 1 64 64 2 40 40 0 0 0
run conv 9 128 128 2 40 40 0 0 0
run conv 9 128 256 2 30 30 0 0 0
run conv 9 256 512 3 30 30 0 0 0
run conv 9 128 64 4 150 150 0 0 0
run conv 9 64 64 2 60 60 0 0 2
run conv 1 128 256 2 40 40 0 0 1
run conv 1 256 1024 2 28 28 0 0 1
run conv 9 64 64 4 150 150 0 0 0 64 0
run conv 9 256 64 2 100 100 0 0 0
run conv 1 64 64 1 506 506 0 0 0
run conv 9 64 64 1 510 510 0 0 0
run conv 9 64 64 1 510 510 0 0 2
run conv 9 128 64 1 508 508 0 0 0
run conv 9 64 128 2 252 252 0 0 0
run conv 9 128 128 2 252 252 0 0 0
run conv 9 256 128 2 252 252 0 0 0
run conv 9 256 256 4 124 124 0 0 0
run conv 9 512 256 4 124 124 0 0 0
run conv 9 512 512 8 60 60 0 0 0
run conv 9 1024 512 8 60 60 0 0 0
run conv 9 1024 1024 8 30 30 0 0 0
run conv 1 1024 2048 8 28 28 0 0 1
run conv 1 128 256 1 252 252 0 0 1
