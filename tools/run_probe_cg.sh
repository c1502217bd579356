#!/bin/bash
# cta_group::2 vs ::1 on representative shapes (correctness + speed)
P=./tools/probe
run() { echo "--- $*"; timeout 60 $P "$@"; echo "    exit=$?"; }
for cg in 2 1; do
run conv 9 64 64 2 40 40 0 0 0 0 -1 0 $cg
run conv 9 128 128 2 41 43 0 0 0 0 -1 0 $cg
run conv 1 64 64 2 40 40 0 0 0 0 -1 0 $cg
run conv 9 256 512 3 30 30 0 0 0 0 -1 0 $cg
run conv 9 64 64 2 60 60 0 0 2 0 -1 0 $cg
run conv 9 64 64 2 60 58 0 0 2 0 -1 0 $cg
run conv 9 256 256 3 24 24 0 0 0 0 -1 0 $cg
run conv 9 512 512 16 26 26 0 0 0 0 -1 0 $cg
run conv 1 128 256 2 40 40 0 0 1 0 -1 0 $cg
run conv 1 64 64 1 506 506 0 0 0 0 -1 0 $cg
run conv 9 64 64 1 510 510 0 0 0 0 -1 0 $cg
run conv 9 128 64 1 508 508 0 0 0 0 -1 0 $cg
run conv 9 64 128 2 252 252 0 0 0 0 -1 0 $cg
run conv 9 128 128 2 252 252 0 0 0 0 -1 0 $cg
run conv 9 256 128 2 252 252 0 0 0 0 -1 0 $cg
run conv 9 256 256 4 124 124 0 0 0 0 -1 0 $cg
run conv 9 512 256 4 124 124 0 0 0 0 -1 0 $cg
run conv 9 1024 512 8 60 60 0 0 0 0 -1 0 $cg
run conv 9 1024 1024 8 30 30 0 0 0 0 -1 0 $cg
run conv 1 1024 2048 8 28 28 0 0 1 0 -1 0 $cg
run conv 1 128 256 1 252 252 0 0 1 0 -1 0 $cg
done
