#!/bin/bash
# round 2, GPU call C: per-tile pipeline traces of the level-1 kernels (TMA-store epilogue) + ncu of c8 and 64->64
P=./tools/probe
tr() { echo "--- $*"; NIND_TRACE=1 timeout 60 $P "$@" 2>&1 | grep -E "CONV|tps|TFLOP|trace|^ +[0-9]+ \||PASS|FAIL"; }
tr conv 9 64 64 32 250 250 0 0 0
tr conv 9 64 64 32 250 250 0 0 0 0 -1 0 0 -1 1
tr conv 9 128 64 32 252 252 0 0 0
tr conv 9 128 64 32 252 252 0 0 0 0 -1 0 0 -1 1
tr conv 9 64 64 32 254 254 0 0 2
tr conv 9 8 64 32 252 252 0 0 0
tr conv 1 128 256 32 126 126 0 0 1
tr conv 9 128 128 32 124 124 0 0 0
timeout 200 ncu --set full --clock-control none --import-source on -k regex:igemm -c 1 -f -o gpurun_out/r02b_c8 $P conv 9 8 64 32 252 252 0 0 0 > gpurun_out/r02b_c8_ncu.log 2>&1
echo "ncu c8 exit $?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:igemm -c 1 -f -o gpurun_out/r02b_6464 $P conv 9 64 64 32 250 250 0 0 0 > gpurun_out/r02b_6464_ncu.log 2>&1
echo "ncu 6464 exit $?"
